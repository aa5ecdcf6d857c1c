#!/bin/bash
# same-box in-step A/B: current library (rolling rows + warp-uniform MMA issue) vs the library built from the previous
# commit (omgsr_b200/lib/ab/libomgsr_b200_head.so: single-lane issue, row-shared halo kernel)
mkdir -p gpurun_out
L=omgsr_b200/lib/libomgsr_b200.so
for rep in 1 2; do
  timeout 300 python scripts/s512_stages.py 2>&1 | tail -1 | sed "s/^/new  S $rep: /"
  timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-230 | sed "s/^/new  F $rep: /"
  cp $L /tmp/lib_new.so; cp omgsr_b200/lib/ab/libomgsr_b200_head.so $L
  timeout 300 python scripts/s512_stages.py 2>&1 | tail -1 | sed "s/^/head S $rep: /"
  timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-230 | sed "s/^/head F $rep: /"
  cp /tmp/lib_new.so $L
done > gpurun_out/r2t_instep_ab.txt 2>&1
cat gpurun_out/r2t_instep_ab.txt
