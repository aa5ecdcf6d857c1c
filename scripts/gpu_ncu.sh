#!/bin/bash
mkdir -p gpurun_out
for m in generic halo; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 2 -c 1 -f -o gpurun_out/conv128_$m python scripts/conv_probe.py $m > gpurun_out/ncu_$m.log 2>&1
  echo "== ncu $m exit $?" >> gpurun_out/summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 2 -c 1 -f -o gpurun_out/conv512_generic python scripts/conv_probe.py generic 16 128 128 512 512 > gpurun_out/ncu_512.log 2>&1
echo "== ncu 512 exit $?" >> gpurun_out/summary.txt
ls -la gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
