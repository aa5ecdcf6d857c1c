#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f1024.csv python scripts/f1024.py 4 1 > gpurun_out/ncu_f.log 2>&1
echo "== ncu f1024 exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/ncu_f.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
