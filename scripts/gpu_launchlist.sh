#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== launch list exit $?" > gpurun_out/summary.txt
python profiles/summarise_launches.py gpurun_out/launches.csv 4 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
