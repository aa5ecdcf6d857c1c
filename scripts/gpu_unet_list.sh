#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_only.py > gpurun_out/ncu_unet.log 2>&1
echo "== launch list exit $?" > gpurun_out/summary.txt
python profiles/summarise_launches.py gpurun_out/launches_unet.csv 2 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
