#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn2 -s 2 -c 1 -f -o gpurun_out/ncu_flash64 python scripts/attn_probe.py 16 5 4096 4096 64 > gpurun_out/ncu_a.log 2>&1
echo "== ncu exit $?" > gpurun_out/summary.txt
python profiles/extract_ncu.py gpurun_out/ncu_flash64.ncu-rep > gpurun_out/ncu_flash64.txt 2>&1
ncu -i gpurun_out/ncu_flash64.ncu-rep --page source --csv > gpurun_out/ncu_flash64_source.csv 2>/dev/null
cat gpurun_out/ncu_flash64.txt >> gpurun_out/summary.txt
python profiles/source_hotspots.py gpurun_out/ncu_flash64_source.csv samples 40 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
