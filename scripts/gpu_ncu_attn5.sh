#!/bin/bash
# ncu --set full of flash_attn5 on the FLUX joint-attention shape
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn5 -s 2 -c 1 -f -o gpurun_out/ncu_attn5 python scripts/attn_probe.py 4 24 4608 4608 128 > gpurun_out/ncu_attn5.log 2>&1
python profiles/extract_ncu.py gpurun_out/ncu_attn5.ncu-rep > gpurun_out/ncu_attn5.txt 2>&1
ncu -i gpurun_out/ncu_attn5.ncu-rep --page source --csv > gpurun_out/ncu_attn5_source.csv 2>/dev/null
python profiles/source_hotspots.py gpurun_out/ncu_attn5_source.csv samples 45 >> gpurun_out/ncu_attn5.txt
cat gpurun_out/ncu_attn5.txt | cut -c1-165
