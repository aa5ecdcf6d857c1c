#!/bin/bash
# ncu --set full of flash_attn3 on the FLUX joint-attention shape
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn3 -s 2 -c 1 -f -o gpurun_out/ncu_attn3 python scripts/attn_probe.py 4 24 4608 4608 128 > gpurun_out/ncu_attn3.log 2>&1
python profiles/extract_ncu.py gpurun_out/ncu_attn3.ncu-rep > gpurun_out/ncu_attn3.txt 2>&1
ncu -i gpurun_out/ncu_attn3.ncu-rep --page source --csv > gpurun_out/ncu_attn3_source.csv 2>/dev/null
python profiles/source_hotspots.py gpurun_out/ncu_attn3_source.csv samples 45 >> gpurun_out/ncu_attn3.txt
cat gpurun_out/ncu_attn3.txt | cut -c1-165
