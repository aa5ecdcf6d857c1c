#!/bin/bash
bash scripts/gpu_final.sh > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_pair128_gn python scripts/conv_probe.py halo_gn > gpurun_out/ncu1b.log 2>&1
python profiles/extract_ncu.py gpurun_out/ncu_halo_pair128_gn.ncu-rep > gpurun_out/ncu_halo_pair128_gn.txt 2>&1
timeout 300 python scripts/s512_stages.py > gpurun_out/stages.txt 2>&1
cat gpurun_out/summary.txt gpurun_out/stages.txt; head -16 gpurun_out/ncu_halo_pair128_gn.txt
