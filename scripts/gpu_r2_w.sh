#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_rr_gn_res2 python scripts/conv_probe.py halo_gn_res > gpurun_out/ncu_h.log 2>&1
python profiles/extract_ncu.py gpurun_out/ncu_halo_rr_gn_res2.ncu-rep > gpurun_out/ncu_halo_rr_gn_res2_summary.txt 2>&1
ncu -i gpurun_out/ncu_halo_rr_gn_res2.ncu-rep --page source --csv > gpurun_out/ncu_halo_rr_gn_res2_source.csv 2>/dev/null
python profiles/source_hotspots.py gpurun_out/ncu_halo_rr_gn_res2_source.csv samples 30 >> gpurun_out/ncu_halo_rr_gn_res2_summary.txt
cat gpurun_out/ncu_halo_rr_gn_res2_summary.txt | cut -c1-170
