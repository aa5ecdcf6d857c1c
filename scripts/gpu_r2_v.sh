#!/bin/bash
# rolling-row kernel with a 7-deep row ring / per-tap-column weight stages: parity, timing, ncu of the light XF variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "rolling_rows or conv_halo or fused_input" > gpurun_out/r2v_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2v_tests.log | cut -c1-200
timeout 200 python scripts/halo_ab.py > gpurun_out/r2v_halo.txt 2>&1; cat gpurun_out/r2v_halo.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_rr_gn2 python scripts/conv_probe.py halo_gn > gpurun_out/ncu_h.log 2>&1
python profiles/extract_ncu.py gpurun_out/ncu_halo_rr_gn2.ncu-rep > gpurun_out/ncu_halo_rr_gn2_summary.txt 2>&1
ncu -i gpurun_out/ncu_halo_rr_gn2.ncu-rep --page source --csv > gpurun_out/ncu_halo_rr_gn2_source.csv 2>/dev/null
python profiles/source_hotspots.py gpurun_out/ncu_halo_rr_gn2_source.csv samples 16 >> gpurun_out/ncu_halo_rr_gn2_summary.txt
cat gpurun_out/ncu_halo_rr_gn2_summary.txt | cut -c1-170
timeout 300 python scripts/s512_stages.py > gpurun_out/r2v_s512_stages.txt 2>&1; tail -1 gpurun_out/r2v_s512_stages.txt
