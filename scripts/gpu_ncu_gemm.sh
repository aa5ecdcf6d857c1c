#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/ncu_gemm_k64 python scripts/gemm_probe.py 65536 1280 64 0 0 nobias > gpurun_out/ncu_g.log 2>&1
echo "== ncu exit $?" > gpurun_out/summary.txt
python profiles/extract_ncu.py gpurun_out/ncu_gemm_k64.ncu-rep > gpurun_out/ncu_gemm_k64.txt 2>&1
ncu -i gpurun_out/ncu_gemm_k64.ncu-rep --page source --csv > gpurun_out/ncu_gemm_k64_source.csv 2>/dev/null
cat gpurun_out/ncu_gemm_k64.txt >> gpurun_out/summary.txt
python profiles/source_hotspots.py gpurun_out/ncu_gemm_k64_source.csv samples 30 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
