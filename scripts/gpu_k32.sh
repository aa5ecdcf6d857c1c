#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "== kernel tests exit $?" > gpurun_out/summary.txt; tail -4 gpurun_out/k_all.log >> gpurun_out/summary.txt
timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
cat gpurun_out/summary.txt
