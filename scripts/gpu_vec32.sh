#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "== kernel tests exit $?" > gpurun_out/summary_k.txt; tail -3 gpurun_out/k_all.log >> gpurun_out/summary_k.txt
bash scripts/gpu_smallk.sh > /dev/null 2>&1
grep -E "force_bn=0" gpurun_out/summary.txt > gpurun_out/summary_smallk.txt
timeout 300 python scripts/s512_stages.py > gpurun_out/stages.txt 2>&1
cat gpurun_out/summary_k.txt gpurun_out/summary_smallk.txt gpurun_out/stages.txt
