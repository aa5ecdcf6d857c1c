#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "attention or attn or flash" > gpurun_out/r2h_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2h_tests.log
timeout 300 python scripts/attn_ab.py > gpurun_out/r2h_attn.txt 2>&1; cat gpurun_out/r2h_attn.txt
