#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2m_tests.log | cut -c1-200
timeout 300 python scripts/f1024.py 4 4 > gpurun_out/r2m_f1024_stages.txt 2>&1; tail -1 gpurun_out/r2m_f1024_stages.txt
