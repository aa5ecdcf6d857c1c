#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "materialised or flash" > gpurun_out/r2z_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2z_tests.log | cut -c1-200
for rep in 1 2; do timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-200; done
timeout 300 python scripts/s512_stages.py 2>&1 | tail -1
