#!/bin/bash
# Cout = 256 through the rolling-row kernel (fused input GroupNorm): parity, then stage timings with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "rolling_rows or conv_halo or fused_input" > gpurun_out/wide_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/wide_tests.log | cut -c1-220
timeout 300 python scripts/s512_stages.py 2>&1 | tail -1 | sed "s/^/wide on  S: /"
OMGSR_NO_WIDE=1 timeout 300 python scripts/s512_stages.py 2>&1 | tail -1 | sed "s/^/wide off S: /"
timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-200 | sed "s/^/wide on  F: /"
OMGSR_NO_WIDE=1 timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-200 | sed "s/^/wide off F: /"
