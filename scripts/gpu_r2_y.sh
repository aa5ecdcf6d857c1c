#!/bin/bash
# exp2 on the FMA pipe for a fraction of the softmax (flash_attn3): parity at the largest fraction, then timing per fraction
mkdir -p gpurun_out
OMGSR_ATTN_POLY=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "flash or attention" > gpurun_out/r2y_tests.log 2>&1
echo "pytest (poly=2) rc=$?"; tail -3 gpurun_out/r2y_tests.log | cut -c1-200
for poly in 0 1 2; do echo "== OMGSR_ATTN_POLY=$poly"; OMGSR_ATTN_POLY=$poly timeout 300 python scripts/attn_ab.py 2>&1 | head -3 | cut -c1-140; done > gpurun_out/r2y_attn_poly.txt 2>&1; cat gpurun_out/r2y_attn_poly.txt
for poly in 0 2; do OMGSR_ATTN_POLY=$poly timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-200 | sed "s/^/POLY=$poly F: /"; done
