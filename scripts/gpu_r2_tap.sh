#!/bin/bash
# residual as an identity tap in the rolling-row kernel: parity, A/B (OMGSR_HALO_RES_TAP=0|1), stages
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "rolling_rows or conv_halo or fused_input" > gpurun_out/tap_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/tap_tests.log | cut -c1-200
for tap in 0 1; do echo "== OMGSR_HALO_RES_TAP=$tap"; OMGSR_HALO_RES_TAP=$tap timeout 200 python scripts/halo_ab.py 2>&1 | head -7; done > gpurun_out/tap_halo_ab.txt 2>&1; cat gpurun_out/tap_halo_ab.txt
for tap in 0 1; do OMGSR_HALO_RES_TAP=$tap timeout 300 python scripts/s512_stages.py 2>&1 | tail -1 | sed "s/^/TAP=$tap S: /"; done
