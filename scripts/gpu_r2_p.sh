#!/bin/bash
# rolling-row halo kernel (conv_halo_rr_kernel): parity first (under a short timeout: a pipeline bug hangs), then A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "rolling_rows" > gpurun_out/r2p_tests_rr.log 2>&1
echo "pytest rr rc=$?"; grep -E "passed|failed|FAILED|max_rel|Error" gpurun_out/r2p_tests_rr.log | head -20 | cut -c1-220
for mode in 0 1; do
  echo "== OMGSR_HALO_ROLLING=$mode"
  OMGSR_HALO_ROLLING=$mode timeout 200 python scripts/halo_ab.py 2>&1 | head -6
done > gpurun_out/r2p_halo_rolling_ab.txt 2>&1; cat gpurun_out/r2p_halo_rolling_ab.txt
for dbg in 2 6; do
  echo "== OMGSR_HALO_ROLLING=1 OMGSR_HALO_DEBUG=$dbg (2 no epilogue, 4 no MMAs)"
  OMGSR_HALO_ROLLING=1 OMGSR_HALO_DEBUG=$dbg timeout 200 python scripts/halo_ab.py 2>&1 | head -6
done > gpurun_out/r2p_halo_rolling_ablation.txt 2>&1; cat gpurun_out/r2p_halo_rolling_ablation.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2p_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2p_tests.log | cut -c1-200
timeout 300 python scripts/s512_stages.py > gpurun_out/r2p_s512_stages.txt 2>&1; tail -1 gpurun_out/r2p_s512_stages.txt
OMGSR_HALO_ROLLING=0 timeout 300 python scripts/s512_stages.py > gpurun_out/r2p_s512_stages_rr0.txt 2>&1; tail -1 gpurun_out/r2p_s512_stages_rr0.txt
