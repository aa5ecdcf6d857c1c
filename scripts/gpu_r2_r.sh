#!/bin/bash
# warp-uniform MMA issue in every tcgen05 kernel: parity (all kernel tests), then kernel timings and stages
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2r_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2r_tests.log | cut -c1-200
for mode in 0 1; do
  echo "== OMGSR_HALO_ROLLING=$mode"
  OMGSR_HALO_ROLLING=$mode timeout 200 python scripts/halo_ab.py 2>&1 | head -7
done > gpurun_out/r2r_halo_ab.txt 2>&1; cat gpurun_out/r2r_halo_ab.txt
timeout 300 python scripts/attn_ab.py > gpurun_out/r2r_attn_ab.txt 2>&1; cat gpurun_out/r2r_attn_ab.txt
timeout 300 python scripts/conv_time.py > gpurun_out/r2r_conv_time.txt 2>&1; tail -30 gpurun_out/r2r_conv_time.txt
timeout 300 python scripts/s512_stages.py > gpurun_out/r2r_s512_stages.txt 2>&1; tail -1 gpurun_out/r2r_s512_stages.txt
timeout 300 python scripts/f1024.py 4 4 > gpurun_out/r2r_f1024_stages.txt 2>&1; tail -1 gpurun_out/r2r_f1024_stages.txt
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2r_tests2.log 2>&1
echo "pytest2 rc=$?"; tail -4 gpurun_out/r2r_tests2.log | cut -c1-200
