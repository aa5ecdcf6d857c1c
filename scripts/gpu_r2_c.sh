#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv or halo or gn or group" > gpurun_out/r2c_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c_tests.log
for dbg in 0 2 6; do
  echo "== OMGSR_HALO_DEBUG=$dbg (1 no stores, 2 no epilogue, 4 no MMAs)"
  OMGSR_HALO_DEBUG=$dbg timeout 120 python scripts/halo_ab.py 2>&1
done > gpurun_out/r2c_halo_new.txt 2>&1
cat gpurun_out/r2c_halo_new.txt
