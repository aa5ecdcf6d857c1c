#!/bin/bash
# one gpurun call: every kernel parity group in its own process (a trapped kernel poisons its CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "gemm" "conv" "materialised" "flash" "groupnorm or layernorm or thin or latent or rmsnorm or tile_blend or colorfix"; do
  tag=$(echo "$grp" | cut -d' ' -f1)
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k "$grp" -p no:cacheprovider > gpurun_out/k_$tag.log 2>&1
  echo "== $tag exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/k_$tag.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
