#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4 6; do
  echo "== OMGSR_HALO_DEBUG=$dbg (1 no stores, 2 no epilogue, 4 no MMAs)"
  OMGSR_HALO_DEBUG=$dbg timeout 120 python scripts/halo_ab.py 2>&1 | grep -v upsample
done > gpurun_out/r2d_ablation.txt 2>&1
cat gpurun_out/r2d_ablation.txt
