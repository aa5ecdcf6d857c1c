#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for f in "" 1; do
  OMGSR_GN_FUSE4=$f timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
done
OMGSR_GN_FUSE4=1 timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
