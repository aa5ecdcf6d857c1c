#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
OMGSR_NO_GN_IN_FUSION=1 timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
