#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pipeline.log 2>&1
echo "== pipeline exit $?" > gpurun_out/summary.txt
grep -E "relerr|psnr|passed|failed|Error|error" gpurun_out/pipeline.log | tail -40 >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "groupnorm or tile_blend or conv" > gpurun_out/k2.log 2>&1
echo "== kernels exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/k2.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
