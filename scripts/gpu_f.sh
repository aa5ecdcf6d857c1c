#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -s -k "rmsnorm or gemm_batched or layernorm or thin" > gpurun_out/k3.log 2>&1
echo "== kernels exit $?" > gpurun_out/summary.txt; grep -E "FAIL|passed|failed" gpurun_out/k3.log | tail -8 >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pipeline.log 2>&1
echo "== pipeline exit $?" >> gpurun_out/summary.txt
grep -E "relerr|psnr|passed|failed|Error|error" gpurun_out/pipeline.log | tail -40 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
