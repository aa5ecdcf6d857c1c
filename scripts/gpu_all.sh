#!/bin/bash
# full GPU regression + timing probes + bench in one call
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "== kernels exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/kernels.log >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pipeline.log 2>&1
echo "== pipeline exit $?" >> gpurun_out/summary.txt
grep -E "relerr|psnr|passed|failed|Error|error" gpurun_out/pipeline.log | tail -30 >> gpurun_out/summary.txt
timeout 600 python scripts/conv_time.py > gpurun_out/conv_time.log 2>&1
cat gpurun_out/conv_time.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt; tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
