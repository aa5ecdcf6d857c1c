#!/bin/bash
# packed fp32 (FFMA2 / FADD2) softmax with sum-triggered rebase (flash_attn3 / 4) + FFMA2 halo transform: tests, A/B, stages
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2o_tests.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2o_tests.log | cut -c1-200
timeout 300 python scripts/attn_ab.py > gpurun_out/r2o_attn_ab.txt 2>&1; cat gpurun_out/r2o_attn_ab.txt
timeout 300 python scripts/halo_ab.py > gpurun_out/r2o_halo.txt 2>&1; cat gpurun_out/r2o_halo.txt
timeout 300 python scripts/s512_stages.py > gpurun_out/r2o_s512_stages.txt 2>&1; tail -2 gpurun_out/r2o_s512_stages.txt
timeout 300 python scripts/f1024.py 4 4 > gpurun_out/r2o_f1024_stages.txt 2>&1; tail -1 gpurun_out/r2o_f1024_stages.txt
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2o_tests2.log 2>&1
echo "pytest2 rc=$?"; tail -6 gpurun_out/r2o_tests2.log | cut -c1-200
