#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "flash or attention" > gpurun_out/v5_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/v5_tests.log | cut -c1-200
OMGSR_ATTN_VARIANT=5 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "flash_attention_fused or flash_attention[" > gpurun_out/v5_tests_b.log 2>&1
echo "pytest(variant 5 default) rc=$?"; tail -2 gpurun_out/v5_tests_b.log | cut -c1-200
timeout 300 python scripts/attn_ab.py > gpurun_out/v5_attn_ab.txt 2>&1; cat gpurun_out/v5_attn_ab.txt | cut -c1-230
for v in 3 5; do OMGSR_ATTN_VARIANT=$v timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-200 | sed "s/^/variant $v F: /"; done
