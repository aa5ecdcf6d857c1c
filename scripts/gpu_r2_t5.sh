#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_encoder_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/t5_tests.log 2>&1
echo "pytest rc=$?"; grep -n "clip text\|t5 encoder\|passed\|failed\|Error\|error" gpurun_out/t5_tests.log | head -16
