#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_text_encoder_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/clip_tests.log 2>&1
echo "pytest rc=$?"; grep -n "clip text\|passed\|failed\|Error\|error" gpurun_out/clip_tests.log | head -12
timeout 600 python scripts/clip_text_time.py > gpurun_out/clip_text_time.txt 2>&1; tail -4 gpurun_out/clip_text_time.txt
