#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_encoder_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/clip_tests2.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/clip_tests2.log | cut -c1-200
timeout 300 python infer/infer_omgsr_s.py --sd_path "random://sd21?seed=0" --lora_path "random://lora?seed=1" --output_dir gpurun_out/tmp_drv --synthetic 4 --batch_size 4 2>&1 | tail -4
