#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
COMMON="--sd_path random://sd21?seed=0 --lora_path random://lora?seed=1 --synthetic 64 --process_size 512 --align_method wavelet"
for extra in "--host_preprocess" "" "--cuda_graph" "--batch_size 16"; do
  rm -rf /tmp/out_drv
  echo "== infer_omgsr_s.py $extra" >> gpurun_out/summary.txt
  timeout 600 python infer/infer_omgsr_s.py $COMMON --output_dir /tmp/out_drv $extra 2>&1 | grep -E "Average|Driver wall" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
