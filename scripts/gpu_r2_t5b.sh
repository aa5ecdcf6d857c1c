#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_encoder_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t5_tests2.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/t5_tests2.log | cut -c1-200
timeout 600 python infer/infer_omgsr_f.py --flux_path "random://tiny?seed=0" --lora_path "random://lora?seed=1" --output_dir gpurun_out/tmp_drv_f --synthetic 2 --process_size 256 2>&1 | tail -3
timeout 900 python infer/infer_omgsr_f.py --flux_path "random://flux_dev?seed=0" --lora_path "random://lora?seed=1&rank=64" --output_dir gpurun_out/tmp_drv_f2 --synthetic 1 2>&1 | tail -3
python - <<'P'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, omgsr_synth as syn
from omgsr_b200.engine_t5 import build_t5_encoder
t5 = build_t5_encoder("random://t5_xxl?seed=3", "cuda")
ids = syn.empty_prompt_ids_t5(t5.cfg, 512).cuda()
for _ in range(2): t5(ids)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(5): t5(ids)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
c = t5.cfg
fl = 2.0 * 512 * c.num_layers * (4 * c.d_model * c.num_heads * c.d_kv + 3 * c.d_model * c.d_ff) + 4.0 * c.num_layers * c.num_heads * 512 * 512 * c.d_kv
print(f"T5-XXL encoder, 512 tokens, batch 1: {ms:.2f} ms ({fl / ms / 1e9:.0f} TFLOP/s), {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB peak")
P
