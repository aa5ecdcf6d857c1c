"""Text-conditioning step: one CLIP text encode of the empty prompt through the sm_100a engine (random-init weights of
the SD-2.1 / CLIP-L shapes), CUDA-event timed, next to transformers' CLIPTextModel in bf16 on the same GPU and fp32 on the
host cores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import omgsr_synth as syn
from omgsr_b200.engine_clip import B200CLIPText

def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

import transformers
for name, B in (("SD21_CLIP_H", 1), ("SD21_CLIP_H", 16), ("FLUX_CLIP_L", 1)):
    cfg = getattr(syn, name)
    sd = syn.random_state_dict(syn.clip_text_shapes(cfg), seed=11)
    ids = syn.empty_prompt_ids(cfg).repeat(B, 1)
    eng = B200CLIPText(sd, cfg, "cuda")
    ids_dev = ids.cuda()
    ms = timeit(lambda: eng(ids_dev))
    hf_cfg = transformers.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, max_position_embeddings=cfg.max_positions,
        hidden_act=cfg.hidden_act, eos_token_id=cfg.eos_token_id, pad_token_id=cfg.pad_token_id, bos_token_id=0)
    m = transformers.CLIPTextModel(hf_cfg).eval(); m.load_state_dict(sd, strict=False)
    with torch.no_grad():
        t0 = time.time(); m(input_ids=ids); cpu_s = time.time() - t0
        mg = m.to("cuda", torch.bfloat16)
        ms_hf = timeit(lambda: mg(input_ids=ids_dev))
    flops = 2.0 * B * cfg.max_positions * cfg.num_layers * (4 * cfg.hidden_size ** 2 + 2 * cfg.hidden_size * cfg.intermediate_size)
    print(f"{name} batch {B}: engine {ms:.3f} ms ({flops / ms / 1e9:.1f} TFLOP/s, {10 * cfg.num_layers + 2} launches), "
          f"transformers bf16 on the GPU {ms_hf:.3f} ms, transformers fp32 on the host {cpu_s * 1e3:.0f} ms", flush=True)
    del eng, m, mg
