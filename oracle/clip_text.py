"""Oracle restatement of the CLIP text tower (TEST INFRASTRUCTURE ONLY: only tests/, smoke() and bench.py's baseline
legs may import this).

The reference computes its prompt embedding once per process, before the super-resolution path:

    infer/infer_omgsr_s.py:19-45   CLIPTextModel.from_pretrained(sd_path, subfolder="text_encoder")(ids)[0]
                                    -> prompt_embeds [1, 77, 1024] (the LAST HIDDEN STATE, after final_layer_norm)
    infer/infer_omgsr_f.py:36-43   FluxPipeline.encode_prompt: CLIPTextModel(ids).pooler_output -> [1, 768]
                                    (+ the T5-XXL encoder for prompt_embeds, not restated here)

The arithmetic lives in transformers (requirements.txt:8 pins 4.55.2; 5.5 is installed in this image), file
src/transformers/models/clip/modeling_clip.py: CLIPTextEmbeddings (token + position embedding), CLIPEncoderLayer
(pre-LN: x + attn(ln1(x)), x + mlp(ln2(x))), CLIPAttention (scale = head_dim ** -0.5, causal mask), CLIPMLP
(fc1 -> act -> fc2, act = "gelu" (erf) for the SD-2.1 tower, "quick_gelu" = x * sigmoid(1.702 x) for OpenAI CLIP-L),
CLIPTextTransformer (final_layer_norm; pooled = row of the EOS token: first position whose id == eos_token_id, or
argmax(ids) for the legacy eos_token_id == 2 configs).

PINNED: tests/test_oracle_cpu.py::test_clip_text_oracle_matches_transformers runs transformers' own CLIPTextModel
(random-init, same weights) next to this file on CPU - tiny configs and the full SD-2.1 / CLIP-L sizes.
"""
import torch
import torch.nn.functional as F


def _act(x, kind):
    if kind == "gelu":
        return F.gelu(x)
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(kind)


def clip_text_forward(sd, cfg, input_ids):
    """sd: state dict with transformers' `text_model.*` names (any float dtype; evaluated in fp32);
    input_ids: int64 [B, T]. Returns (last_hidden_state [B, T, C], pooler_output [B, C])."""
    f = lambda k: sd[k].to(torch.float32)
    B, T = input_ids.shape
    C, H = cfg.hidden_size, cfg.num_heads
    D = C // H
    x = f("text_model.embeddings.token_embedding.weight")[input_ids] + \
        f("text_model.embeddings.position_embedding.weight")[:T][None]
    mask = torch.full((T, T), float("-inf"), device=x.device).triu(1)
    for i in range(cfg.num_layers):
        b = f"text_model.encoder.layers.{i}"
        h = F.layer_norm(x, (C,), f(b + ".layer_norm1.weight"), f(b + ".layer_norm1.bias"), cfg.layer_norm_eps)
        q = F.linear(h, f(b + ".self_attn.q_proj.weight"), f(b + ".self_attn.q_proj.bias")).view(B, T, H, D).transpose(1, 2)
        k = F.linear(h, f(b + ".self_attn.k_proj.weight"), f(b + ".self_attn.k_proj.bias")).view(B, T, H, D).transpose(1, 2)
        v = F.linear(h, f(b + ".self_attn.v_proj.weight"), f(b + ".self_attn.v_proj.bias")).view(B, T, H, D).transpose(1, 2)
        p = torch.softmax(q @ k.transpose(-1, -2) * (D ** -0.5) + mask, dim=-1)
        a = (p @ v).transpose(1, 2).reshape(B, T, C)
        x = x + F.linear(a, f(b + ".self_attn.out_proj.weight"), f(b + ".self_attn.out_proj.bias"))
        h = F.layer_norm(x, (C,), f(b + ".layer_norm2.weight"), f(b + ".layer_norm2.bias"), cfg.layer_norm_eps)
        h = _act(F.linear(h, f(b + ".mlp.fc1.weight"), f(b + ".mlp.fc1.bias")), cfg.hidden_act)
        x = x + F.linear(h, f(b + ".mlp.fc2.weight"), f(b + ".mlp.fc2.bias"))
    x = F.layer_norm(x, (C,), f("text_model.final_layer_norm.weight"), f("text_model.final_layer_norm.bias"),
                     cfg.layer_norm_eps)
    pooled = x[torch.arange(B, device=x.device), eos_positions(input_ids, cfg.eos_token_id)]
    return x, pooled


def eos_positions(input_ids, eos_token_id):
    """modeling_clip.py CLIPTextTransformer.forward: legacy configs (eos_token_id == 2) take argmax(ids) - the EOS token
    has the largest id of the vocabulary; otherwise the FIRST position holding eos_token_id."""
    if eos_token_id == 2:
        return input_ids.to(torch.int).argmax(dim=-1)
    return (input_ids.to(torch.int) == eos_token_id).int().argmax(dim=-1)
