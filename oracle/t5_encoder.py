"""Oracle restatement of the T5 v1.1 encoder stack (TEST INFRASTRUCTURE ONLY: only tests/, smoke() and bench.py's
baseline legs may import this).

The reference gets OMGSR-F's `prompt_embeds` [1, 512, 4096] from `FluxPipeline.encode_prompt` (infer/infer_omgsr_f.py:
36-43), which runs `T5EncoderModel(google/t5-v1_1-xxl)(ids)[0]` once per process. The arithmetic lives in transformers
(requirements.txt:8 pins 4.55.2; 5.5 installed here), src/transformers/models/t5/modeling_t5.py:

  T5LayerNorm          x * rsqrt(mean(x^2) + eps) * w            (no mean subtraction, no bias, statistics in fp32)
  T5Attention          q k^T WITHOUT the 1/sqrt(d) factor + position_bias, softmax in fp32, no projection biases
  position_bias        relative_attention_bias[bucket(j - i)] of the FIRST block, shared by all blocks; bidirectional
                       buckets: half for the sign, exact up to max_exact = buckets / 4, logarithmic up to max_distance
  T5DenseGatedActDense wo(gelu_new(wi_0 x) * wi_1 x),  gelu_new = 0.5 x (1 + tanh(sqrt(2/pi)(x + 0.044715 x^3)))
  T5Block / T5Stack    x + attn(ln(x)), x + ff(ln(x)); final_layer_norm

(encode_prompt passes no attention mask to the T5 encoder: pad tokens are attended like any other.)
PINNED: tests/test_oracle_cpu.py::test_t5_encoder_oracle_matches_transformers runs transformers' own T5EncoderModel on
the same random weights (tiny and small configs, CPU).
"""
import math

import torch
import torch.nn.functional as F


def relative_position_bucket(rel, num_buckets=32, max_distance=128):
    """modeling_t5.py T5Attention._relative_position_bucket, bidirectional=True. rel = key position - query position."""
    nb = num_buckets // 2
    out = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) *
                         (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(is_small, rel, large)


def position_bias(table, T, num_buckets, max_distance):
    """table: relative_attention_bias.weight [buckets, heads] -> [heads, T, T] fp32."""
    ctx = torch.arange(T, device=table.device)[:, None]
    mem = torch.arange(T, device=table.device)[None, :]
    b = relative_position_bucket(mem - ctx, num_buckets, max_distance)
    return table.float()[b].permute(2, 0, 1).contiguous()


def _rms(x, w, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def _gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def t5_encoder_forward(sd, cfg, input_ids):
    """sd: transformers' T5EncoderModel key names (evaluated in fp32); input_ids int64 [B, T] -> [B, T, d_model]."""
    f = lambda k: sd[k].to(torch.float32)
    B, T = input_ids.shape
    H, D = cfg.num_heads, cfg.d_kv
    x = f("shared.weight")[input_ids]
    bias = position_bias(f("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"), T, cfg.num_buckets,
                         cfg.max_distance)
    for i in range(cfg.num_layers):
        b = f"encoder.block.{i}.layer"
        h = _rms(x, f(b + ".0.layer_norm.weight"), cfg.layer_norm_eps)
        q = F.linear(h, f(b + ".0.SelfAttention.q.weight")).view(B, T, H, D).transpose(1, 2)
        k = F.linear(h, f(b + ".0.SelfAttention.k.weight")).view(B, T, H, D).transpose(1, 2)
        v = F.linear(h, f(b + ".0.SelfAttention.v.weight")).view(B, T, H, D).transpose(1, 2)
        p = torch.softmax(q @ k.transpose(-1, -2) + bias[None], dim=-1)
        a = (p @ v).transpose(1, 2).reshape(B, T, H * D)
        x = x + F.linear(a, f(b + ".0.SelfAttention.o.weight"))
        h = _rms(x, f(b + ".1.layer_norm.weight"), cfg.layer_norm_eps)
        g = _gelu_new(F.linear(h, f(b + ".1.DenseReluDense.wi_0.weight"))) * F.linear(h, f(b + ".1.DenseReluDense.wi_1.weight"))
        x = x + F.linear(g, f(b + ".1.DenseReluDense.wo.weight"))
    return _rms(x, f("encoder.final_layer_norm.weight"), cfg.layer_norm_eps)
