"""Oracle restatement of the un-vendored nets the reference calls (TEST INFRASTRUCTURE ONLY).

The reference imports these from diffusers==0.34.0 (requirements.txt:4), which is neither under
/root/reference nor installable offline; the forward passes below restate diffusers' published
architecture as plain fp32-capable PyTorch functions over a diffusers-keyed
state dict. Sub-module names are corroborated by the attribute names the reference itself dereferences
(infer/vaehook.py:137-171,248-359) and by its LoRA target lists (train/train_omgsr_s.py:61-72,
train/train_omgsr_f.py:155-169). PARITY UNPINNED: no reference fixture exists for these nets.
"""
import math

import torch
import torch.nn.functional as F


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


GN_HOOK = None  # set by oracle/vaehook.py to share GroupNorm statistics across tiles


def _gn(sd, name, x, eps, groups=32):
    if GN_HOOK is not None:
        return GN_HOOK(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def _ln(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def resnet_block(sd, name, x, eps, temb=None):
    """diffusers ResnetBlock2D (mirrored task-by-task in infer/vaehook.py:248-276)."""
    h = _conv(sd, name + ".conv1", F.silu(_gn(sd, name + ".norm1", x, eps)))
    if temb is not None:
        h = h + _lin(sd, name + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = _conv(sd, name + ".conv2", F.silu(_gn(sd, name + ".norm2", h, eps)))
    if name + ".conv_shortcut.weight" in sd:
        x = _conv(sd, name + ".conv_shortcut", x, padding=0)
    return x + h


def vae_attention(sd, name, x, eps):
    """single-head attention of the VAE mid block (infer/vaehook.py:137-171 restates the same math)."""
    b, c, hh, ww = x.shape
    h = _gn(sd, name + ".group_norm", x, eps).reshape(b, c, hh * ww).transpose(1, 2)
    q, k, v = _lin(sd, name + ".to_q", h), _lin(sd, name + ".to_k", h), _lin(sd, name + ".to_v", h)
    a = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1) @ v  # materialised, as infer/vaehook.py:160-161
    a = _lin(sd, name + ".to_out.0", a)
    return x + a.transpose(1, 2).reshape(b, c, hh, ww)


def vae_encode_moments(sd, cfg, x):
    """AutoencoderKL.encode up to the posterior parameters (mean | logvar)."""
    eps = cfg.eps
    h = _conv(sd, "encoder.conv_in", x)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, eps)
        if i < n - 1:
            h = F.pad(h, (0, 1, 0, 1))  # diffusers Downsample2D with padding=0: asymmetric pad
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = resnet_block(sd, "encoder.mid_block.resnets.0", h, eps)
    h = vae_attention(sd, "encoder.mid_block.attentions.0", h, eps)
    h = resnet_block(sd, "encoder.mid_block.resnets.1", h, eps)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, eps)))
    if cfg.use_quant_conv:
        h = _conv(sd, "quant_conv", h, padding=0)
    return h


def posterior_sample(moments, eps_noise):
    """DiagonalGaussianDistribution.sample(): mean + exp(0.5 * clamp(logvar, -30, 20)) * eps."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps_noise


def vae_decode(sd, cfg, z):
    eps = cfg.eps
    if cfg.use_quant_conv:
        z = _conv(sd, "post_quant_conv", z, padding=0)
    h = _conv(sd, "decoder.conv_in", z)
    h = resnet_block(sd, "decoder.mid_block.resnets.0", h, eps)
    h = vae_attention(sd, "decoder.mid_block.attentions.0", h, eps)
    h = resnet_block(sd, "decoder.mid_block.resnets.1", h, eps)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, eps)
        if i < n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, eps)))


# ------------------------------------------------------------------------------------------------
# UNet2DConditionModel (SD-2.1-base shape)
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim, flip_sin_to_cos=True, max_period=10000.0):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t.float()[:, None] * freqs[None]
    emb = torch.cat([torch.sin(args), torch.cos(args)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def _mha(q, k, v, heads):
    b, m, c = q.shape
    d = c // heads
    q = q.reshape(b, m, heads, d).transpose(1, 2)
    k = k.reshape(b, -1, heads, d).transpose(1, 2)
    v = v.reshape(b, -1, heads, d).transpose(1, 2)
    return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, m, c)


def transformer2d(sd, name, x, ctx, heads):
    b, c, hh, ww = x.shape
    h = _gn(sd, name + ".norm", x, 1e-6).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    h = _lin(sd, name + ".proj_in", h)
    tb = name + ".transformer_blocks.0"
    n = _ln(sd, tb + ".norm1", h)
    a = _mha(_lin(sd, tb + ".attn1.to_q", n), _lin(sd, tb + ".attn1.to_k", n), _lin(sd, tb + ".attn1.to_v", n), heads)
    h = h + _lin(sd, tb + ".attn1.to_out.0", a)
    n = _ln(sd, tb + ".norm2", h)
    a = _mha(_lin(sd, tb + ".attn2.to_q", n), _lin(sd, tb + ".attn2.to_k", ctx), _lin(sd, tb + ".attn2.to_v", ctx), heads)
    h = h + _lin(sd, tb + ".attn2.to_out.0", a)
    n = _ln(sd, tb + ".norm3", h)
    val, gate = _lin(sd, tb + ".ff.net.0.proj", n).chunk(2, dim=-1)
    h = h + _lin(sd, tb + ".ff.net.2", val * F.gelu(gate))
    h = _lin(sd, name + ".proj_out", h)
    return x + h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)


def unet_forward(sd, cfg, x, timestep, ctx):
    b = x.shape[0]
    ch = cfg.block_out_channels
    eps = cfg.norm_eps
    t = torch.full((b,), float(timestep), dtype=torch.float32)
    temb = timestep_embedding(t, ch[0]).to(x.dtype)
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", temb)))
    if ctx.shape[0] != b:
        ctx = ctx.expand(b, -1, -1)
    h = _conv(sd, "conv_in", x)
    skips = [h]
    n = len(ch)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, eps, temb)
            if cfg.down_has_attn[i]:
                h = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", h, ctx, cfg.num_heads[i])
            skips.append(h)
        if i < n - 1:
            h = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1)
            skips.append(h)
    h = resnet_block(sd, "mid_block.resnets.0", h, eps, temb)
    h = transformer2d(sd, "mid_block.attentions.0", h, ctx, cfg.num_heads[-1])
    h = resnet_block(sd, "mid_block.resnets.1", h, eps, temb)
    rev_attn = list(reversed(cfg.down_has_attn))
    rev_heads = list(reversed(cfg.num_heads))
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, eps, temb)
            if rev_attn[i]:
                h = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", h, ctx, rev_heads[i])
        if i < n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h)
    return _conv(sd, "conv_out", F.silu(_gn(sd, "conv_norm_out", h, eps)))
