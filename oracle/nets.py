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
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
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
    t = torch.full((b,), float(timestep), dtype=torch.float32, device=x.device)
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


# ------------------------------------------------------------------------------------------------
# FluxTransformer2DModel (FLUX.1-dev shape) — diffusers 0.34.0 semantics, SURVEY.md Appendix C.
# Cross-checked against torchtitan's BFL FLUX (experiments/flux/model/{layers,math}.py) for the block math;
# parameter names are diffusers' (train/train_omgsr_f.py:155-169 lists the same module names).
# ------------------------------------------------------------------------------------------------
def flux_rope(ids, axes_dims, theta=10000.0):
    """FluxPosEmbed: per axis cos/sin of pos * theta^(-2i/d), each repeated twice -> [S, sum(axes_dims)]."""
    cos_out, sin_out = [], []
    pos = ids.double()
    for i, d in enumerate(axes_dims):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
        ang = pos[:, i:i + 1] * freqs[None]
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rope(x, cos, sin):
    """x [B,H,S,D]; pairs (x0,x1) -> (x0 cos - x1 sin, x1 cos + x0 sin), computed in fp32 and cast back."""
    xf = x.float()
    x2 = xf.reshape(*xf.shape[:-1], -1, 2)
    rot = torch.stack([-x2[..., 1], x2[..., 0]], dim=-1).flatten(3)
    return (xf * cos[None, None] + rot * sin[None, None]).to(x.dtype)


def _rms(x, w, eps=1e-6):
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype) * w


def _sincos256(v):
    half = 128
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=v.device) / half)
    args = v.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)  # flip_sin_to_cos=True


def _mlp_embed(sd, name, x):
    return _lin(sd, name + ".linear_2", F.silu(_lin(sd, name + ".linear_1", x)))


def _ln_plain(x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def _heads(x, h):
    b, s, c = x.shape
    return x.reshape(b, s, h, c // h).transpose(1, 2)  # [B,H,S,D]


def flux_forward(sd, cfg, hidden_states, timestep, guidance, pooled, encoder_hidden_states, txt_ids, img_ids):
    """hidden_states [B,T,64] packed latents; timestep [1] or [B] in sigma units (0..1); guidance [B];
    pooled [B,768]; encoder_hidden_states [B,512,4096]; txt_ids [512,3]; img_ids [T,3] -> [B,T,64]."""
    H = cfg.num_heads
    dt = hidden_states.dtype
    B = hidden_states.shape[0]
    dev = hidden_states.device
    x = _lin(sd, "x_embedder", hidden_states)
    # diffusers casts timestep / guidance to the hidden dtype BEFORE the x1000 (bf16 weights => the net is conditioned
    # on bf16(0.5051) * 1000 = 504, fp32 weights => 505.11); the fp32 evaluation is the parity target
    t = timestep.to(device=dev, dtype=dt).expand(B) * 1000
    temb = _mlp_embed(sd, "time_text_embed.timestep_embedder", _sincos256(t).to(dt))
    if cfg.guidance_embeds:
        g = guidance.to(device=dev, dtype=dt) * 1000
        temb = temb + _mlp_embed(sd, "time_text_embed.guidance_embedder", _sincos256(g).to(dt))
    temb = temb + _mlp_embed(sd, "time_text_embed.text_embedder", pooled)
    c = _lin(sd, "context_embedder", encoder_hidden_states)
    cos, sin = flux_rope(torch.cat([txt_ids, img_ids], dim=0).float().to(dev), cfg.axes_dims_rope)
    st = F.silu(temb)
    n_txt = c.shape[1]
    for i in range(cfg.num_layers):
        b = f"transformer_blocks.{i}"
        sh1, sc1, g1, sh2, sc2, g2 = _lin(sd, b + ".norm1.linear", st).chunk(6, dim=1)
        csh1, csc1, cg1, csh2, csc2, cg2 = _lin(sd, b + ".norm1_context.linear", st).chunk(6, dim=1)
        xn = _ln_plain(x) * (1 + sc1[:, None]) + sh1[:, None]
        cn = _ln_plain(c) * (1 + csc1[:, None]) + csh1[:, None]
        q = _rms(_heads(_lin(sd, b + ".attn.to_q", xn), H), sd[b + ".attn.norm_q.weight"])
        k = _rms(_heads(_lin(sd, b + ".attn.to_k", xn), H), sd[b + ".attn.norm_k.weight"])
        v = _heads(_lin(sd, b + ".attn.to_v", xn), H)
        cq = _rms(_heads(_lin(sd, b + ".attn.add_q_proj", cn), H), sd[b + ".attn.norm_added_q.weight"])
        ck = _rms(_heads(_lin(sd, b + ".attn.add_k_proj", cn), H), sd[b + ".attn.norm_added_k.weight"])
        cv = _heads(_lin(sd, b + ".attn.add_v_proj", cn), H)
        q, k, v = torch.cat([cq, q], 2), torch.cat([ck, k], 2), torch.cat([cv, v], 2)  # [text ; image]
        a = F.scaled_dot_product_attention(apply_rope(q, cos, sin), apply_rope(k, cos, sin), v)
        a = a.transpose(1, 2).reshape(B, -1, H * cfg.head_dim)
        ca, xa = a[:, :n_txt], a[:, n_txt:]
        x = x + g1[:, None] * _lin(sd, b + ".attn.to_out.0", xa)
        xn = _ln_plain(x) * (1 + sc2[:, None]) + sh2[:, None]
        x = x + g2[:, None] * _lin(sd, b + ".ff.net.2", F.gelu(_lin(sd, b + ".ff.net.0.proj", xn), approximate="tanh"))
        c = c + cg1[:, None] * _lin(sd, b + ".attn.to_add_out", ca)
        cn = _ln_plain(c) * (1 + csc2[:, None]) + csh2[:, None]
        c = c + cg2[:, None] * _lin(sd, b + ".ff_context.net.2",
                                    F.gelu(_lin(sd, b + ".ff_context.net.0.proj", cn), approximate="tanh"))
    h = torch.cat([c, x], dim=1)
    for i in range(cfg.num_single_layers):
        b = f"single_transformer_blocks.{i}"
        sh, sc, gt = _lin(sd, b + ".norm.linear", st).chunk(3, dim=1)
        hn = _ln_plain(h) * (1 + sc[:, None]) + sh[:, None]
        m = F.gelu(_lin(sd, b + ".proj_mlp", hn), approximate="tanh")
        q = _rms(_heads(_lin(sd, b + ".attn.to_q", hn), H), sd[b + ".attn.norm_q.weight"])
        k = _rms(_heads(_lin(sd, b + ".attn.to_k", hn), H), sd[b + ".attn.norm_k.weight"])
        v = _heads(_lin(sd, b + ".attn.to_v", hn), H)
        a = F.scaled_dot_product_attention(apply_rope(q, cos, sin), apply_rope(k, cos, sin), v)
        a = a.transpose(1, 2).reshape(B, -1, H * cfg.head_dim)
        h = h + gt[:, None] * _lin(sd, b + ".proj_out", torch.cat([a, m], dim=2))
    x = h[:, n_txt:]
    scale, shift = _lin(sd, "norm_out.linear", st).chunk(2, dim=1)  # diffusers AdaLayerNormContinuous: scale first
    x = _ln_plain(x) * (1 + scale[:, None]) + shift[:, None]
    return _lin(sd, "proj_out", x)
