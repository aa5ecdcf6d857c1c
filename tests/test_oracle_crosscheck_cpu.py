"""The oracle's FLUX transformer and FLUX autoencoder (oracle/nets.py, written from recall of diffusers 0.34.0, which is
not installable here) cross-checked against an INDEPENDENT implementation that is present in this image: torchtitan's
port of Black Forest Labs' reference code (torchtitan/experiments/flux/model/{model,layers,math,autoencoder}.py).
The same seeded weights are re-packed from diffusers' parameter layout into BFL's (fused qkv, linear1 = q|k|v|mlp,
LDM-style autoencoder names, (shift, scale) order of the final modulation) and the two forward passes must agree to
fp32 rounding. This does not make the reference itself importable, but it removes "one person's recall" as the only
source of the two largest nets of the OMGSR-F path (SURVEY.md section 8c lists this cross-check)."""
import pytest
import torch

tt = pytest.importorskip("torchtitan.experiments.flux.model.model")


def _copy(dst, src):
    assert dst.shape == src.shape, (tuple(dst.shape), tuple(src.shape))
    dst.data.copy_(src)


def _lin(mod, sd, name):
    _copy(mod.weight, sd[name + ".weight"])
    _copy(mod.bias, sd[name + ".bias"])


def test_flux_transformer_matches_bfl_reference_port():
    from torchtitan.experiments.flux.model.args import FluxModelArgs
    from omgsr_b200 import weights as W
    from oracle import nets
    cfg = W.TINY_FLUX
    D = cfg.num_heads * cfg.head_dim
    sd = W.random_state_dict(W.flux_shapes(cfg), 11)
    # BFL's port has no guidance embedder (schnell layout): make the oracle's contribute exactly zero
    sd["time_text_embed.guidance_embedder.linear_2.weight"].zero_()
    sd["time_text_embed.guidance_embedder.linear_2.bias"].zero_()
    m = tt.FluxModel(FluxModelArgs(in_channels=cfg.in_channels, out_channels=cfg.in_channels,
                                   vec_in_dim=cfg.pooled_projection_dim, context_in_dim=cfg.joint_attention_dim,
                                   hidden_size=D, num_heads=cfg.num_heads, depth=cfg.num_layers,
                                   depth_single_blocks=cfg.num_single_layers, axes_dim=tuple(cfg.axes_dims_rope)))
    _lin(m.img_in, sd, "x_embedder")
    _lin(m.txt_in, sd, "context_embedder")
    _lin(m.time_in.in_layer, sd, "time_text_embed.timestep_embedder.linear_1")
    _lin(m.time_in.out_layer, sd, "time_text_embed.timestep_embedder.linear_2")
    _lin(m.vector_in.in_layer, sd, "time_text_embed.text_embedder.linear_1")
    _lin(m.vector_in.out_layer, sd, "time_text_embed.text_embedder.linear_2")
    cat = lambda names, suffix: torch.cat([sd[n + suffix] for n in names], 0)
    for i, blk in enumerate(m.double_blocks):
        b = f"transformer_blocks.{i}"
        _lin(blk.img_mod.lin, sd, b + ".norm1.linear")
        _lin(blk.txt_mod.lin, sd, b + ".norm1_context.linear")
        for attn, q, k, v, nq, nk, out in ((blk.img_attn, "to_q", "to_k", "to_v", "norm_q", "norm_k", "to_out.0"),
                                           (blk.txt_attn, "add_q_proj", "add_k_proj", "add_v_proj", "norm_added_q",
                                            "norm_added_k", "to_add_out")):
            names = [f"{b}.attn.{q}", f"{b}.attn.{k}", f"{b}.attn.{v}"]
            _copy(attn.qkv.weight, cat(names, ".weight"))
            _copy(attn.qkv.bias, cat(names, ".bias"))
            _copy(attn.norm.query_norm.weight, sd[f"{b}.attn.{nq}.weight"])
            _copy(attn.norm.key_norm.weight, sd[f"{b}.attn.{nk}.weight"])
            _lin(attn.proj, sd, f"{b}.attn.{out}")
        _lin(blk.img_mlp[0], sd, b + ".ff.net.0.proj")
        _lin(blk.img_mlp[2], sd, b + ".ff.net.2")
        _lin(blk.txt_mlp[0], sd, b + ".ff_context.net.0.proj")
        _lin(blk.txt_mlp[2], sd, b + ".ff_context.net.2")
    for i, blk in enumerate(m.single_blocks):
        b = f"single_transformer_blocks.{i}"
        _lin(blk.modulation.lin, sd, b + ".norm.linear")
        names = [b + ".attn.to_q", b + ".attn.to_k", b + ".attn.to_v", b + ".proj_mlp"]
        _copy(blk.linear1.weight, cat(names, ".weight"))
        _copy(blk.linear1.bias, cat(names, ".bias"))
        _lin(blk.linear2, sd, b + ".proj_out")
        _copy(blk.norm.query_norm.weight, sd[b + ".attn.norm_q.weight"])
        _copy(blk.norm.key_norm.weight, sd[b + ".attn.norm_k.weight"])
    # final modulation: diffusers AdaLayerNormContinuous emits (scale, shift), BFL's LastLayer (shift, scale)
    w, bias = sd["norm_out.linear.weight"], sd["norm_out.linear.bias"]
    _copy(m.final_layer.adaLN_modulation[1].weight, torch.cat([w[D:], w[:D]], 0))
    _copy(m.final_layer.adaLN_modulation[1].bias, torch.cat([bias[D:], bias[:D]], 0))
    _lin(m.final_layer.linear, sd, "proj_out")

    g = torch.Generator().manual_seed(5)
    B, hw, n_txt = 2, 6, 16
    img = torch.randn(B, hw * hw, cfg.in_channels, generator=g)
    txt = torch.randn(B, n_txt, cfg.joint_attention_dim, generator=g)
    pooled = torch.randn(B, cfg.pooled_projection_dim, generator=g)
    img_ids = torch.zeros(hw, hw, 3)
    img_ids[..., 1] += torch.arange(hw)[:, None]
    img_ids[..., 2] += torch.arange(hw)[None, :]
    img_ids = img_ids.reshape(hw * hw, 3)
    txt_ids = torch.zeros(n_txt, 3)
    t = torch.tensor([0.5051124692])
    with torch.no_grad():
        ours = nets.flux_forward(sd, cfg, img, t, torch.ones(B), pooled, txt, txt_ids, img_ids)
        theirs = m(img, img_ids[None].expand(B, -1, -1), txt, txt_ids[None].expand(B, -1, -1), t.expand(B), pooled)
    err = (ours - theirs).abs().max().item() / theirs.abs().max().item()
    print(f"oracle FLUX vs BFL port: max rel err {err:.2e} (output scale {theirs.abs().max():.3f})")
    assert err < 1e-4


def _res(dst, sd, name):
    for d, s in (("norm1", "norm1"), ("conv1", "conv1"), ("norm2", "norm2"), ("conv2", "conv2")):
        _lin(getattr(dst, d), sd, f"{name}.{s}")
    if name + ".conv_shortcut.weight" in sd:
        _lin(dst.nin_shortcut, sd, name + ".conv_shortcut")


def _attn(dst, sd, name):
    _lin(dst.norm, sd, name + ".group_norm")
    for d, s in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
        w = sd[f"{name}.{s}.weight"]
        _copy(getattr(dst, d).weight, w.reshape(w.shape[0], w.shape[1], 1, 1))
        _copy(getattr(dst, d).bias, sd[f"{name}.{s}.bias"])


def test_flux_autoencoder_matches_bfl_reference_port():
    from torchtitan.experiments.flux.model import autoencoder as ae
    from omgsr_b200 import weights as W
    from oracle import nets
    cfg = W.TINY_FLUX_VAE
    ch = cfg.block_out_channels[0]
    mult = tuple(c // ch for c in cfg.block_out_channels)
    n = len(mult)
    sd = W.random_state_dict(W.vae_shapes(cfg), 3)
    enc = ae.Encoder(resolution=64, in_channels=3, ch=ch, ch_mult=mult, num_res_blocks=cfg.layers_per_block,
                     z_channels=cfg.latent_channels)
    dec = ae.Decoder(ch=ch, out_ch=3, ch_mult=mult, num_res_blocks=cfg.layers_per_block, in_channels=3, resolution=64,
                     z_channels=cfg.latent_channels)
    _lin(enc.conv_in, sd, "encoder.conv_in")
    for i in range(n):
        for j in range(cfg.layers_per_block):
            _res(enc.down[i].block[j], sd, f"encoder.down_blocks.{i}.resnets.{j}")
        if i < n - 1:
            _lin(enc.down[i].downsample.conv, sd, f"encoder.down_blocks.{i}.downsamplers.0.conv")
    _res(enc.mid.block_1, sd, "encoder.mid_block.resnets.0")
    _attn(enc.mid.attn_1, sd, "encoder.mid_block.attentions.0")
    _res(enc.mid.block_2, sd, "encoder.mid_block.resnets.1")
    _lin(enc.norm_out, sd, "encoder.conv_norm_out")
    _lin(enc.conv_out, sd, "encoder.conv_out")
    _lin(dec.conv_in, sd, "decoder.conv_in")
    _res(dec.mid.block_1, sd, "decoder.mid_block.resnets.0")
    _attn(dec.mid.attn_1, sd, "decoder.mid_block.attentions.0")
    _res(dec.mid.block_2, sd, "decoder.mid_block.resnets.1")
    for i in range(n):  # diffusers up_blocks.0 runs first (lowest resolution) = LDM up[n-1]
        up = dec.up[n - 1 - i]
        for j in range(cfg.layers_per_block + 1):
            _res(up.block[j], sd, f"decoder.up_blocks.{i}.resnets.{j}")
        if i < n - 1:
            _lin(up.upsample.conv, sd, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    _lin(dec.norm_out, sd, "decoder.conv_norm_out")
    _lin(dec.conv_out, sd, "decoder.conv_out")
    g = torch.Generator().manual_seed(8)
    x = torch.rand(2, 3, 64, 48, generator=g) * 2 - 1
    z = torch.randn(2, cfg.latent_channels, 8, 6, generator=g)
    with torch.no_grad():
        e_ours, e_theirs = nets.vae_encode_moments(sd, cfg, x), enc(x)
        d_ours, d_theirs = nets.vae_decode(sd, cfg, z), dec(z)
    e_err = (e_ours - e_theirs).abs().max().item() / e_theirs.abs().max().item()
    d_err = (d_ours - d_theirs).abs().max().item() / d_theirs.abs().max().item()
    print(f"oracle FLUX VAE vs BFL port: encoder {e_err:.2e}, decoder {d_err:.2e}")
    assert e_err < 1e-4 and d_err < 1e-4
