"""Golden vectors for the tiled VAE, produced by the REFERENCE's own infer/vaehook.py (VAEHook, task queue,
GroupNormParam) running on CPU over a stub net that exposes the diffusers attribute names the hook
dereferences (infer/vaehook.py:137-171,248-359) and carries this repo's seeded TINY_VAE weights.

Run in the authoring container only:  python tests/golden/make_golden_vaehook.py
Writes tests/golden/vaehook_fixtures.pt (inputs are regenerated from seeds by the tests; only outputs and
a checksum of the input are stored)."""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _conv(sd, name, stride=1, padding=1):
    w = sd[name + ".weight"]
    m = nn.Conv2d(w.shape[1], w.shape[0], w.shape[2], stride=stride, padding=padding)
    m.weight.data.copy_(w)
    m.bias.data.copy_(sd[name + ".bias"])
    return m


def _gn(sd, name, eps=1e-6):
    w = sd[name + ".weight"]
    m = nn.GroupNorm(32, w.shape[0], eps=eps)
    m.weight.data.copy_(w)
    m.bias.data.copy_(sd[name + ".bias"])
    return m


def _lin(sd, name):
    w = sd[name + ".weight"]
    m = nn.Linear(w.shape[1], w.shape[0])
    m.weight.data.copy_(w)
    m.bias.data.copy_(sd[name + ".bias"])
    return m


class Res(nn.Module):
    def __init__(self, sd, name):
        super().__init__()
        self.norm1, self.conv1 = _gn(sd, name + ".norm1"), _conv(sd, name + ".conv1")
        self.norm2, self.conv2 = _gn(sd, name + ".norm2"), _conv(sd, name + ".conv2")
        self.in_channels, self.out_channels = self.conv1.in_channels, self.conv1.out_channels
        self.use_in_shortcut = self.in_channels != self.out_channels
        self.conv_shortcut = _conv(sd, name + ".conv_shortcut", padding=0) if self.use_in_shortcut else None


class Attn(nn.Module):
    """the subset of diffusers.Attention that attn_forward_new touches (single head, no mask)."""

    def __init__(self, sd, name):
        super().__init__()
        self.group_norm = _gn(sd, name + ".group_norm")
        self.to_q, self.to_k, self.to_v = _lin(sd, name + ".to_q"), _lin(sd, name + ".to_k"), _lin(sd, name + ".to_v")
        self.to_out = nn.ModuleList([_lin(sd, name + ".to_out.0"), nn.Identity()])
        self.norm_cross = False
        self.scale = self.to_q.in_features ** -0.5

    def prepare_attention_mask(self, mask, n, b):
        return None

    def head_to_batch_dim(self, t):
        return t

    def batch_to_head_dim(self, t):
        return t

    def get_attention_scores(self, q, k, mask):
        return torch.softmax(torch.bmm(q, k.transpose(1, 2)) * self.scale, dim=-1)


class Down(nn.Module):
    def __init__(self, sd, name):
        super().__init__()
        self.conv = _conv(sd, name + ".conv", stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class Up(nn.Module):
    def __init__(self, sd, name):
        super().__init__()
        self.conv = _conv(sd, name + ".conv")

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Block(nn.Module):
    pass


def build_stub(sd, cfg, decoder):
    net = nn.Module()
    pre = "decoder" if decoder else "encoder"
    net.conv_in = _conv(sd, pre + ".conv_in")
    n = len(cfg.block_out_channels)
    blocks = nn.ModuleList()
    for i in range(n):
        b = Block()
        kind = "up_blocks" if decoder else "down_blocks"
        nres = cfg.layers_per_block + (1 if decoder else 0)
        b.resnets = nn.ModuleList([Res(sd, f"{pre}.{kind}.{i}.resnets.{j}") for j in range(nres)])
        if i < n - 1:
            if decoder:
                b.upsamplers = nn.ModuleList([Up(sd, f"{pre}.{kind}.{i}.upsamplers.0")])
            else:
                b.downsamplers = nn.ModuleList([Down(sd, f"{pre}.{kind}.{i}.downsamplers.0")])
        blocks.append(b)
    if decoder:
        net.up_blocks = blocks
    else:
        net.down_blocks = blocks
    net.mid_block = Block()
    net.mid_block.resnets = nn.ModuleList([Res(sd, pre + ".mid_block.resnets.0"), Res(sd, pre + ".mid_block.resnets.1")])
    net.mid_block.attentions = nn.ModuleList([Attn(sd, pre + ".mid_block.attentions.0")])
    net.conv_norm_out = _gn(sd, pre + ".conv_norm_out")
    net.conv_out = _conv(sd, pre + ".conv_out")
    net.give_pre_end = False
    net.tanh_out = False
    return net


def main():
    # the reference's `infer` is a namespace package and would lose against this repo's regular `infer`
    # package: bind the name to the reference directory explicitly for this process
    import types
    pkg = types.ModuleType("infer")
    pkg.__path__ = [os.path.join(REF, "infer")]
    sys.modules["infer"] = pkg
    import infer.vaehook as vaehook
    vaehook.devices.device = torch.device("cpu")  # GroupNormParam.summary builds its weights on devices.device
    from omgsr_b200 import weights as W
    from oracle import nets
    cfg = W.TINY_VAE
    sd = W.random_state_dict(W.vae_shapes(cfg), 0)
    gold = {}
    with torch.no_grad():
        # sanity: the stub net untiled == oracle nets (so the fixture really pins nets.* + vaehook together)
        enc = build_stub(sd, cfg, False)
        dec = build_stub(sd, cfg, True)
        g = torch.Generator().manual_seed(4321)
        x = torch.rand(1, 3, 192, 160, generator=g) * 2 - 1
        z = torch.randn(1, 4, 40, 48, generator=g)
        hook_e = vaehook.VAEHook(enc, 64, False, False, False, False)
        hook_d = vaehook.VAEHook(dec, 16, True, False, False, False)
        gold["enc_in_sum"] = x.double().sum()
        gold["dec_in_sum"] = z.double().sum()
        gold["enc_tiled"] = hook_e(x)   # [1, 8, 24, 20] pre-quant moments, fp32
        gold["dec_tiled"] = hook_d(z)   # [1, 3, 320, 384]
        # fast mode (GroupNorm statistics estimated once on a nearest-exact downsampled copy, infer/vaehook.py:714-735)
        gold["enc_tiled_fast"] = vaehook.VAEHook(build_stub(sd, cfg, False), 64, False, False, True, False)(x)
        gold["dec_tiled_fast"] = vaehook.VAEHook(build_stub(sd, cfg, True), 16, True, True, False, False)(z)
        # fast encoder with color_fix=True: only the sites before the first downsample use estimated statistics
        # (infer/vaehook.py:664-668), the rest are synchronised across tiles as in the exact mode
        gold["enc_tiled_fast_colorfix"] = vaehook.VAEHook(build_stub(sd, cfg, False), 64, False, False, True, True)(x)
        gold["enc_split"] = hook_e.split_tiles(192, 160)
        gold["dec_split"] = hook_d.split_tiles(40, 48)
        # untiled references from the oracle nets for the same inputs (pre quant_conv / no post_quant)
        cfg_nq = W.VAEConfig(block_out_channels=cfg.block_out_channels, use_quant_conv=False)
        gold["enc_untiled_oracle"] = nets.vae_encode_moments(sd, cfg_nq, x)
        gold["dec_untiled_oracle"] = nets.vae_decode(sd, cfg_nq, z)
    torch.save(gold, os.path.join(OUT, "vaehook_fixtures.pt"))
    print("enc tiled vs untiled max diff", (gold["enc_tiled"] - gold["enc_untiled_oracle"]).abs().max().item())
    print("dec tiled vs untiled max diff", (gold["dec_tiled"] - gold["dec_untiled_oracle"]).abs().max().item())
    print("enc fast vs exact tiled max diff", (gold["enc_tiled_fast"] - gold["enc_tiled"]).abs().max().item())
    print("enc fast+color_fix vs fast max diff", (gold["enc_tiled_fast_colorfix"] - gold["enc_tiled_fast"]).abs().max().item())
    print("dec fast vs exact tiled max diff", (gold["dec_tiled_fast"] - gold["dec_tiled"]).abs().max().item())
    print("wrote vaehook_fixtures.pt", os.path.getsize(os.path.join(OUT, "vaehook_fixtures.pt")))


if __name__ == "__main__":
    main()
