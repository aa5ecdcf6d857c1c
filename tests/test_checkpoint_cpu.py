"""on-disk formats: diffusers-style directories and PEFT adapter directories round-trip through
the dependency-free safetensors reader/writer and merge to the same weights as the in-memory path."""
import json
import os

import torch


def test_safetensors_roundtrip(tmp_path):
    from omgsr_b200.checkpoint import read_safetensors, write_safetensors
    t = {"a.weight": torch.randn(3, 5), "b": torch.randn(7).bfloat16(), "c": torch.arange(6).reshape(2, 3)}
    p = str(tmp_path / "x.safetensors")
    write_safetensors(p, t)
    back = read_safetensors(p)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and torch.equal(back[k], t[k])
    try:  # cross-check against the real library when it happens to be installed
        from safetensors.torch import load_file
        ref = load_file(p)
        for k in t:
            assert torch.equal(ref[k], t[k])
    except ImportError:
        pass


def test_sd_checkpoint_and_peft_adapter_dirs(tmp_path):
    from omgsr_b200 import weights as W
    from omgsr_b200.checkpoint import load_peft_adapter, load_sd_checkpoint, write_safetensors
    from infer.omgsr_s_infer_model import build_s_state
    vcfg, ucfg = W.TINY_VAE, W.TINY_UNET
    vae_sd = W.random_state_dict(W.vae_shapes(vcfg), 0)
    unet_sd = W.random_state_dict(W.unet_shapes(ucfg), 1)
    root = tmp_path / "sd"
    for sub, sd, cfg in (("vae", vae_sd, {"block_out_channels": list(vcfg.block_out_channels), "latent_channels": 4,
                                          "scaling_factor": 0.18215}),
                         ("unet", unet_sd, {"block_out_channels": list(ucfg.block_out_channels),
                                            "attention_head_dim": list(ucfg.num_heads), "cross_attention_dim": 128,
                                            "down_block_types": ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]})):
        os.makedirs(root / sub)
        write_safetensors(str(root / sub / "diffusion_pytorch_model.safetensors"), sd)
        json.dump(cfg, open(root / sub / "config.json", "w"))
    # PEFT adapters as train/train_omgsr_s.py:516-517 writes them
    lora = tmp_path / "lora"
    vae_ad = W.random_lora(W.vae_shapes(vcfg), W.LORA_TARGETS_SD, 4, 7, prefix="encoder.")
    unet_ad = W.random_lora(W.unet_shapes(ucfg), W.LORA_TARGETS_SD, 8, 8)
    for sub, ad, strip in (("vae_encoder_lora_adapter", vae_ad, "encoder."), ("unet_lora_adapter", unet_ad, "")):
        os.makedirs(lora / sub)
        flat = {}
        for m, (a, b) in ad.items():
            flat[f"base_model.model.{m[len(strip):]}.lora_A.weight"] = a
            flat[f"base_model.model.{m[len(strip):]}.lora_B.weight"] = b
        write_safetensors(str(lora / sub / "adapter_model.safetensors"), flat)
        json.dump({"r": 4, "lora_alpha": 4}, open(lora / sub / "adapter_config.json", "w"))
    v2, u2, vc2, uc2 = load_sd_checkpoint(str(root))
    assert vc2.block_out_channels == vcfg.block_out_channels and uc2.num_heads == ucfg.num_heads
    assert uc2.down_has_attn == ucfg.down_has_attn
    assert all(torch.equal(v2[k], vae_sd[k]) for k in vae_sd)
    got = load_peft_adapter(str(lora / "vae_encoder_lora_adapter"), prefix="encoder.")
    assert set(got) == set(vae_ad)
    # the module's loader path == in-memory merge
    mv, mu, _, _ = build_s_state(str(root), str(lora))
    ev, eu = W.merge_lora(vae_sd, vae_ad), W.merge_lora(unet_sd, unet_ad)
    assert all(torch.allclose(mv[k], ev[k]) for k in ev) and all(torch.allclose(mu[k], eu[k]) for k in eu)
    changed = sum(int(not torch.equal(mv[k], vae_sd[k])) for k in vae_sd)
    assert changed == len(vae_ad)
