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


def test_flux_checkpoint_and_adapter_dirs(tmp_path):
    """load_flux_checkpoint (diffusers FLUX directory: vae/ + sharded transformer/) and the FLUX adapter directories
    {flux_adapter, vae_encoder_adapter} the reference loads (infer/omgsr_f_infer_model.py:115-116; target module names
    train/train_omgsr_f.py:155-169) round-trip to the same merged weights as the in-memory path."""
    from omgsr_b200 import weights as W
    from omgsr_b200.checkpoint import load_flux_checkpoint, load_peft_adapter, write_safetensors
    from infer.omgsr_f_infer_model import build_f_state
    vcfg, fcfg = W.TINY_FLUX_VAE, W.TINY_FLUX
    vae_sd = W.random_state_dict(W.vae_shapes(vcfg), 0)
    flux_sd = W.random_state_dict(W.flux_shapes(fcfg), 1)
    root = tmp_path / "flux"
    os.makedirs(root / "vae")
    os.makedirs(root / "transformer")
    write_safetensors(str(root / "vae" / "diffusion_pytorch_model.safetensors"), vae_sd)
    json.dump({"block_out_channels": list(vcfg.block_out_channels), "latent_channels": 16, "scaling_factor": 0.3611,
               "shift_factor": 0.1159, "use_quant_conv": False}, open(root / "vae" / "config.json", "w"))
    # the real checkpoint is sharded into three files: split the keys the same way
    keys = sorted(flux_sd)
    for i in range(3):
        part = {k: flux_sd[k] for k in keys[i::3]}
        write_safetensors(str(root / "transformer" / f"diffusion_pytorch_model-0000{i + 1}-of-00003.safetensors"), part)
    json.dump({"in_channels": 64, "num_layers": fcfg.num_layers, "num_single_layers": fcfg.num_single_layers,
               "num_attention_heads": fcfg.num_heads, "attention_head_dim": 128,
               "joint_attention_dim": fcfg.joint_attention_dim, "pooled_projection_dim": fcfg.pooled_projection_dim,
               "axes_dims_rope": [16, 56, 56], "guidance_embeds": True}, open(root / "transformer" / "config.json", "w"))
    v2, f2, vc2, fc2 = load_flux_checkpoint(str(root))
    assert fc2 == fcfg and vc2.latent_channels == 16 and not vc2.use_quant_conv and vc2.shift_factor == 0.1159
    assert set(f2) == set(flux_sd) and all(torch.equal(f2[k], flux_sd[k]) for k in flux_sd)
    # adapters: rank 8, lora_alpha 16 -> scale 2 must be folded in
    lora = tmp_path / "lora"
    vae_ad = W.random_lora(W.vae_shapes(vcfg), W.LORA_TARGETS_SD, 8, 7, prefix="encoder.")
    flux_ad = W.random_lora(W.flux_shapes(fcfg), W.LORA_TARGETS_FLUX, 8, 8)
    # suffix matching as PEFT does: `attn.to_q` etc. also hit the single blocks; proj_mlp / proj_out / norm*.linear do not
    mods = set(flux_ad)
    assert "single_transformer_blocks.0.attn.to_q" in mods and "transformer_blocks.0.attn.add_k_proj" in mods
    assert "x_embedder" in mods and "transformer_blocks.1.ff_context.net.2" in mods
    assert not any(m.endswith(("proj_mlp", "proj_out", "norm.linear", "norm1.linear", "context_embedder")) for m in mods)
    for sub, ad, strip in (("vae_encoder_adapter", vae_ad, "encoder."), ("flux_adapter", flux_ad, "")):
        os.makedirs(lora / sub)
        flat = {}
        for m, (a, b) in ad.items():
            # PEFT writes the adapter name into the key for non-default adapters: accept both spellings
            mid = ".default" if sub == "flux_adapter" else ""
            flat[f"base_model.model.{m[len(strip):]}.lora_A{mid}.weight"] = a
            flat[f"base_model.model.{m[len(strip):]}.lora_B{mid}.weight"] = b
        write_safetensors(str(lora / sub / "adapter_model.safetensors"), flat)
        json.dump({"r": 8, "lora_alpha": 16}, open(lora / sub / "adapter_config.json", "w"))
    got = load_peft_adapter(str(lora / "flux_adapter"))
    assert set(got) == mods
    mv, mf, _, _ = build_f_state(str(root), str(lora), torch.float32)
    ev, ef = W.merge_lora(vae_sd, vae_ad, scale=2.0), W.merge_lora(flux_sd, flux_ad, scale=2.0)
    assert all(torch.allclose(mv[k], ev[k], atol=1e-6) for k in ev)
    assert all(torch.allclose(mf[k], ef[k], atol=1e-6) for k in ef)
    assert sum(int(not torch.equal(mf[k], flux_sd[k])) for k in flux_sd) == len(flux_ad)


def test_clip_text_checkpoint_folder_roundtrip(tmp_path):
    """`<sd_path>/text_encoder` as diffusers writes it (config.json + model.safetensors with transformers' key names)
    -> read_clip_text_folder: same tensors, config fields mapped (infer/infer_omgsr_s.py:19-22 loads exactly this
    folder)."""
    import json
    import omgsr_synth as syn
    from omgsr_b200.checkpoint import write_safetensors
    cfg = syn.TINY_CLIP_QUICK
    sd = syn.random_state_dict(syn.clip_text_shapes(cfg), seed=3)
    folder = tmp_path / "text_encoder"
    folder.mkdir()
    write_safetensors(str(folder / "model.safetensors"), sd)
    json.dump({"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
               "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
               "max_position_embeddings": cfg.max_positions, "hidden_act": cfg.hidden_act, "layer_norm_eps": 1e-5,
               "eos_token_id": cfg.eos_token_id, "pad_token_id": cfg.pad_token_id}, open(folder / "config.json", "w"))
    try:
        from omgsr_b200.engine_clip import read_clip_text_folder
    except ImportError as e:   # the product package needs its CUDA library to import
        pytest.skip(str(e))
    sd2, cfg2 = read_clip_text_folder(str(folder))
    assert cfg2 == cfg
    assert set(sd2) == set(sd) and all(torch.equal(sd2[k], sd[k]) for k in sd)
