"""CPU tests that PIN the oracle against fixtures produced by the reference's own code
(tests/golden/make_golden.py, make_golden_vaehook.py), plus the host-side logic of the drop-in modules."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "reference_fixtures.pt"), weights_only=False)


def test_colorfix_oracle_matches_reference(gold):
    from oracle import colorfix
    g = gold["colorfix"]
    c, s = g["content"], g["style"]
    assert torch.allclose(colorfix.wavelet_reconstruction(c, s), g["wavelet"], atol=1e-6)
    assert torch.allclose(colorfix.adaptive_instance_normalization(c, s), g["adain"], atol=1e-6)
    assert torch.allclose(colorfix.atrous_blur(c, 4), g["blur_r4"], atol=1e-7)
    # the telescoping identity the CUDA kernel uses: out = content + LP5(style - content)
    _, low = colorfix.wavelet_decomposition(s - c)
    assert torch.allclose(c + low, g["wavelet"], atol=2e-6)


def test_colorfix_pil_wrappers_match_reference(gold):
    from PIL import Image
    from oracle import colorfix
    g = gold["colorfix_pil"]
    tgt = Image.fromarray(g["target"].numpy(), "RGB")
    src = Image.fromarray(g["source"].numpy(), "RGB")
    assert np.array_equal(np.asarray(colorfix.wavelet_color_fix(tgt, src)), g["wavelet"].numpy())
    assert np.array_equal(np.asarray(colorfix.adain_color_fix(tgt, src)), g["adain"].numpy())


def test_tile_planner_matches_reference(gold):
    from infer.vaehook import VAEHook
    from oracle import vaehook as ov
    for (is_dec, tile, h, w), (ib, ob) in gold["split_tiles"].items():
        hk = VAEHook(None, tile, is_dec, False, False, False)
        with contextlib.redirect_stdout(io.StringIO()):
            a, b = hk.split_tiles(h, w)
        assert a == ib and b == ob, (is_dec, tile, h, w)
        assert ov.split_tiles(h, w, tile, is_dec) == (ib, ob)
    hk = VAEHook(None, 256, False, False, False, False)
    for (lo, hi), v in gold["best_tile"].items():
        assert hk.get_best_tile_size(lo, hi) == v
        assert ov.get_best_tile_size(lo, hi) == v


def test_crop_margins_match_reference(gold):
    from infer.vaehook import VAEHook
    g = gold["crop"]
    x = g["x"]
    m = VAEHook.crop_margins([0, 6, 0, 5], [0, 37, 0, 29], True)
    assert torch.equal(x[:, :, m[2]:x.size(2) + m[3], m[0]:x.size(3) + m[1]], g["dec"])
    m = VAEHook.crop_margins([64, 448, 0, 320], [12, 52, 0, 36], False)
    assert torch.equal(x[:, :, m[2]:x.size(2) + m[3], m[0]:x.size(3) + m[1]], g["enc"])


def test_groupnorm_helpers_match_reference(gold):
    from oracle import ops
    g = gold["groupnorm"]
    x = g["x"]
    xr = x.reshape(2, 32, -1)
    var, mean = torch.var_mean(xr, dim=2, unbiased=False)
    assert torch.allclose(var.flatten(), g["var"], atol=1e-6) and torch.allclose(mean.flatten(), g["mean"], atol=1e-6)
    ref = ops.groupnorm_nhwc(x.permute(0, 2, 3, 1), g["w"], g["b"], 32, 1e-6).permute(0, 3, 1, 2)
    assert torch.allclose(ref, g["out"], atol=1e-5)


def test_flux_helpers_match_reference(gold):
    from infer import omgsr_f_infer_model as fm
    g = gold["flux"]
    sched = fm.get_flux_setting_timesteps()
    assert len(sched) == 1000
    assert torch.allclose(torch.tensor(sched, dtype=torch.float64), g["schedule"], atol=0, rtol=0)
    assert sched[-(244 + 1)] == g["t_curr_244"]
    lat = g["lat"]
    packed = fm._pack_latents(lat, 2, 16, 8, 12)
    assert torch.equal(packed, g["packed"])
    assert torch.equal(fm._unpack_latents(packed, 64, 96, 8), g["unpacked"])
    from infer.infer_omgsr_f import _prepare_latent_image_ids
    assert torch.equal(_prepare_latent_image_ids(6, 4, "cpu", torch.float32), g["img_ids"])


def test_gaussian_weights_match_reference(gold):
    from oracle import omgsr_s
    for (tw, th), ref in gold["gauss"].items():
        mine = omgsr_s.gaussian_weights(tw, th)
        assert mine.dtype == torch.float64
        assert torch.allclose(mine, ref[0, 0], rtol=1e-14, atol=0)
        assert ref.shape == (1, 4, th, tw)


def test_tile_offsets_cover_reference_cases():
    from oracle import omgsr_s
    # latent 128 / tile 64 / overlap 32 -> offsets {0,32,64}; 256/128/64 -> {0,64,128}
    offs = omgsr_s.tile_offsets(128, 128, 64, 32)
    assert sorted(set(o[0] for o in offs)) == [0, 32, 64] and len(offs) == 9
    offs = omgsr_s.tile_offsets(256, 256, 128, 64)
    assert sorted(set(o[1] for o in offs)) == [0, 64, 128]
    # non-square, last tile forced to dim - T
    offs = omgsr_s.tile_offsets(80, 100, 64, 32)
    assert offs[-1] == (100 - 64, 80 - 64)


def test_vaehook_oracle_matches_reference_vaehook():
    """oracle/vaehook.py (+ oracle/nets.py) against the reference's own VAEHook run on a stub net (CPU)."""
    from omgsr_b200 import weights as W
    from oracle import nets
    from oracle import vaehook as ov
    g = torch.load(os.path.join(GOLD, "vaehook_fixtures.pt"), weights_only=False)
    cfg = W.VAEConfig(block_out_channels=W.TINY_VAE.block_out_channels, use_quant_conv=False)
    sd = W.random_state_dict(W.vae_shapes(W.TINY_VAE), 0)
    gen = torch.Generator().manual_seed(4321)
    x = torch.rand(1, 3, 192, 160, generator=gen) * 2 - 1
    z = torch.randn(1, 4, 40, 48, generator=gen)
    assert x.double().sum() == g["enc_in_sum"] and z.double().sum() == g["dec_in_sum"]
    with torch.no_grad():
        assert torch.allclose(nets.vae_encode_moments(sd, cfg, x), g["enc_untiled_oracle"], atol=1e-5)
        enc = ov.tiled_forward(lambda t: nets.vae_encode_moments(sd, cfg, t), x, 64, False)
        dec = ov.tiled_forward(lambda t: nets.vae_decode(sd, cfg, t), z, 16, True)
    assert enc.shape == g["enc_tiled"].shape and dec.shape == g["dec_tiled"].shape
    assert (enc - g["enc_tiled"]).abs().max() < 2e-4, (enc - g["enc_tiled"]).abs().max()
    assert (dec - g["dec_tiled"]).abs().max() < 2e-4, (dec - g["dec_tiled"]).abs().max()
    # fast mode: statistics estimated once on a nearest-exact downsampled copy, then replayed by every tile
    with torch.no_grad():
        enc_f = ov.tiled_forward_fast(lambda t: nets.vae_encode_moments(sd, cfg, t), x, 64, False)
        dec_f = ov.tiled_forward_fast(lambda t: nets.vae_decode(sd, cfg, t), z, 16, True)
    assert (enc_f - g["enc_tiled_fast"]).abs().max() < 2e-4, (enc_f - g["enc_tiled_fast"]).abs().max()
    assert (dec_f - g["dec_tiled_fast"]).abs().max() < 2e-4, (dec_f - g["dec_tiled_fast"]).abs().max()
    assert (enc_f - enc).abs().max() > 1e-2  # and it really is a different result from the exact mode
    # reference quirk: color_fix=True changes nothing (its 'downsample' branch never matches OMGSR's task names)
    assert torch.equal(g["enc_tiled_fast_colorfix"], g["enc_tiled_fast"])


def test_weight_tables_reproduce_public_parameter_counts():
    from omgsr_b200 import weights as W
    n = lambda shapes: sum(int(np.prod(s)) for s in shapes.values())
    assert n(W.vae_shapes(W.SD21_VAE)) == 83_653_863
    assert n(W.unet_shapes(W.SD21_UNET)) == 865_910_724
    assert abs(n(W.flux_shapes(W.FLUX_DEV)) - 11.90e9) < 0.01e9


def test_lora_merge_matches_definition():
    from omgsr_b200 import weights as W
    shapes = W.vae_shapes(W.TINY_VAE)
    sd = W.random_state_dict(shapes, 0)
    ad = W.random_lora(shapes, W.LORA_TARGETS_SD, 4, 1, prefix="encoder.")
    assert all(k.startswith("encoder.") for k in ad)
    assert "encoder.conv_in" in ad and "encoder.down_blocks.0.downsamplers.0.conv" in ad
    assert "encoder.mid_block.attentions.0.to_out.0" in ad and "decoder.conv_in" not in ad
    merged = W.merge_lora(sd, ad)
    a, b = ad["encoder.down_blocks.1.resnets.0.conv1"]
    x = torch.randn(1, a.shape[1], 9, 9)
    base = torch.nn.functional.conv2d(x, sd["encoder.down_blocks.1.resnets.0.conv1.weight"], padding=1)
    lora = torch.nn.functional.conv2d(torch.nn.functional.conv2d(x, a, padding=1), b)  # PEFT Conv2d LoRA: B(A(x))
    full = torch.nn.functional.conv2d(x, merged["encoder.down_blocks.1.resnets.0.conv1.weight"], padding=1)
    assert torch.allclose(full, base + lora, atol=1e-5)
    assert torch.equal(merged["decoder.conv_in.weight"], sd["decoder.conv_in.weight"])


def test_ddpm_alpha_matches_survey_constant():
    from omgsr_b200 import weights as W
    a = W.ddpm_alphas_cumprod()[273].item()
    assert abs(a - 0.6357423) < 1e-6


def test_mid_timestep_search_matches_reference_loop():
    """vectorised SNR sweep == the reference's per-timestep python loop (mid_timestep/mid_timestep_sd.py:59-80,
    mid_timestep_flux.py:94-108), restated here literally on random latents."""
    import math
    from mid_timestep.calibrate import MidTimestepSearch, ddpm_tables, flux_tables
    from omgsr_b200 import weights as W
    from infer.omgsr_f_infer_model import get_flux_setting_timesteps
    g = torch.Generator().manual_seed(3)
    abar = W.ddpm_alphas_cumprod()
    sigmas = get_flux_setting_timesteps()
    for family in ("sd", "flux"):
        sig, noise = ddpm_tables(abar) if family == "sd" else flux_tables(sigmas)
        search = MidTimestepSearch(sig, noise)
        acc = {t: 0.0 for t in range(999, -1, -1)}
        n = 0
        for _ in range(3):
            hq = torch.randn(2, 4, 8, 8, generator=g)
            lq = hq + 0.6 * torch.randn(2, 4, 8, 8, generator=g)
            search.update(hq, lq)
            for t in acc:
                if family == "sd":
                    s2, n2 = math.sqrt(abar[t]) ** 2, math.sqrt(1 - abar[t]) ** 2
                else:
                    s2, n2 = (1 - sigmas[-(t + 1)]) ** 2, sigmas[-(t + 1)] ** 2
                snr1 = torch.mean(hq ** 2) * s2 / n2 if n2 > 0 else torch.tensor(float("inf"))
                snr2 = torch.mean(hq ** 2) / torch.mean((lq - hq) ** 2)
                acc[t] += torch.abs(snr1 - snr2).item() * 2
            n += 2
        ref = {t: acc[t] / n for t in acc}
        t_ref = min(ref, key=ref.get)
        t_got, loss, _ = search.result()
        assert t_got == t_ref, (family, t_got, t_ref)
        assert abs(loss - ref[t_ref]) <= 1e-4 * max(1.0, ref[t_ref])


def test_lora_merge_is_adapter_forward():
    """oracle/lora.py (PEFT merge rule, restated independently) == what the un-merged adapter computes:
    base(x) + B(A(x)) for a Linear and for a 3x3 Conv2d (stride / padding of the base conv on A, 1x1 B); and the product
    path's merge (omgsr_b200.weights.merge_lora) produces the same weights."""
    import torch.nn.functional as F
    from omgsr_b200 import weights as W
    from oracle import lora as OL
    g = torch.Generator().manual_seed(3)
    r = 4
    # Linear
    w, b = torch.randn(24, 16, generator=g), torch.randn(24, generator=g)
    a, bb = torch.randn(r, 16, generator=g), torch.randn(24, r, generator=g)
    x = torch.randn(5, 16, generator=g)
    merged = OL.merge_lora({"m.weight": w, "m.bias": b}, {"m": (a, bb)})
    assert torch.allclose(F.linear(x, merged["m.weight"], b), OL.adapter_forward(x, w, b, a, bb), atol=1e-5)
    # Conv2d 3x3, stride 2, padding 1
    w, b = torch.randn(12, 8, 3, 3, generator=g), torch.randn(12, generator=g)
    a, bb = torch.randn(r, 8, 3, 3, generator=g), torch.randn(12, r, 1, 1, generator=g)
    x = torch.randn(2, 8, 10, 10, generator=g)
    merged_c = OL.merge_lora({"c.weight": w}, {"c": (a, bb)})
    assert torch.allclose(F.conv2d(x, merged_c["c.weight"], b, stride=2, padding=1),
                          OL.adapter_forward(x, w, b, a, bb, stride=2, padding=1), atol=1e-4)
    # lora_alpha != r scales the delta
    m2 = OL.merge_lora({"c.weight": w}, {"c": (a, bb)}, lora_alpha=2 * r)
    assert torch.allclose(m2["c.weight"] - w, 2 * (merged_c["c.weight"] - w), atol=1e-5)
    # product merge == oracle merge
    prod = W.merge_lora({"c.weight": w, "m.weight": merged["m.weight"]}, {"c": (a, bb)})
    assert torch.allclose(prod["c.weight"], merged_c["c.weight"], atol=1e-6)


def test_oracle_synth_states_equal_product_states():
    """bench.py's reference arm builds its weights through oracle/synth.py without importing the product package;
    they must be the very tensors build_s_state / build_f_state give the CUDA path."""
    from infer.omgsr_s_infer_model import build_s_state
    from infer.omgsr_f_infer_model import build_f_state
    from oracle import synth
    pv, pu, pvc, puc = build_s_state("random://tiny?seed=0", "random://lora?seed=1")
    ov, ou, ovc, ouc = synth.s_state("tiny", 0, 1)
    assert pvc == ovc and puc == ouc and set(pv) == set(ov) and set(pu) == set(ou)
    assert all(torch.allclose(pv[k], ov[k], atol=1e-6) for k in pv) and all(torch.allclose(pu[k], ou[k], atol=1e-6) for k in pu)
    pv, pf, _, pfc = build_f_state("random://tiny?seed=0", "random://lora?seed=1&rank=8", torch.float32)
    ov, of_, _, ofc = synth.f_state("tiny", 0, 1, 8, torch.float32)
    assert pfc == ofc and all(torch.allclose(pf[k], of_[k], atol=1e-6) for k in pf)
    assert all(torch.allclose(pv[k], ov[k], atol=1e-6) for k in pv)


def test_reference_arm_does_not_load_the_cuda_library():
    """`bench.py --impl reference` must not import omgsr_b200 / infer (the CPU arm dlopen()ed the .so)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, types; sys.argv=['bench.py']; import importlib.util as u; "
            f"s=u.spec_from_file_location('bench', r'{os.path.join(root, 'bench.py')}'); b=u.module_from_spec(s); s.loader.exec_module(b); "
            "o=b._oracle_s(); x=b._s_inputs(1); "
            "bad=[m for m in sys.modules if m.startswith(('omgsr_b200','infer'))]; "
            "maps=open('/proc/self/maps').read(); "
            "print('BAD', bad, 'libomgsr' in maps)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "BAD [] False" in r.stdout, r.stdout


# ------------------------------------------------------------------------------------------------
# text conditioning: the oracle's CLIP text tower against transformers' own CLIPTextModel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["TINY_CLIP", "TINY_CLIP_QUICK", "SD21_CLIP_H", "FLUX_CLIP_L"])
def test_clip_text_oracle_matches_transformers(name):
    """oracle/clip_text.py == transformers.CLIPTextModel on the same random weights (last hidden state and pooled
    output), tiny configs and the full SD-2.1 (23 x 1024, gelu) / OpenAI CLIP-L (12 x 768, quick_gelu) towers; ids: the
    empty prompt [BOS, EOS, pad...] and a random prompt with an early EOS."""
    transformers = pytest.importorskip("transformers")
    import omgsr_synth as syn
    from oracle import clip_text
    cfg = getattr(syn, name)
    sd = syn.random_state_dict(syn.clip_text_shapes(cfg), seed=11)
    hf_cfg = transformers.CLIPTextConfig(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, max_position_embeddings=cfg.max_positions,
        hidden_act=cfg.hidden_act, layer_norm_eps=cfg.layer_norm_eps, eos_token_id=cfg.eos_token_id,
        pad_token_id=cfg.pad_token_id, bos_token_id=0)
    model = transformers.CLIPTextModel(hf_cfg).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(5)
    rnd = torch.randint(1, cfg.eos_token_id - 1, (1, cfg.max_positions), generator=g)
    rnd[0, 9] = cfg.eos_token_id
    rnd[0, 10:] = cfg.pad_token_id
    ids = torch.cat([syn.empty_prompt_ids(cfg), rnd])
    with torch.no_grad():
        ref = model(input_ids=ids)
        h, pooled = clip_text.clip_text_forward(sd, cfg, ids)
    scale = float(ref.last_hidden_state.abs().max())
    assert float((h - ref.last_hidden_state).abs().max()) <= 2e-5 * scale
    assert float((pooled - ref.pooler_output).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("name", ["TINY_T5", "SMALL_T5"])
def test_t5_encoder_oracle_matches_transformers(name):
    """oracle/t5_encoder.py == transformers.T5EncoderModel (v1.1: gated gelu_new feed-forward, RMS norms, shared
    relative position bias) on the same random weights, 96 tokens (distances beyond max_exact and up to the logarithmic
    buckets) incl. the empty prompt [EOS, pad...]."""
    transformers = pytest.importorskip("transformers")
    import omgsr_synth as syn
    from oracle import t5_encoder
    cfg = getattr(syn, name)
    sd = syn.random_state_dict(syn.t5_encoder_shapes(cfg), seed=21)
    hf_cfg = transformers.T5Config(
        vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.num_layers,
        num_heads=cfg.num_heads, relative_attention_num_buckets=cfg.num_buckets,
        relative_attention_max_distance=cfg.max_distance, feed_forward_proj="gated-gelu",
        layer_norm_epsilon=cfg.layer_norm_eps, dropout_rate=0.0)
    model = transformers.T5EncoderModel(hf_cfg).eval()
    sd_hf = dict(sd)
    sd_hf["encoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = model.load_state_dict(sd_hf, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    g = torch.Generator().manual_seed(6)
    ids = torch.cat([syn.empty_prompt_ids_t5(cfg, 96), torch.randint(2, cfg.vocab_size, (1, 96), generator=g)])
    with torch.no_grad():
        ref = model(input_ids=ids).last_hidden_state
        out = t5_encoder.t5_encoder_forward(sd, cfg, ids)
    assert float((out - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


def test_t5_relative_position_buckets_match_transformers():
    """the bucket function of the product engine (omgsr_b200/engine_t5.py) and of the oracle == transformers'
    T5Attention._relative_position_bucket (bidirectional, 32 buckets, max distance 128) for every offset that occurs
    in a 512-token sequence; the empty-prompt ids of both tokenisers."""
    transformers = pytest.importorskip("transformers")
    from transformers.models.t5.modeling_t5 import T5Attention
    import omgsr_synth as syn
    from oracle import t5_encoder
    rel = torch.arange(-511, 512)[None, :]
    ref = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(t5_encoder.relative_position_bucket(rel, 32, 128), ref)
    try:
        from omgsr_b200.engine_t5 import _relative_position_bucket
    except ImportError as e:  # the product package needs its CUDA library to import
        pytest.skip(str(e))
    assert torch.equal(_relative_position_bucket(rel, 32, 128), ref)
    ids = syn.empty_prompt_ids(syn.SD21_CLIP_H)
    assert ids.shape == (1, 77) and ids[0, :3].tolist() == [49406, 49407, 0]
    assert syn.empty_prompt_ids(syn.FLUX_CLIP_L)[0, :3].tolist() == [49406, 49407, 49407]
    t5 = syn.empty_prompt_ids_t5(syn.FLUX_T5_XXL)
    assert t5.shape == (1, 512) and t5[0, :2].tolist() == [1, 0]
