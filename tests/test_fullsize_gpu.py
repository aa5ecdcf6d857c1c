"""Full-size parity of the CUDA path against the fp32 oracle evaluated ON THE GPU (TF32 off), bit-reproducibility of
repeated runs, and the cache / boundary regressions flagged in round 1.

BASELINE.json configs covered at their real widths:
  * configs[1]  OMGSR-S-512, batch 16 (two images of the batch against the oracle; UNet output at 2e-2)
  * configs[2]  OMGSR-S 256->1024 with `_init_tiled_vae()` defaults (4x4 encoder tiles, 9 latent tiles) vs oracle/vaehook.py
  * configs[3]  OMGSR-F-1024 at FLUX.1-dev widths (19 + 38 blocks, 3072 wide): latent, DiT output, image
Bars (north_star): latent max-relative error <= 2e-2, SR output PSNR >= 40 dB vs the fp32 result.
The fp32 oracle of the 12 B-parameter DiT is 47.6 GB of weights and ~90 TFLOP per image: a few seconds on a B200 in
plain fp32 (it does not fit a CPU test budget, which is why round 1 only had property tests at this size)."""
import contextlib
import io
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def psnr(a, b, peak=2.0):
    mse = (a.float().cpu() - b.float().cpu()).pow(2).mean().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-20))


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@contextlib.contextmanager
def exact_fp32():
    """plain fp32 on the GPU: no TF32 in matmuls or cuDNN convolutions (the reference's own
    infer/devices.py silently turns TF32 on; the parity target is true fp32)."""
    m, c = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = m, c


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


@pytest.fixture(scope="module")
def sd21():
    """SD-2.1-shape OMGSR-S module (bf16 CUDA path) + the fp32 oracle on the same device, same merged weights."""
    from infer.omgsr_s_infer_model import OMGSR_S_Infer, build_s_state
    from omgsr_b200 import weights as W
    from oracle.omgsr_s import OracleS
    vae_sd, unet_sd, vae_cfg, unet_cfg = build_s_state("random://sd21?seed=0", "random://lora?seed=1")
    state = {"vae": vae_sd, "unet": unet_sd, "vae_cfg": vae_cfg, "unet_cfg": unet_cfg}
    with _quiet():
        model = OMGSR_S_Infer(state, None, 273, "cuda", torch.bfloat16)
    oracle = OracleS(vae_sd, unet_sd, vae_cfg, unet_cfg, 273, W.ddpm_alphas_cumprod()).to("cuda", torch.float32)
    yield model, oracle
    del model, oracle
    torch.cuda.empty_cache()


def _s_inputs(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    lq = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    ctx = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(seed + 1))
    eps = torch.randn(B, 4, H // 8, W // 8, generator=torch.Generator().manual_seed(seed + 2))
    return lq, ctx, eps


def test_s512_batch16_vs_fp32_oracle(sd21):
    """the benched configuration itself (BASELINE configs[1]: batch 16): images 0 and 11 of the batch against the
    oracle, plus the UNet output on the same latent at the 2e-2 bar."""
    from oracle import nets
    model, oracle = sd21
    lq, ctx, eps = _s_inputs(16, 512, 512)
    with exact_fp32():
        model.vae.posterior_eps = eps
        z = model._encode(lq.cuda().bfloat16())
        with _quiet():
            out, _ = model(lq.cuda().bfloat16(), ctx.cuda().bfloat16(), 64, 32)
        e = model.unet(z[:2], 273, encoder_hidden_states=ctx.cuda().bfloat16()).sample
        e_ref = nets.unet_forward(oracle.unet_sd, oracle.unet_cfg, z[:2].float(), 273, ctx.cuda())
        e_unet = relerr(e, e_ref)
        res = []
        for i in (0, 11):
            z_ref = oracle.encode(lq[i:i + 1].cuda(), eps[i:i + 1].cuda())
            ref = oracle.forward(lq[i:i + 1].cuda(), ctx.cuda(), 64, 32, eps[i:i + 1].cuda())
            res.append((relerr(z[i:i + 1], z_ref), psnr(out[i:i + 1], ref)))
    model.vae.posterior_eps = None
    print(f"S-512 batch 16: UNet eps relerr {e_unet:.4f}; image 0: latent {res[0][0]:.4f}, {res[0][1]:.2f} dB; "
          f"image 11: latent {res[1][0]:.4f}, {res[1][1]:.2f} dB")
    assert e_unet <= 2e-2
    for lat, p in res:
        assert lat <= 2e-2 and p >= 40.0


def test_c3_s1024_tiled_vae_full_size(sd21):
    """BASELINE configs[2] at SD-2.1 widths: 256->1024, `_init_tiled_vae()` defaults (encoder tile 256 + pad 32 ->
    4x4 tiles; decoder tile 256: latent 128 <= 278 stays untiled, as in the reference), latent 128^2 -> 3x3 tiles of 64.
    Oracle: reference VAEHook semantics (oracle/vaehook.py) around the plain-torch VAE, per-tile UNet + Gaussian blend."""
    from oracle import nets
    from oracle import vaehook as ov
    from oracle.omgsr_s import gaussian_weights, tile_offsets
    from omgsr_b200 import weights as W
    model, oracle = sd21
    lq, ctx, eps = _s_inputs(1, 1024, 1024, seed=99)
    cfg, sd = oracle.vae_cfg, oracle.vae_sd
    cfg_nq = W.VAEConfig(block_out_channels=cfg.block_out_channels, use_quant_conv=False)
    with _quiet():
        model._init_tiled_vae()
    try:
        with exact_fp32():
            x = lq.cuda()
            mom = ov.tiled_forward(lambda t: nets.vae_encode_moments(sd, cfg_nq, t), x, 256, False)
            mom = torch.nn.functional.conv2d(mom, sd["quant_conv.weight"], sd["quant_conv.bias"])
            z = nets.posterior_sample(mom, eps.cuda()) * cfg.scaling_factor
            T, ov_ = 64, 32
            wts = gaussian_weights(T, T).cuda()
            noise, contrib = torch.zeros_like(z), torch.zeros_like(z)
            for ox, oy in tile_offsets(128, 128, T, ov_):
                o = nets.unet_forward(oracle.unet_sd, oracle.unet_cfg, z[:, :, oy:oy + T, ox:ox + T], 273, ctx.cuda())
                noise[:, :, oy:oy + T, ox:ox + T] += (o * wts).float()
                contrib[:, :, oy:oy + T, ox:ox + T] += wts.float()
            noise /= contrib
            z0 = (z - (1 - oracle.alpha_t).sqrt() * noise) / oracle.alpha_t.sqrt() / cfg.scaling_factor
            z0 = torch.nn.functional.conv2d(z0, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
            ref = ov.tiled_forward(lambda t: nets.vae_decode(sd, cfg_nq, t), z0, 256, True).clamp(-1, 1)
            model.vae.posterior_eps = eps
            z_got = model._encode(x.bfloat16())
            with _quiet():
                out, t = model(x.bfloat16(), ctx.cuda().bfloat16(), T, ov_)
    finally:
        model.vae.posterior_eps = None
        model.vae.encoder.__dict__.pop("forward", None)
        model.vae.decoder.__dict__.pop("forward", None)
    e_lat, p = relerr(z_got, z), psnr(out, ref)
    print(f"C3 S 256->1024 (tiled VAE 4x4 + 9 latent tiles) full size: latent relerr {e_lat:.4f}, image {p:.2f} dB, "
          f"{t * 1e3:.1f} ms")
    assert e_lat <= 2e-2 and p >= 40.0


def test_bit_reproducible_runs(sd21):
    """GroupNorm statistics are reduced in a fixed order (no floating-point atomics): the same input gives the same
    bits, at tiny and full widths, through the fused conv epilogues, the stand-alone statistics kernel (UNet 10/20/40
    channels per group), the four-launch upsample path and the tiled VAE."""
    from omgsr_b200 import capi
    from omgsr_b200.pack import pack_conv_weight, pack_upsample_weight
    model, _ = sd21
    lq, ctx, eps = _s_inputs(2, 512, 512, seed=5)
    with torch.no_grad(), _quiet():
        model.vae.posterior_eps = eps
        a, _ = model(lq.cuda().bfloat16(), ctx.cuda().bfloat16(), 64, 32)
        b, _ = model(lq.cuda().bfloat16(), ctx.cuda().bfloat16(), 64, 32)
        model.vae.posterior_eps = None
    assert torch.equal(a, b), f"S-512 runs differ: {psnr(a, b):.1f} dB"
    # kernel level
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 96, 160, 128, generator=g).cuda().bfloat16()
    w = pack_conv_weight(torch.randn(128, 128, 3, 3, generator=g) / 34).cuda()
    wu = pack_upsample_weight(torch.randn(128, 128, 3, 3, generator=g) / 34).cuda()
    bias = torch.zeros(128, device="cuda")
    for kw in ({}, {"upsample2x": True}, {"flags": capi.CONV_NO_HALO}, {"stride": 2, "pad": (0, 0, 1, 1)}):
        outs = []
        for _ in range(3):
            st = torch.empty(3, 32, 2, device="cuda")
            y = capi.conv2d(x, wu if kw.get("upsample2x") else w, bias, gn_stats=st, **kw)
            outs.append((y.clone(), st.clone()))
        ref = capi.groupnorm_stats(outs[0][0])
        assert all(torch.equal(o[1], outs[0][1]) and torch.equal(o[0], outs[0][0]) for o in outs), kw
        # and the fused statistics are the statistics of the stored tensor
        assert relerr(outs[0][1], ref) <= 1e-4, (kw, relerr(outs[0][1], ref))
    for C in (320, 640, 1280, 2560):
        xs = torch.randn(2, 32, 32, C, generator=g).cuda().bfloat16()
        s1, s2 = capi.groupnorm_stats(xs), capi.groupnorm_stats(xs)
        assert torch.equal(s1, s2)
        xf = xs.float().reshape(2, 32 * 32, 32, C // 32)
        want = torch.stack([xf.sum((1, 3)), (xf * xf).sum((1, 3))], -1)
        assert relerr(s1, want) <= 1e-4


def test_unet_prompt_cache_and_timestep_guard():
    """the cross-attention K/V cache must not alias a freed prompt tensor; a different prompt of the same
    shape right after the first one (both temporaries) must be honoured. And any timestep other than the folded one
    raises instead of being silently ignored."""
    from infer.omgsr_s_infer_model import OMGSR_S_Infer, build_s_state
    vae_sd, unet_sd, vae_cfg, unet_cfg = build_s_state("random://tiny?seed=0", "random://lora?seed=5")
    state = {"vae": vae_sd, "unet": unet_sd, "vae_cfg": vae_cfg, "unet_cfg": unet_cfg}
    with _quiet():
        model = OMGSR_S_Infer(state, None, 273, "cuda", torch.bfloat16)
    z = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(1)).cuda().bfloat16()

    def prompt(seed):  # a fresh fp32 tensor every time: converted to a bf16 temporary inside the engine
        return torch.randn(1, 77, 128, generator=torch.Generator().manual_seed(seed)).cuda()

    with torch.no_grad():
        model.unet.cache_kv = False
        want1 = model.unet(z, 273, encoder_hidden_states=prompt(1)).sample
        want2 = model.unet(z, 273, encoder_hidden_states=prompt(2)).sample
        model.unet.cache_kv = True
        got1 = model.unet(z, 273, encoder_hidden_states=prompt(1)).sample
        got2 = model.unet(z, 273, encoder_hidden_states=prompt(2)).sample
        p = prompt(2)
        got2b = model.unet(z, 273, encoder_hidden_states=p).sample   # cache fill ...
        got2c = model.unet(z, 273, encoder_hidden_states=p).sample   # ... and hit
        p.mul_(0.5)                                                   # in-place change of the SAME object: version bump
        got3 = model.unet(z, 273, encoder_hidden_states=p).sample
        want3 = model.unet(z, 273, encoder_hidden_states=prompt(2) * 0.5).sample
    assert not torch.equal(want1, want2)
    assert torch.equal(got1, want1) and torch.equal(got2, want2) and torch.equal(got2b, want2) and torch.equal(got2c, want2)
    assert torch.equal(got3, want3)
    with pytest.raises(ValueError):
        model.unet(z, 100, encoder_hidden_states=prompt(1))
    with pytest.raises(ValueError):
        model.unet(z, torch.tensor([272]), encoder_hidden_states=prompt(1))


def test_flux_conditioning_cache_and_bf16_timestep_switch():
    """the AdaLN modulation vectors are cached per (t, guidance, pooled prompt); `bf16_timestep_quirk` reproduces the
    conditioning of the reference's bf16 run (diffusers casts t to bf16 before x1000: 504 instead of 505.11)."""
    import torch.nn.functional as F
    from infer.omgsr_f_infer_model import OMGSR_F_Infer, build_f_state
    from oracle import nets
    vae_sd, flux_sd, vae_cfg, flux_cfg = build_f_state("random://tiny?seed=0", "random://lora?seed=1&rank=8", torch.float32)
    state = {"vae": vae_sd, "flux": flux_sd, "vae_cfg": vae_cfg, "flux_cfg": flux_cfg}
    with _quiet():
        model = OMGSR_F_Infer(state, None, "cuda", torch.bfloat16)
    ft = model.flux_transformer
    pp = torch.randn(1, flux_cfg.pooled_projection_dim, generator=torch.Generator().manual_seed(3)).cuda().bfloat16()
    t, gd = torch.tensor([model.t_curr]), torch.full((2,), 1.0)

    def expected(t1000):
        sd = {k: v.float() for k, v in flux_sd.items()}
        emb = lambda name, x: F.linear(F.silu(F.linear(x, sd[name + ".linear_1.weight"], sd[name + ".linear_1.bias"])),
                                       sd[name + ".linear_2.weight"], sd[name + ".linear_2.bias"])
        temb = emb("time_text_embed.timestep_embedder", nets._sincos256(torch.tensor([t1000])))
        temb = temb + emb("time_text_embed.guidance_embedder", nets._sincos256(torch.tensor([1000.0])))
        temb = temb + emb("time_text_embed.text_embedder", pp.float().cpu())
        sc, sh = F.linear(F.silu(temb), sd["norm_out.linear.weight"], sd["norm_out.linear.bias"]).chunk(2, dim=1)
        return torch.cat([1 + sc, sh], 1)  # the engine folds the 1 of (1 + scale) into the bias

    def out_mod(mod):
        o = ft.mod_off[("out",)]
        return mod[:1, o:o + 2 * flux_cfg.dim]

    with torch.no_grad():
        m1 = ft._conditioning(t, gd, pp, 2, "cuda")
        m2 = ft._conditioning(torch.tensor([model.t_curr]), torch.full((2,), 1.0), pp, 2, "cuda")
        assert m2 is m1, "same (t, guidance, prompt) must hit the cache"
        m3 = ft._conditioning(torch.tensor([0.25]), gd, pp, 2, "cuda")
        assert m3 is not m1 and not torch.equal(m3, m1)
        e_fp32 = relerr(out_mod(m1), expected(model.t_curr * 1000.0))
        ft.bf16_timestep_quirk = True
        mq = ft._conditioning(t, gd, pp, 2, "cuda")
        assert mq is not m1
        e_quirk = relerr(out_mod(mq), expected(504.0))
    print(f"modulation vs fp32 expectation: t=505.11 {e_fp32:.4f}; quirk t=504 {e_quirk:.4f}")
    assert e_fp32 <= 2e-2 and e_quirk <= 2e-2


def test_mid_timestep_calibration_on_encoder_kernels():
    """mid_timestep/calibrate.py on the CUDA VAE encoder against the same search over oracle latents."""
    from mid_timestep.calibrate import MidTimestepSearch, calibrate, ddpm_tables
    from omgsr_b200 import weights as W
    from omgsr_b200.engine_vae import B200AutoencoderKL
    from oracle import nets
    sd = W.random_state_dict(W.vae_shapes(W.TINY_VAE), 0)
    vae = B200AutoencoderKL(sd, W.TINY_VAE, "cuda", torch.bfloat16)
    g = torch.Generator().manual_seed(7)
    pairs, want = [], MidTimestepSearch(*ddpm_tables(W.ddpm_alphas_cumprod()))
    eps = torch.randn(2, 4, 16, 16, generator=g)
    for _ in range(2):
        hq = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
        lq = (hq + 0.35 * torch.randn(2, 3, 128, 128, generator=g)).clamp(-1, 1)
        pairs.append((lq.cuda().bfloat16(), hq.cuda().bfloat16()))
        with torch.no_grad():
            zl = nets.posterior_sample(nets.vae_encode_moments(sd, W.TINY_VAE, lq), eps) * 0.18215
            zh = nets.posterior_sample(nets.vae_encode_moments(sd, W.TINY_VAE, hq), eps) * 0.18215
        want.update(zh, zl)
    vae.posterior_eps = eps
    t_got, loss_got, _ = calibrate(vae, pairs, family="sd")
    t_want, loss_want, _ = want.result()
    print(f"mid-timestep search: CUDA encoder t* = {t_got}, oracle t* = {t_want}")
    assert 0 < t_want < 999 and abs(t_got - t_want) <= 3


def test_f1024_full_size_vs_fp32_oracle():
    """BASELINE configs[3] at FLUX.1-dev widths against the fp32 oracle run on the GPU: latent <= 2e-2, DiT output on
    the same packed latent <= 2e-2, image >= 40 dB."""
    from infer.infer_omgsr_f import _prepare_latent_image_ids
    from infer.omgsr_f_infer_model import OMGSR_F_Infer, build_f_state, encode_images
    from oracle.omgsr_f import OracleF, pack_latents
    torch.cuda.empty_cache()
    vae_sd, flux_sd, vae_cfg, flux_cfg = build_f_state("random://flux_dev?seed=0", "random://lora?seed=1&rank=64",
                                                       torch.float32, "cuda")
    state = {"vae": vae_sd, "flux": flux_sd, "vae_cfg": vae_cfg, "flux_cfg": flux_cfg}
    with _quiet():
        model = OMGSR_F_Infer(state, None, "cuda", torch.bfloat16)
    oracle = OracleF(vae_sd, flux_sd, vae_cfg, flux_cfg).to("cuda", torch.float32)
    g = torch.Generator().manual_seed(77)
    lq = (torch.rand(1, 3, 1024, 1024, generator=g) * 2 - 1).cuda()
    pe = torch.randn(1, 512, 4096, generator=g).cuda()
    pp = torch.randn(1, 768, generator=g).cuda()
    tid = torch.zeros(512, 3, device="cuda")
    ids = _prepare_latent_image_ids(64, 64, "cuda", torch.float32)
    eps = torch.randn(1, 16, 128, 128, generator=g).cuda()
    with exact_fp32():
        model.vae.posterior_eps = eps
        z = encode_images(lq.bfloat16(), model.vae, torch.bfloat16)
        z_ref = oracle.encode(lq, eps)
        packed = pack_latents(z.float())
        v = model._dit(pack_latents(z), pe.bfloat16(), pp.bfloat16(), tid.bfloat16(), ids.bfloat16())
        v_ref = oracle.dit(packed, pe, pp, tid, ids)
        with _quiet():
            out, t = model(lq.bfloat16(), pe.bfloat16(), pp.bfloat16(), tid.bfloat16(), ids.bfloat16(), 128, 64)
        ref = oracle.forward(lq, pe, pp, tid, ids, 128, 64, eps)
    e_lat, e_dit, p = relerr(z, z_ref), relerr(v, v_ref), psnr(out, ref)
    rms = ((v.float() - v_ref).pow(2).mean().sqrt() / v_ref.pow(2).mean().sqrt()).item()
    print(f"F-1024 full size vs fp32 oracle: latent relerr {e_lat:.4f}, DiT relerr {e_dit:.4f} (rms {rms:.4f}), "
          f"image {p:.2f} dB, {t * 1e3:.0f} ms; ref std {ref.std().item():.3f}")
    del model, oracle, state, flux_sd
    torch.cuda.empty_cache()
    assert e_lat <= 2e-2
    assert e_dit <= 2e-2
    assert p >= 40.0
