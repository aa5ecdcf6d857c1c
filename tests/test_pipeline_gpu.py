"""End-to-end parity of the OMGSR-S module (CUDA path through the C ABI) against the fp32 CPU oracle on
identical synthetic inputs, identical random-init + LoRA-merged weights and the same posterior noise.
Bars from BASELINE.json north_star: SR output PSNR >= 40 dB vs the fp32 result, latent max-relative
error <= 2e-2."""
import io
import contextlib
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


def _build(spec, lora, dtype=torch.bfloat16):
    from infer.omgsr_s_infer_model import OMGSR_S_Infer, build_s_state
    from omgsr_b200 import weights as W
    from oracle.omgsr_s import OracleS
    vae_sd, unet_sd, vae_cfg, unet_cfg = build_s_state(spec, lora)
    state = {"vae": vae_sd, "unet": unet_sd, "vae_cfg": vae_cfg, "unet_cfg": unet_cfg}
    with contextlib.redirect_stdout(io.StringIO()):
        model = OMGSR_S_Infer(state, None, 273, "cuda", dtype)
    oracle = OracleS(vae_sd, unet_sd, vae_cfg, unet_cfg, 273, W.ddpm_alphas_cumprod())
    return model, oracle


def _inputs(B, H, W, ctx_dim, cz, seed=1234):
    g = torch.Generator().manual_seed(seed)
    lq = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    ctx = torch.randn(1, 77, ctx_dim, generator=torch.Generator().manual_seed(seed + 1))
    eps = torch.randn(B, cz, H // 8, W // 8, generator=torch.Generator().manual_seed(seed + 2))
    return lq, ctx, eps


@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (2, 128, 96)])
def test_s_tiny_stages(B, H, W):
    """stage-by-stage: encoder latent, UNet eps, full image (tiny widths, same topology)."""
    from oracle import nets
    model, oracle = _build("random://tiny?seed=0", "random://lora?seed=5")
    lq, ctx, eps = _inputs(B, H, W, 128, 4)
    with torch.no_grad():
        z_ref = oracle.encode(lq, eps)
        model.vae.posterior_eps = eps
        z = model._encode(lq.cuda().bfloat16())
        e_latent = relerr(z, z_ref)
        print(f"latent relerr {e_latent:.4f}")
        assert e_latent <= 2e-2
        # UNet on the SAME latent (isolates the denoiser)
        zq = z.float().cpu()
        e_ref = nets.unet_forward(oracle.unet_sd, oracle.unet_cfg, zq, 273, ctx)
        e = model.unet(z, 273, encoder_hidden_states=ctx.cuda().bfloat16()).sample
        e_unet = relerr(e, e_ref)
        print(f"unet relerr {e_unet:.4f}")
        assert e_unet <= 2e-2
        # decoder on the same latent
        d_ref = nets.vae_decode(oracle.vae_sd, oracle.vae_cfg, zq / 0.18215)
        d = model.vae.decode((z.float() / 0.18215).bfloat16()).sample
        print(f"decoder psnr {psnr(d, d_ref):.2f}")
        assert psnr(d, d_ref) >= 40.0
        # whole module
        with contextlib.redirect_stdout(io.StringIO()):
            out, t = model(lq.cuda().bfloat16(), ctx.cuda().bfloat16(), 64, 32)
        ref = oracle.forward(lq, ctx, 64, 32, eps)
        p = psnr(out, ref)
        print(f"e2e psnr {p:.2f} dB, out std {ref.std():.3f}")
        assert out.shape == ref.shape and t > 0
        assert p >= 40.0


def test_s_tiny_latent_tiling():
    """latent 16x16 with tile 8 / overlap 4 -> 3x3 tiles batched into one UNet call + Gaussian blend."""
    model, oracle = _build("random://tiny?seed=3", "random://lora?seed=7")
    lq, ctx, eps = _inputs(1, 128, 128, 128, 4, seed=77)
    with torch.no_grad():
        model.vae.posterior_eps = eps
        with contextlib.redirect_stdout(io.StringIO()):
            out, _ = model(lq.cuda().bfloat16(), ctx.cuda().bfloat16(), 8, 4)
        ref = oracle.forward(lq, ctx, 8, 4, eps)
    p = psnr(out, ref)
    print(f"tiled e2e psnr {p:.2f} dB")
    assert p >= 40.0


def test_s_full_size_512():
    """BASELINE config 1 vs the CUDA path: SD-2.1 shapes, one 512x512 image, oracle in fp32 on the host."""
    model, oracle = _build("random://sd21?seed=0", "random://lora?seed=1")
    lq, ctx, eps = _inputs(1, 512, 512, 1024, 4)
    with torch.no_grad():
        model.vae.posterior_eps = eps
        z = model._encode(lq.cuda().bfloat16())
        z_ref = oracle.encode(lq, eps)
        e_latent = relerr(z, z_ref)
        with contextlib.redirect_stdout(io.StringIO()):
            out, t = model(lq.cuda().bfloat16(), ctx.cuda().bfloat16(), 64, 32)
        ref = oracle.forward(lq, ctx, 64, 32, eps)
    p = psnr(out, ref)
    print(f"S-512 full size: latent relerr {e_latent:.4f}, e2e psnr {p:.2f} dB, {t*1e3:.1f} ms, ref std {ref.std():.3f}")
    assert e_latent <= 2e-2
    assert p >= 40.0
