"""Mid-timestep calibration: t* = argmin_t |SNR(z_t) - SNR(z_L)| averaged over paired
(LQ, HQ) images — the offline search that yields 273 (OMGSR-S-512) / 244 (OMGSR-F-1024) in the reference
(mid_timestep/mid_timestep_sd.py:59-80, mid_timestep/mid_timestep_flux.py:94-108).

The only heavy part is the VAE encode of every pair, which runs on the sm_100a encoder (omgsr_b200.engine_vae);
the SNR sweep over the 1000 timesteps is a vectorised host computation on two scalars per batch.
The reference reads pairs from its Real-ESRGAN degradation dataset (out of scope here); this module takes any
iterable of (lq, hq) image batches in [-1, 1]."""
import math

import torch


def snr_losses(hq_latent, lq_latent, signal_scale_sq, noise_power):
    """|SNR1(t) - SNR2| for every t. SNR1(t) = mean(hq^2) * signal_scale_sq[t] / noise_power[t] is the SNR of the
    noised HQ latent at t; SNR2 = mean(hq^2) / mean((lq - hq)^2) is the SNR of the LQ latent seen as a noisy HQ."""
    p_hq = hq_latent.float().pow(2).mean()
    snr2 = p_hq / (lq_latent.float() - hq_latent.float()).pow(2).mean()
    snr1 = p_hq * signal_scale_sq / noise_power
    return (snr1 - snr2).abs()


def ddpm_tables(alphas_cumprod):
    """SD: z_t = sqrt(abar) z + sqrt(1 - abar) eps"""
    a = alphas_cumprod.double()
    return a, 1.0 - a


def flux_tables(sigmas):
    """FLUX: z_t = (1 - sigma) z + sigma eps, with sigma_t = schedule[-(t+1)]"""
    s = torch.tensor([sigmas[-(t + 1)] for t in range(len(sigmas))], dtype=torch.float64)
    return (1.0 - s) ** 2, s ** 2


class MidTimestepSearch:
    """accumulates the batch-size-weighted loss per timestep, like the reference's loss_accumulators."""

    def __init__(self, signal_scale_sq, noise_power):
        self.sig, self.noise = signal_scale_sq, noise_power
        self.acc = torch.zeros_like(signal_scale_sq)
        self.count = 0

    def update(self, hq_latent, lq_latent):
        b = hq_latent.shape[0]
        self.acc += snr_losses(hq_latent.cpu(), lq_latent.cpu(), self.sig, self.noise).double() * b
        self.count += b

    def result(self):
        avg = self.acc / max(self.count, 1)
        # the reference scans t = 999 .. 0 and python's min() keeps the first minimum, i.e. the largest t on ties
        best = float(avg.min())
        t = max(i for i in range(len(avg)) if float(avg[i]) == best)
        return t, float(avg[t]), avg


def calibrate(vae, pairs, family="sd", alphas_cumprod=None, sigmas=None, generator=None):
    """vae: B200AutoencoderKL; pairs: iterable of (lq [B,3,H,W], hq [B,3,H,W]) CUDA tensors in [-1,1]."""
    if family == "sd":
        from omgsr_b200.weights import ddpm_alphas_cumprod
        sig, noise = ddpm_tables(alphas_cumprod if alphas_cumprod is not None else ddpm_alphas_cumprod())
    else:
        from infer.omgsr_f_infer_model import get_flux_setting_timesteps
        sig, noise = flux_tables(sigmas if sigmas is not None else get_flux_setting_timesteps())
    search = MidTimestepSearch(sig, noise)
    sf, sh = vae.config.scaling_factor, (vae.config.shift_factor if family != "sd" else 0.0)
    with torch.no_grad():
        for lq, hq in pairs:
            zl = (vae.encode(lq).latent_dist.sample(generator) - sh) * sf
            zh = (vae.encode(hq).latent_dist.sample(generator) - sh) * sf
            search.update(zh, zl)
    return search.result()
