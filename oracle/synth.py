"""Seeded synthetic weights for the oracle / the reference arm of bench.py (TEST INFRASTRUCTURE ONLY).

Builds exactly the state dicts `infer.omgsr_s_infer_model.build_s_state("random://sd21?seed=S", "random://lora?seed=L")`
and `build_f_state("random://flux_dev?seed=S", "random://lora?seed=L&rank=R")` build - same `omgsr_synth` tables and
generator seeds - but merges the adapters with the oracle's own LoRA rule (oracle/lora.py) and imports nothing from the
product package, so `bench.py --impl reference` never loads the CUDA library."""
import torch

import omgsr_synth as S

from .lora import merge_lora


def s_state(name="sd21", seed=0, lora_seed=1, vae_rank=16, unet_rank=32):
    vae_cfg, unet_cfg = (S.SD21_VAE, S.SD21_UNET) if name == "sd21" else (S.TINY_VAE, S.TINY_UNET)
    vae_sd = S.random_state_dict(S.vae_shapes(vae_cfg), seed)
    unet_sd = S.random_state_dict(S.unet_shapes(unet_cfg), seed + 1)
    if lora_seed is not None:
        vae_ad = S.random_lora(S.vae_shapes(vae_cfg), S.LORA_TARGETS_SD, vae_rank, lora_seed, prefix="encoder.")
        unet_ad = S.random_lora(S.unet_shapes(unet_cfg), S.LORA_TARGETS_SD, unet_rank, lora_seed + 1)
        vae_sd, unet_sd = merge_lora(vae_sd, vae_ad), merge_lora(unet_sd, unet_ad)
    return vae_sd, unet_sd, vae_cfg, unet_cfg


def f_state(name="flux_dev", seed=0, lora_seed=1, rank=64, weight_dtype=torch.bfloat16, device="cpu"):
    vae_cfg, flux_cfg = (S.FLUX_VAE, S.FLUX_DEV) if name == "flux_dev" else (S.TINY_FLUX_VAE, S.TINY_FLUX)
    vae_sd = S.random_state_dict(S.vae_shapes(vae_cfg), seed)
    flux_sd = S.random_state_dict(S.flux_shapes(flux_cfg), seed + 1, dtype=weight_dtype, device=device)
    if lora_seed is not None:
        vae_ad = S.random_lora(S.vae_shapes(vae_cfg), S.LORA_TARGETS_SD, rank, lora_seed, prefix="encoder.")
        flux_ad = S.random_lora(S.flux_shapes(flux_cfg), S.LORA_TARGETS_FLUX, rank, lora_seed + 1, device=device)
        vae_sd, flux_sd = merge_lora(vae_sd, vae_ad), merge_lora(flux_sd, flux_ad)
    return vae_sd, flux_sd, vae_cfg, flux_cfg
