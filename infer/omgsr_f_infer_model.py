"""OMGSR-F one-step SR inference module — B200-native drop-in for the reference's
infer/omgsr_f_infer_model.py (same module-level helpers, class name, constructor and method signatures).

Dataflow (reference lines in brackets): FLUX-VAE encode, posterior sample, (z - shift) * scale [:15-18,:325] ->
2x2 pack to tokens [:21-26] -> one FLUX DiT forward at sigma = schedule[-(244+1)] = 0.5051 [:191-200] -> Euler
step to sigma 0 [:202] -> unpack, un-scale [:204-210] -> VAE decode [:211]; latents above tile_size^2 use the
overlapping latent-tile path [:214-320].
"""
import math
import os
import time
from typing import Callable

import torch

from omgsr_b200 import capi
from omgsr_b200 import weights as W
from .vaehook import VAEHook


def encode_images(pixels: torch.Tensor, vae: torch.nn.Module, weight_dtype):
    pixel_latents = vae.encode(pixels.to(vae.dtype)).latent_dist.sample()
    pixel_latents = (pixel_latents - vae.config.shift_factor) * vae.config.scaling_factor
    return pixel_latents.to(weight_dtype)


def _pack_latents(latents, batch_size, num_channels_latents, height, width):
    """[B,C,h,w] -> [B,(h/2)(w/2),4C], channel order (c, dy, dx). CUDA tensors use the pack kernel."""
    if latents.is_cuda and latents.dtype == torch.bfloat16:
        return capi.pack_latents(latents)
    x = latents.reshape(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
    return x.permute(0, 2, 4, 1, 3, 5).reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)


def _unpack_latents(latents, height, width, vae_scale_factor):
    """inverse of _pack_latents; height/width are PIXEL sizes (latent size = 2 * (pixels // (2 * vae_scale_factor)))."""
    batch_size, _, channels = latents.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    if latents.is_cuda and latents.dtype == torch.bfloat16:
        return capi.unpack_latents(latents, channels // 4, h, w)
    x = latents.reshape(batch_size, h // 2, w // 2, channels // 4, 2, 2)
    return x.permute(0, 3, 1, 4, 2, 5).reshape(batch_size, channels // 4, h, w)


def time_shift(mu: float, sigma: float, t: torch.Tensor):
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_lin_function(x1: float = 256, y1: float = 0.5, x2: float = 4096, y2: float = 1.15) -> Callable[[float], float]:
    slope = (y2 - y1) / (x2 - x1)
    icpt = y1 - slope * x1
    return lambda x: slope * x + icpt


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                 shift: bool = True) -> list:
    # num_steps + 1 points from 1 to 0, resolution-dependent shift towards high sigma
    ts = torch.linspace(1, 0, num_steps + 1)
    if shift:
        mu = get_lin_function(y1=base_shift, y2=max_shift)(image_seq_len)
        ts = time_shift(mu, 1.0, ts)
    return ts.tolist()


def get_flux_setting_timesteps(n=999):
    return get_schedule(n, (1024 // 8) * (1024 // 8) // 4, shift=True)
