"""Oracle restatement of OMGSR-F inference (TEST INFRASTRUCTURE ONLY).

Follows infer/omgsr_f_infer_model.py of the reference: encode_images (:15-18), _pack/_unpack_latents (:21-41),
the sigma schedule (:43-76, t_curr = schedule[-(244+1)], t_prev = 0), _forward_no_tile (:174-212),
_forward_tile (:214-320), forward (:322-336). Nets from oracle/nets.py; posterior noise is an explicit argument."""
import math

import torch

from . import nets
from .omgsr_s import gaussian_weights, tile_offsets


def pack_latents(z):
    b, c, h, w = z.shape
    return z.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, (h // 2) * (w // 2), c * 4)


def unpack_latents(t, h, w):
    b, _, ch = t.shape
    return t.reshape(b, h // 2, w // 2, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(b, ch // 4, h, w)


def flux_schedule(n=999, image_seq_len=(1024 // 8) * (1024 // 8) // 4):
    ts = torch.linspace(1, 0, n + 1)
    m = (1.15 - 0.5) / (4096 - 256)
    mu = m * image_seq_len + (0.5 - m * 256)
    return (math.exp(mu) / (math.exp(mu) + (1 / ts - 1) ** 1.0)).tolist()


class OracleF:
    def __init__(self, vae_sd, flux_sd, vae_cfg, flux_cfg, mid_timestep=244, guidance_scale=1.0):
        self.vae_sd, self.flux_sd, self.vae_cfg, self.flux_cfg = vae_sd, flux_sd, vae_cfg, flux_cfg
        sched = flux_schedule()
        self.t_curr, self.t_prev = sched[-(mid_timestep + 1)], sched[-1]
        self.guidance_scale = guidance_scale

    def to(self, device=None, dtype=None):
        mv = lambda sd: {k: v.to(device=device, dtype=dtype if v.is_floating_point() else None) for k, v in sd.items()}
        self.vae_sd, self.flux_sd = mv(self.vae_sd), mv(self.flux_sd)
        return self

    def encode(self, lq_img, eps_noise):
        mom = nets.vae_encode_moments(self.vae_sd, self.vae_cfg, lq_img)
        z = nets.posterior_sample(mom, eps_noise)
        return (z - self.vae_cfg.shift_factor) * self.vae_cfg.scaling_factor

    def dit(self, packed, prompt_embeds, pooled, text_ids, img_ids):
        b = packed.shape[0]
        return nets.flux_forward(self.flux_sd, self.flux_cfg, packed, torch.tensor([self.t_curr]),
                                 torch.full((b,), float(self.guidance_scale)), pooled.expand(b, -1),
                                 prompt_embeds.expand(b, -1, -1), text_ids, img_ids)

    def forward_no_tile(self, z, prompt_embeds, pooled, text_ids, img_ids):
        _, _, h, w = z.shape
        x = pack_latents(z)
        v = self.dit(x, prompt_embeds, pooled, text_ids, img_ids)
        x = x + (self.t_prev - self.t_curr) * v
        z0 = unpack_latents(x, h, w) / self.vae_cfg.scaling_factor + self.vae_cfg.shift_factor
        return nets.vae_decode(self.vae_sd, self.vae_cfg, z0)

    def forward_tile(self, z, prompt_embeds, pooled, text_ids, img_ids, tile_size, tile_overlap):
        _, _, h, w = z.shape
        tile_size = min(tile_size, min(h, w))
        wts = gaussian_weights(tile_size, tile_size).to(z.device)
        noise = torch.zeros(z.shape, device=z.device)
        contrib = torch.zeros(z.shape, device=z.device)
        for ox, oy in tile_offsets(h, w, tile_size, tile_overlap):
            tile = z[:, :, oy:oy + tile_size, ox:ox + tile_size]
            v = unpack_latents(self.dit(pack_latents(tile), prompt_embeds, pooled, text_ids, img_ids), tile_size, tile_size)
            noise[:, :, oy:oy + tile_size, ox:ox + tile_size] += (v * wts).float()
            contrib[:, :, oy:oy + tile_size, ox:ox + tile_size] += wts.float()
        noise /= contrib
        z0 = (z + (self.t_prev - self.t_curr) * noise) / self.vae_cfg.scaling_factor + self.vae_cfg.shift_factor
        return nets.vae_decode(self.vae_sd, self.vae_cfg, z0)

    def forward(self, lq_img, prompt_embeds, pooled, text_ids, img_ids, tile_size, tile_overlap, eps_noise):
        z = self.encode(lq_img, eps_noise)
        _, _, h, w = z.shape
        if h * w <= tile_size * tile_size:
            return self.forward_no_tile(z, prompt_embeds, pooled, text_ids, img_ids)
        return self.forward_tile(z, prompt_embeds, pooled, text_ids, img_ids, tile_size, tile_overlap)
