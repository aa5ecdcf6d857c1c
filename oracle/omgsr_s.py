"""Oracle restatement of OMGSR-S inference (TEST INFRASTRUCTURE ONLY).

Follows infer/omgsr_s_infer_model.py of the reference line by line: forward (:170-184), _forward_no_tile
(:74-86), _forward_tile (:88-168), _gaussian_weights (:56-71); the nets come from oracle/nets.py. The
unconditional torch.cuda.synchronize() calls (:171,181) are dropped so it runs on CPU; the posterior noise
is an explicit argument so both implementations consume the same epsilon."""
import math

import numpy as np
import torch

from . import nets


def gaussian_weights(tile_width, tile_height):
    # infer/omgsr_s_infer_model.py:56-71 (note the asymmetric midpoints: (W-1)/2 along x, H/2 along y)
    var = 0.01
    mid = (tile_width - 1) / 2
    xs = [math.exp(-(x - mid) * (x - mid) / (tile_width * tile_width) / (2 * var)) / math.sqrt(2 * math.pi * var)
          for x in range(tile_width)]
    mid = tile_height / 2
    ys = [math.exp(-(y - mid) * (y - mid) / (tile_height * tile_height) / (2 * var)) / math.sqrt(2 * math.pi * var)
          for y in range(tile_height)]
    return torch.tensor(np.outer(ys, xs))  # float64 [H, W]


def tile_offsets(h, w, tile_size, tile_overlap):
    """tile grid of _forward_tile (:94-118): returns [(ofs_x, ofs_y)] in the reference's visiting order
    ('row' walks x, 'col' walks y; index row*grid_cols + col)."""
    grid_rows = 0
    cur = 0
    while cur < w:
        cur = max(grid_rows * tile_size - tile_overlap * grid_rows, 0) + tile_size
        grid_rows += 1
    grid_cols = 0
    cur = 0
    while cur < h:
        cur = max(grid_cols * tile_size - tile_overlap * grid_cols, 0) + tile_size
        grid_cols += 1
    out = []
    ofs_x = ofs_y = 0
    for row in range(grid_rows):
        for col in range(grid_cols):
            if col < grid_cols - 1 or row < grid_rows - 1:
                ofs_x = max(row * tile_size - tile_overlap * row, 0)
                ofs_y = max(col * tile_size - tile_overlap * col, 0)
            if row == grid_rows - 1:
                ofs_x = w - tile_size
            if col == grid_cols - 1:
                ofs_y = h - tile_size
            out.append((ofs_x, ofs_y))
    return out


class OracleS:
    def __init__(self, vae_sd, unet_sd, vae_cfg, unet_cfg, mid_timestep, alphas_cumprod):
        self.vae_sd, self.unet_sd, self.vae_cfg, self.unet_cfg = vae_sd, unet_sd, vae_cfg, unet_cfg
        self.mid_timestep = mid_timestep
        self.alpha_t = alphas_cumprod[mid_timestep]

    def to(self, device=None, dtype=None):
        """move / cast the weights (the GPU-eager arm of bench.py and the full-size parity tests run the oracle on
        the device: fp32 with TF32 off for parity, bf16 as the PyTorch-eager baseline)."""
        mv = lambda sd: {k: v.to(device=device, dtype=dtype if v.is_floating_point() else None) for k, v in sd.items()}
        self.vae_sd, self.unet_sd = mv(self.vae_sd), mv(self.unet_sd)
        return self

    def encode(self, lq_img, eps_noise):
        mom = nets.vae_encode_moments(self.vae_sd, self.vae_cfg, lq_img)
        return nets.posterior_sample(mom, eps_noise) * self.vae_cfg.scaling_factor  # :173

    def _tail(self, lq_latent, model_pred):
        den = (lq_latent - (1 - self.alpha_t).sqrt() * model_pred) / self.alpha_t.sqrt()  # :80-82
        return nets.vae_decode(self.vae_sd, self.vae_cfg, den / self.vae_cfg.scaling_factor).clamp(-1, 1)  # :83-85

    def forward_no_tile(self, lq_latent, prompt_embeds):
        pred = nets.unet_forward(self.unet_sd, self.unet_cfg, lq_latent, self.mid_timestep, prompt_embeds)
        return self._tail(lq_latent, pred)

    def forward_tile(self, lq_latent, prompt_embeds, tile_size, tile_overlap):
        _, _, h, w = lq_latent.shape
        tile_size = min(tile_size, min(h, w))
        wts = gaussian_weights(tile_size, tile_size).to(lq_latent.device)
        offs = tile_offsets(h, w, tile_size, tile_overlap)
        noise = torch.zeros(lq_latent.shape, device=lq_latent.device)
        contrib = torch.zeros(lq_latent.shape, device=lq_latent.device)
        for ox, oy in offs:  # each tile is a batch-1 denoiser call (:129-133)
            tile = lq_latent[:, :, oy:oy + tile_size, ox:ox + tile_size]
            out = nets.unet_forward(self.unet_sd, self.unet_cfg, tile, self.mid_timestep, prompt_embeds)
            noise[:, :, oy:oy + tile_size, ox:ox + tile_size] += (out * wts).float()  # fp64 weights -> fp32 acc (:160)
            contrib[:, :, oy:oy + tile_size, ox:ox + tile_size] += wts.float()
        noise /= contrib
        return self._tail(lq_latent, noise.to(lq_latent.dtype))

    def forward(self, lq_img, prompt_embeds, tile_size, tile_overlap, eps_noise):
        z = self.encode(lq_img, eps_noise)
        _, _, h, w = z.shape
        if h * w <= tile_size * tile_size:  # :175
            return self.forward_no_tile(z, prompt_embeds)
        return self.forward_tile(z, prompt_embeds, tile_size, tile_overlap)
