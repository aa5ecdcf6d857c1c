"""Oracle restatement of the reference's tiled VAE (infer/vaehook.py) — TEST INFRASTRUCTURE ONLY.

Semantics reproduced (non-fast mode, the reference's and OMGSR's default; fast mode: tiled_forward_fast):
  * inputs with max(H, W) <= 2*pad + tile_size bypass tiling (:554-556); pad = 32 px (encoder) / 11 latent px
    (decoder) (:546)
  * tile planner split_tiles (:577-634) / get_best_tile_size (:562-575)
  * each padded tile runs the whole net on its own (zero padding at tile borders, per-tile attention :243-246)
  * at every GroupNorm the per-tile biased variance and mean (get_var_mean :371-381) are averaged across tiles
    with weights pixels_i / sum(pixels) (GroupNormParam.summary :489-508) and applied with eps 1e-6
    (custom_group_norm :384-413)
  * the valid region of each tile is cropped (:416-427) and written into an fp32 result (:800-806)
Pinned by tests/golden/vaehook_fixtures.pt, which is produced by the reference's own VAEHook on CPU.

The nets are the plain functions of oracle/nets.py; tiles run in threads that meet at a barrier inside the
GroupNorm hook, which reproduces the lock-step of the reference's task queue without duplicating the nets."""
import math
import threading

import torch
import torch.nn.functional as F

from . import nets


def get_best_tile_size(lowerbound, upperbound):
    divider = 32
    while divider >= 2:
        rem = lowerbound % divider
        if rem == 0:
            return lowerbound
        cand = lowerbound - rem + divider
        if cand <= upperbound:
            return cand
        divider //= 2
    return lowerbound


def split_tiles(h, w, tile_size, is_decoder):
    pad = 11 if is_decoder else 32
    n_h = max(math.ceil((h - 2 * pad) / tile_size), 1)
    n_w = max(math.ceil((w - 2 * pad) / tile_size), 1)
    rh = get_best_tile_size(math.ceil((h - 2 * pad) / n_h), tile_size)
    rw = get_best_tile_size(math.ceil((w - 2 * pad) / n_w), tile_size)
    ins, outs = [], []
    for i in range(n_h):
        for j in range(n_w):
            core = [pad + j * rw, min(pad + (j + 1) * rw, w), pad + i * rh, min(pad + (i + 1) * rh, h)]
            out = [core[0] if core[0] > pad else 0, core[1] if core[1] < w - pad else w,
                   core[2] if core[2] > pad else 0, core[3] if core[3] < h - pad else h]
            outs.append([v * 8 if is_decoder else v // 8 for v in out])
            ins.append([max(0, core[0] - pad), min(w, core[1] + pad), max(0, core[2] - pad), min(h, core[3] + pad)])
    return ins, outs


class _SharedGroupNorm:
    """GroupNorm hook shared by the tile threads: barrier, pixel-weighted average of (mean, var), apply."""

    def __init__(self, n_tiles):
        self.n = n_tiles
        self.barrier = threading.Barrier(n_tiles)
        self.slots = [None] * n_tiles
        self.local = threading.local()

    def __call__(self, x, groups, weight, bias, eps):
        i = self.local.idx
        b, c = x.shape[:2]
        xr = x.reshape(b, groups, -1)
        var, mean = torch.var_mean(xr, dim=2, unbiased=False)  # per (b, group)
        self.slots[i] = (var, mean, x.shape[2] * x.shape[3])
        self.barrier.wait()
        pix = torch.tensor([s[2] for s in self.slots], dtype=torch.float32, device=x.device)
        wts = pix / pix.sum()
        var = sum(w * s[0] for w, s in zip(wts, self.slots))
        mean = sum(w * s[1] for w, s in zip(wts, self.slots))
        self.barrier.wait()  # everybody has read the slots before the next site overwrites them
        y = (xr - mean[..., None]) / torch.sqrt(var[..., None] + 1e-6)
        y = y.reshape(x.shape)
        return y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


class _EstimatedGroupNorm:
    """fast mode (infer/vaehook.py:637-676, 714-735): the first pass over a downsampled copy records (var, mean) per
    (image, group) at every GroupNorm site (GroupNormParam.from_tile :511-534); every tile then replays them in site
    order, so the tiles are independent of each other."""

    def __init__(self):
        self.sites = []
        self.recording = True
        self.k = 0

    def __call__(self, x, groups, weight, bias, eps):
        b = x.shape[0]
        xr = x.reshape(b, groups, -1)
        if self.recording:
            self.sites.append(torch.var_mean(xr, dim=2, unbiased=False))
            var, mean = self.sites[-1]
        else:
            var, mean = self.sites[self.k]
            self.k += 1
        y = ((xr - mean[..., None]) / torch.sqrt(var[..., None] + 1e-6)).reshape(x.shape)
        return y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def _stitch(x, results, ins, outs, is_decoder):
    _, _, h, w = x.shape
    oh, ow = (h * 8, w * 8) if is_decoder else (h // 8, w // 8)
    out = torch.zeros((x.shape[0], results[0].shape[1], oh, ow), dtype=torch.float32, device=x.device)
    for i, tile in enumerate(results):
        scaled = [v * 8 if is_decoder else v // 8 for v in ins[i]]
        m = [outs[i][k] - scaled[k] for k in range(4)]
        ob = outs[i]
        out[:, :, ob[2]:ob[3], ob[0]:ob[1]] = tile[:, :, m[2]:tile.shape[2] + m[3], m[0]:tile.shape[3] + m[1]].float()
    return out


def tiled_forward_fast(fn, x, tile_size, is_decoder):
    """VAEHook with fast_encoder / fast_decoder = True (color_fix False)."""
    pad = 11 if is_decoder else 32
    _, _, h, w = x.shape
    if max(h, w) <= 2 * pad + tile_size:
        return fn(x)
    ins, outs = split_tiles(h, w, tile_size, is_decoder)
    # downsample to the tile size with nearest-exact, then restore the per-channel mean / std the resampling distorted
    # and clamp to the original range (:716-731)
    down = F.interpolate(x, scale_factor=tile_size / max(h, w), mode="nearest-exact")
    std_old, mean_old = torch.std_mean(x, dim=[0, 2, 3], keepdim=True)
    std_new, mean_new = torch.std_mean(down, dim=[0, 2, 3], keepdim=True)
    down = torch.clamp((down - mean_new) / std_new * std_old + mean_old, min=x.min(), max=x.max())
    hook = _EstimatedGroupNorm()
    prev = nets.GN_HOOK
    nets.GN_HOOK = hook
    try:
        fn(down)
        hook.recording = False
        results = []
        for bx in ins:
            hook.k = 0
            results.append(fn(x[:, :, bx[2]:bx[3], bx[0]:bx[1]]))
    finally:
        nets.GN_HOOK = prev
    return _stitch(x, results, ins, outs, is_decoder)


def tiled_forward(fn, x, tile_size, is_decoder):
    """fn(tile) -> tile output (a function of oracle/nets.py that calls nets._gn for every GroupNorm)."""
    pad = 11 if is_decoder else 32
    _, _, h, w = x.shape
    if max(h, w) <= 2 * pad + tile_size:
        return fn(x)
    ins, outs = split_tiles(h, w, tile_size, is_decoder)
    hook = _SharedGroupNorm(len(ins))
    results = [None] * len(ins)
    errors = []

    def work(i):
        try:
            hook.local.idx = i
            bx = ins[i]
            results[i] = fn(x[:, :, bx[2]:bx[3], bx[0]:bx[1]])
        except Exception as e:  # pragma: no cover
            errors.append(e)
            hook.barrier.abort()

    prev = nets.GN_HOOK
    nets.GN_HOOK = hook
    try:
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(ins))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        nets.GN_HOOK = prev
    if errors:
        raise errors[0]
    oh, ow = (h * 8, w * 8) if is_decoder else (h // 8, w // 8)
    out = torch.zeros((x.shape[0], results[0].shape[1], oh, ow), dtype=torch.float32, device=x.device)
    for i, tile in enumerate(results):
        scaled = [v * 8 if is_decoder else v // 8 for v in ins[i]]
        m = [outs[i][k] - scaled[k] for k in range(4)]
        ob = outs[i]
        out[:, :, ob[2]:ob[3], ob[0]:ob[1]] = tile[:, :, m[2]:tile.shape[2] + m[3], m[0]:tile.shape[3] + m[1]].float()
    return out
