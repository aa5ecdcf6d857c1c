"""AutoencoderKL (SD-2.1 / FLUX VAE shapes) driven through the sm_100a kernels.

Mirrors the module surface the reference touches on diffusers' AutoencoderKL
(infer/omgsr_s_infer_model.py:11,24,84,173; infer/omgsr_f_infer_model.py:15-18,99,211; the sub-module
attribute names dereferenced by infer/vaehook.py:296-329,353): `.encoder`, `.decoder` (each with a
patchable `.forward`), `.quant_conv`, `.post_quant_conv`, `.config.scaling_factor/.shift_factor`,
`.dtype`, `.encode(x).latent_dist.sample()`, `.decode(z).sample`.

Internally everything is NHWC bf16; the NCHW<->NHWC change is fused into the thin conv_in / conv_out
kernels at the module boundary. Weights are repacked once at construction (omgsr_b200.pack).
"""
import math
from types import SimpleNamespace

import torch

from . import capi
from .pack import pack_conv_in_gemm_weight, pack_conv_in_weight, pack_conv_weight, pack_upsample_weight, pad8


def _dev(t, device, dtype=None):
    return t.detach().to(device=device, dtype=dtype if dtype is not None else t.dtype).contiguous()


class _Res:
    """ResnetBlock2D parameters in kernel layout."""

    def __init__(self, sd, name, device, extra_bias1=None):
        f32 = torch.float32
        self.g1, self.b1 = _dev(sd[name + ".norm1.weight"], device, f32), _dev(sd[name + ".norm1.bias"], device, f32)
        self.g2, self.b2 = _dev(sd[name + ".norm2.weight"], device, f32), _dev(sd[name + ".norm2.bias"], device, f32)
        self.w1 = _dev(pack_conv_weight(sd[name + ".conv1.weight"]), device)
        bias1 = sd[name + ".conv1.bias"].float()
        if extra_bias1 is not None:  # UNet: constant time-embedding projection folded into the conv bias
            bias1 = bias1 + extra_bias1.float()
        self.c1b = _dev(bias1, device, f32)
        self.w2 = _dev(pack_conv_weight(sd[name + ".conv2.weight"]), device)
        self.c2b = _dev(sd[name + ".conv2.bias"], device, f32)
        self.cin = sd[name + ".conv1.weight"].shape[1]
        self.cout = sd[name + ".conv1.weight"].shape[0]
        self.ws = None
        if name + ".conv_shortcut.weight" in sd:
            self.ws = _dev(pack_conv_weight(sd[name + ".conv_shortcut.weight"]), device)
            self.sb = _dev(sd[name + ".conv_shortcut.bias"], device, f32)


def thin_conv_in(x, w_simt, b_simt, w_gemm, b_gemm, cout, out=None):
    """3x3 conv from a thin NCHW input (Cin <= 16) to NHWC bf16. From 64 Ki pixels on: im2col (K padded to a multiple
    of 64) + tensor-core GEMM; below that the direct CUDA-core kernel (launch-latency bound either way)."""
    if x.shape[0] * x.shape[2] * x.shape[3] >= (1 << 16):
        return capi.conv_in_gemm(x.contiguous(), w_gemm, b_gemm, out=out)
    return capi.conv_in_nchw(x.contiguous(), w_simt, b_simt, cout, out=out)


def _stats_buf(x_like_B, device):
    return torch.empty((x_like_B, 32, 2), device=device, dtype=torch.float32)


class GnDeferred:
    """a GroupNorm + SiLU that has NOT been applied: the raw tensor plus the per-(image, channel) coefficients. The
    consuming 3x3 conv normalises its operand tiles in shared memory (capi.conv2d_gn), so the normalised tensor never
    exists in HBM."""
    __slots__ = ("x", "coef")

    def __init__(self, x, coef):
        self.x, self.coef = x, coef


def _deferrable(x, silu, fuse_cout):
    """may the GroupNorm of NHWC view `x` be left to a following plain 3x3 conv with `fuse_cout` output channels?"""
    if not (silu and fuse_cout):
        return False
    B, H, W, C = x.shape
    dense = x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1)
    return dense and capi.conv_gn_in_fusable(B, H, W, C, fuse_cout)


def conv_with_stats(x, w, bias, want_stats, **kw):
    """conv2d that also returns the GroupNorm (sum, sumsq) statistics of its output when the shape allows the fused
    epilogue path (None otherwise: the consumer then runs the separate statistics kernel). `x` may be a GnDeferred."""
    if isinstance(x, GnDeferred):
        assert set(kw) <= {"res", "out"}, f"a deferred GroupNorm feeds a plain 3x3 conv only, got {sorted(kw)}"
        B, H, W, _ = x.x.shape
        st = _stats_buf(B, x.x.device) if want_stats and capi.gn_fusable(B, H, W, w.shape[0]) else None
        return capi.conv2d_gn(x.x, x.coef, w, bias, res=kw.get("res"), out=kw.get("out"), gn_stats=st), st
    B, H, W = x.shape[0], x.shape[1], x.shape[2]
    if kw.get("stride", 1) == 2 and (H % 2 or W % 2):
        # odd extents (VAEHook fast mode estimates on a downsampled copy of arbitrary size): the stride-2 kernel views
        # the input as pixel pairs, so run it on a copy zero-padded to even extents - the same zeros the (0,1,0,1)
        # padding of the reference's Downsample2D supplies - and drop the one extra output row / column
        xp = torch.zeros((B, H + H % 2, W + W % 2, x.shape[3]), device=x.device, dtype=x.dtype)
        xp[:, :H, :W] = x
        y = capi.conv2d(xp, w, bias, **kw)
        oh, ow = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
        return y[:, :oh, :ow].contiguous(), None
    if kw.get("upsample2x"):
        H, W = 2 * H, 2 * W
    elif kw.get("stride", 1) == 2:
        H, W = H // 2, W // 2
    cout = w.shape[1] if kw.get("upsample2x") else w.shape[0]
    if want_stats and capi.gn_fusable(B, H, W, cout):
        st = _stats_buf(B, x.device)
        return capi.conv2d(x, w, bias, gn_stats=st, **kw), st
    return capi.conv2d(x, w, bias, **kw), None


def resnet_gen(r, x, out=None, x_stats=None, want_stats=False):
    """ResnetBlock2D as a coroutine: yields (x, gamma, beta, silu, stats|None) at each GroupNorm and receives the
    normalised tensor, so a driver can share statistics across tiles (infer/vaehook.py task queue). `stats` are
    the fused-epilogue (sum, sumsq) of the producing conv when available. Returns (y, stats of y | None)."""
    # 6th field: Cout of the plain 3x3 conv that consumes the normalised tensor (the driver may then defer the
    # normalisation to that conv, see GnDeferred)
    h = yield (x, r.g1, r.b1, True, x_stats, r.w1.shape[0])
    h, st = conv_with_stats(h, r.w1, r.c1b, True)
    h = yield (h, r.g2, r.b2, True, st, r.w2.shape[0])
    if r.ws is not None:
        skip = capi.conv2d(x, r.ws, r.sb, ksize=1, pad=(0, 0, 0, 0))
    else:
        skip = x
    return conv_with_stats(h, r.w2, r.c2b, want_stats, res=skip, out=out)


def drive_plain(gen, eps):
    """run a GroupNorm-yielding coroutine with ordinary per-tensor statistics."""
    try:
        req = next(gen)
        while True:
            x, g, b, silu, st, *rest = req
            if st is None:
                st = capi.groupnorm_stats(x)
            if _deferrable(x, silu, rest[0] if rest else None):
                hw = x.shape[1] * x.shape[2]
                req = gen.send(GnDeferred(x, capi.groupnorm_coef(st, g, b, x.shape[0], hw, 32, eps)))
            else:
                req = gen.send(capi.groupnorm_apply(x, st, g, b, 32, eps, silu))
    except StopIteration as stop:
        return stop.value


def drive_tiled(gens, eps, n_img, weight=1.0, tile_parallel=False):
    """advance several tile-group coroutines in lock step; at every GroupNorm the per-(tile,image)
    statistics of all groups are combined on the device (pixel-weighted mean of mean and of variance,
    GroupNormParam.summary, infer/vaehook.py:489-508) and applied to every group."""
    reqs = [next(g) for g in gens]
    while True:
        sums, pix = [], []
        for (x, _, _, _, st, *_rest) in reqs:
            sums.append(st if st is not None else capi.groupnorm_stats(x))
            hw = x.numel() // (x.shape[0] * x.shape[-1])
            pix.append(torch.full((x.shape[0],), float(hw) * weight, device=x.device, dtype=torch.float32))
        C = reqs[0][0].shape[-1]
        if tile_parallel:
            # tiles of this image are spread over several GPUs: all-reduce the weighted numerators (tiny), then divide
            from .dist import reduce_gn_partials
            num, wsum = capi.groupnorm_combine(torch.cat(sums, 0), torch.cat(pix, 0), n_img, C, raw=True)
            shared = reduce_gn_partials(num, wsum).contiguous()
        else:
            shared = capi.groupnorm_combine(torch.cat(sums, 0), torch.cat(pix, 0), n_img, C)
        outs = []
        for (x, g, b, silu, _st, *rest) in reqs:
            if _deferrable(x, silu, rest[0] if rest else None):
                coef = capi.groupnorm_coef(shared, g, b, x.shape[0], x.shape[1] * x.shape[2], 32, eps, shared_mod=n_img)
                outs.append(GnDeferred(x, coef))
            else:
                outs.append(capi.groupnorm_apply(x, shared, g, b, 32, eps, silu, shared_mod=n_img))
        nxt, done = [], []
        for g, o in zip(gens, outs):
            try:
                nxt.append(g.send(o))
            except StopIteration as stop:
                done.append(stop.value)
        if done:
            if nxt:
                raise RuntimeError("tile coroutines fell out of step")
            return done
        reqs = nxt


def drive_record(gen, eps):
    """VAEHook fast mode, first pass (infer/vaehook.py:637-676): run the coroutine on the downsampled copy with
    ordinary per-tensor statistics and record (mean, biased var) per (image, group) at every GroupNorm site."""
    sites = []
    try:
        req = next(gen)
        while True:
            x, g, b, silu, st, *_rest = req
            if st is None:
                st = capi.groupnorm_stats(x)
            n = float(x.shape[1] * x.shape[2] * (x.shape[3] // 32))
            mean = st[..., 0] / n
            sites.append(torch.stack([mean, (st[..., 1] / n - mean * mean).clamp_min_(0.0)], -1).contiguous())
            req = gen.send(capi.groupnorm_apply(x, st, g, b, 32, eps, silu))
    except StopIteration:
        return sites


def drive_fixed(gen, eps, n_img, sites):
    """VAEHook fast mode, tile pass: every GroupNorm site applies the statistics recorded by drive_record
    ([n_img, 32, 2] (mean, var) rows shared by batch index % n_img), so the tiles never synchronise."""
    k = 0
    try:
        req = next(gen)
        while True:
            x, g, b, silu, _st, *rest = req
            shared = sites[k]
            k += 1
            if _deferrable(x, silu, rest[0] if rest else None):
                coef = capi.groupnorm_coef(shared, g, b, x.shape[0], x.shape[1] * x.shape[2], 32, eps, shared_mod=n_img)
                req = gen.send(GnDeferred(x, coef))
            else:
                req = gen.send(capi.groupnorm_apply(x, shared, g, b, 32, eps, silu, shared_mod=n_img))
    except StopIteration as stop:
        return stop.value


def resnet_forward(r, x, eps, out=None):
    """x NHWC view -> NHWC (written into `out` if given). GroupNorm+SiLU -> conv1 -> GroupNorm+SiLU -> conv2 + skip."""
    return drive_plain(resnet_gen(r, x, out), eps)[0]


class _MidAttn:
    """single-head attention of the VAE mid block (d = C = 512): fused qkv projection, materialised fp32
    scores through the batched tcgen05 GEMMs (a 128 x 512 fp32 O tile alone would fill TMEM)."""

    def __init__(self, sd, name, device):
        f32 = torch.float32
        self.g, self.b = _dev(sd[name + ".group_norm.weight"], device, f32), _dev(sd[name + ".group_norm.bias"], device, f32)
        wq, wk, wv = (sd[f"{name}.{p}.weight"] for p in ("to_q", "to_k", "to_v"))
        if wq.dim() == 4:  # older checkpoints store the projections as 1x1 convs
            wq, wk, wv = (w.reshape(w.shape[0], w.shape[1]) for w in (wq, wk, wv))
        self.wqkv = _dev(torch.cat([wq, wk, wv], 0).to(torch.bfloat16), device)
        self.bqkv = _dev(torch.cat([sd[f"{name}.{p}.bias"] for p in ("to_q", "to_k", "to_v")], 0), device, f32)
        wo = sd[name + ".to_out.0.weight"]
        self.wo = _dev(wo.reshape(wo.shape[0], wo.shape[1]).to(torch.bfloat16), device)
        self.bo = _dev(sd[name + ".to_out.0.bias"], device, f32)
        self.c = wq.shape[0]

    def gen(self, x):
        B, H, W, C = x.shape
        T = H * W
        h = yield (x, self.g, self.b, False, None)
        qkv = capi.gemm(h.view(B * T, C), self.wqkv, self.bqkv)  # [B*T, 3C]
        q, k, v = (qkv[:, i * C:(i + 1) * C].unflatten(0, (B, T)).unsqueeze(2) for i in range(3))  # [B,T,1,C]
        a = capi.attention_materialised(q, k, v, 1.0 / math.sqrt(C))  # [B,T,1,C]
        if not x.is_contiguous():
            raise RuntimeError("VAE mid attention expects a dense NHWC input")
        out = capi.gemm(a.view(B * T, C), self.wo, self.bo, res=x.view(B * T, C))
        return out.view(B, H, W, C)


class B200Encoder(torch.nn.Module):
    """diffusers `Encoder`: [B,3,H,W] -> [B,2*Cz,H/8,W/8] (pre-quant_conv), NCHW in / NCHW out."""

    def __init__(self, sd, cfg, device, dtype):
        super().__init__()
        self.cfg, self.dev, self.out_dtype = cfg, device, dtype
        f32 = torch.float32
        ch = cfg.block_out_channels
        self.conv_in_w = _dev(pack_conv_in_weight(sd["encoder.conv_in.weight"]), device)
        self.conv_in_b = _dev(pad8(sd["encoder.conv_in.bias"]), device, f32)
        self.conv_in_wg = _dev(pack_conv_in_gemm_weight(sd["encoder.conv_in.weight"]), device)
        self.conv_in_bg = _dev(sd["encoder.conv_in.bias"], device, f32)
        self.levels = []
        for i in range(len(ch)):
            res = [_Res(sd, f"encoder.down_blocks.{i}.resnets.{j}", device) for j in range(cfg.layers_per_block)]
            down = None
            if i < len(ch) - 1:
                n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                down = (_dev(pack_conv_weight(sd[n + ".weight"]), device), _dev(sd[n + ".bias"], device, f32))
            self.levels.append((res, down))
        self.mid0 = _Res(sd, "encoder.mid_block.resnets.0", device)
        self.attn = _MidAttn(sd, "encoder.mid_block.attentions.0", device)
        self.mid1 = _Res(sd, "encoder.mid_block.resnets.1", device)
        self.gno = _dev(sd["encoder.conv_norm_out.weight"], device, f32)
        self.bno = _dev(sd["encoder.conv_norm_out.bias"], device, f32)
        self.set_conv_out(sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"])
        self._p = torch.nn.Parameter(torch.zeros(1, device=device, dtype=dtype), requires_grad=False)

    def set_conv_out(self, w, b):
        self.cout = w.shape[0]
        self.conv_out_w = _dev(pack_conv_weight(w), self.dev)
        self.conv_out_b = _dev(b, self.dev, torch.float32)

    def trunk_gen(self, x):
        """NCHW image -> NHWC bf16 features after conv_norm_out + SiLU (input of conv_out). GroupNorm statistics
        travel with the tensors: every conv whose output feeds a GroupNorm produces them in its epilogue."""
        # 3 -> C conv_in at full image resolution runs as im2col (K = 27 -> 64) + tensor-core GEMM
        # (the CUDA-core kernel took 4.7 ms of a 105 ms step at 16 x 512^2)
        h = thin_conv_in(x, self.conv_in_w, self.conv_in_b, self.conv_in_wg, self.conv_in_bg,
                         self.cfg.block_out_channels[0])
        st = None
        for res, down in self.levels:
            for r in res:
                h, st = yield from resnet_gen(r, h, x_stats=st, want_stats=True)
            if down is not None:  # F.pad(x,(0,1,0,1)) + stride-2 conv
                h, st = conv_with_stats(h, down[0], down[1], True, stride=2, pad=(0, 0, 1, 1))
        h, _ = yield from resnet_gen(self.mid0, h, x_stats=st)
        h = yield from self.attn.gen(h)
        h, st = yield from resnet_gen(self.mid1, h, want_stats=True)
        h = yield (h, self.gno, self.bno, True, st)
        return h

    def moments_nhwc(self, x, w=None, b=None):
        """fp32 NHWC conv_out result (optionally with quant_conv folded into w/b)."""
        h = drive_plain(self.trunk_gen(x), self.cfg.eps)
        return capi.conv2d(h, self.conv_out_w if w is None else w, self.conv_out_b if b is None else b,
                           out_f32=True, cout=self.cout)

    def forward_gen(self, x):
        """coroutine form of forward for the tiled driver: NCHW tile batch -> NCHW fp32 moments."""
        h = yield from self.trunk_gen(x.to(self.out_dtype))
        m = capi.conv2d(h, self.conv_out_w, self.conv_out_b, out_f32=True, cout=self.cout)
        return m.permute(0, 3, 1, 2).contiguous()

    def run_tiled(self, gens, tiles_per_group, n_img, weight=1.0, tile_parallel=False):
        return drive_tiled(gens, self.cfg.eps, n_img, weight, tile_parallel)

    def run_tiled_fast(self, gens, n_img, estimate_input):
        """VAEHook fast mode: statistics from one pass over `estimate_input`, replayed by every tile group."""
        sites = drive_record(self.forward_gen(estimate_input), self.cfg.eps)
        return [drive_fixed(g, self.cfg.eps, n_img, sites) for g in gens]

    def forward(self, x):
        m = self.moments_nhwc(x)  # [B,h,w,2Cz] fp32
        return m.permute(0, 3, 1, 2).contiguous().to(self.out_dtype)


class B200Decoder(torch.nn.Module):
    """diffusers `Decoder`: [B,Cz,h,w] -> [B,3,8h,8w], NCHW in / NCHW out."""

    def __init__(self, sd, cfg, device, dtype):
        super().__init__()
        self.cfg, self.dev, self.out_dtype = cfg, device, dtype
        f32 = torch.float32
        ch = list(reversed(cfg.block_out_channels))
        self.conv_in_w = _dev(pack_conv_in_weight(sd["decoder.conv_in.weight"]), device)
        self.conv_in_b = _dev(pad8(sd["decoder.conv_in.bias"]), device, f32)
        self.conv_in_wg = _dev(pack_conv_in_gemm_weight(sd["decoder.conv_in.weight"]), device)
        self.conv_in_bg = _dev(sd["decoder.conv_in.bias"], device, f32)
        self.mid0 = _Res(sd, "decoder.mid_block.resnets.0", device)
        self.attn = _MidAttn(sd, "decoder.mid_block.attentions.0", device)
        self.mid1 = _Res(sd, "decoder.mid_block.resnets.1", device)
        self.levels = []
        for i in range(len(ch)):
            res = [_Res(sd, f"decoder.up_blocks.{i}.resnets.{j}", device) for j in range(cfg.layers_per_block + 1)]
            up = None
            if i < len(ch) - 1:
                n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                up = (_dev(pack_upsample_weight(sd[n + ".weight"]), device), _dev(sd[n + ".bias"], device, f32))
            self.levels.append((res, up))
        self.gno = _dev(sd["decoder.conv_norm_out.weight"], device, f32)
        self.bno = _dev(sd["decoder.conv_norm_out.bias"], device, f32)
        self.conv_out_w = _dev(pack_conv_weight(sd["decoder.conv_out.weight"]), device)
        self.conv_out_b = _dev(sd["decoder.conv_out.bias"], device, f32)
        self._p = torch.nn.Parameter(torch.zeros(1, device=device, dtype=dtype), requires_grad=False)

    def trunk_gen(self, z):
        ch0 = self.cfg.block_out_channels[-1]
        h = thin_conv_in(z, self.conv_in_w, self.conv_in_b, self.conv_in_wg, self.conv_in_bg, ch0)
        h, _ = yield from resnet_gen(self.mid0, h)
        h = yield from self.attn.gen(h)
        h, st = yield from resnet_gen(self.mid1, h, want_stats=True)
        for res, up in self.levels:
            for r in res:
                h, st = yield from resnet_gen(r, h, x_stats=st, want_stats=True)
            if up is not None:  # nearest-2x folded into 4 parity convs; all four accumulate into one statistics buffer
                h, st = conv_with_stats(h, up[0], up[1], True, upsample2x=True)
        h = yield (h, self.gno, self.bno, True, st)
        return h

    def _conv_out(self, h, clamp, out_dtype):
        """decoder.conv_out (C -> 3). Large images: the tensor-core conv with the three output channels in one
        masked 64-wide tile (reads the 1 GB input once through TMA; the CUDA-core kernel was LSU-bound at
        2.9 ms for 16 x 512^2), then a 3-channel NHWC->NCHW (+clamp). Small tiles: the direct kernel."""
        B, H, W, C = h.shape
        if B * H * W >= (1 << 18) and C % 64 == 0:
            y = capi.conv2d(h, self.conv_out_w, self.conv_out_b, cout=self.cfg.out_channels)
            return capi.nhwc_to_nchw(y, out_dtype, clamp=clamp)
        return capi.conv_out_nchw(h, self.conv_out_w, self.conv_out_b, self.cfg.out_channels, clamp=clamp,
                                  out_dtype=out_dtype)

    def forward_gen(self, z):
        """coroutine form of forward for the tiled driver: NCHW latent tile batch -> NCHW fp32 image tiles."""
        h = yield from self.trunk_gen(z.to(self.out_dtype))
        return self._conv_out(h, False, torch.float32)

    def run_tiled(self, gens, tiles_per_group, n_img, weight=1.0, tile_parallel=False):
        return drive_tiled(gens, self.cfg.eps, n_img, weight, tile_parallel)

    def run_tiled_fast(self, gens, n_img, estimate_input):
        """VAEHook fast mode: statistics from one pass over `estimate_input`, replayed by every tile group."""
        sites = drive_record(self.forward_gen(estimate_input), self.cfg.eps)
        return [drive_fixed(g, self.cfg.eps, n_img, sites) for g in gens]

    def forward(self, z, clamp=False):
        h = drive_plain(self.trunk_gen(z), self.cfg.eps)
        return self._conv_out(h, clamp, self.out_dtype)


class _Conv1x1(torch.nn.Module):
    """quant_conv / post_quant_conv on tiny NCHW latents (thin conv kernel -> NHWC -> NCHW)."""

    def __init__(self, w, b, device, dtype):
        super().__init__()
        self.w = _dev(pack_conv_in_weight(w), device)
        self.b = _dev(pad8(b), device, torch.float32)
        self.bf = _dev(b, device, torch.float32)
        self.wf = _dev(w.reshape(w.shape[0], w.shape[1]), device, torch.float32)
        self.cout, self.out_dtype = w.shape[0], dtype

    def forward(self, x):
        y = capi.conv_in_nchw(x.contiguous(), self.w, self.b, self.cout, ksize=1)
        return capi.nhwc_to_nchw(y, self.out_dtype)


class _Posterior:
    """DiagonalGaussianDistribution stand-in: `.sample()` draws eps from torch's global CUDA generator like
    diffusers' randn_tensor; tests inject a fixed eps through `B200AutoencoderKL.posterior_eps`."""

    def __init__(self, moments_nhwc_f32, cz, out_dtype, eps=None):
        self.m, self.cz, self.out_dtype, self.eps = moments_nhwc_f32, cz, out_dtype, eps

    def sample(self, generator=None, shift=0.0, scale=1.0):
        """mean + std * eps; `shift` / `scale` fold the caller's (z - shift_factor) * scaling_factor into the kernel."""
        B, h, w, _ = self.m.shape
        eps = self.eps
        if eps is None:
            from .dist import broadcast_
            eps = broadcast_(torch.randn((B, self.cz, h, w), device=self.m.device, dtype=torch.float32,
                                         generator=generator))
        return capi.posterior_sample(self.m, eps.to(device=self.m.device, dtype=torch.float32), self.cz, shift, scale,
                                     self.out_dtype)

    def mode(self):
        return self.m[..., :self.cz].permute(0, 3, 1, 2).contiguous().to(self.out_dtype)


class B200AutoencoderKL(torch.nn.Module):
    def __init__(self, sd, cfg, device="cuda", dtype=torch.bfloat16):
        super().__init__()
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("B200AutoencoderKL: boundary dtype must be bf16 or fp32 (compute is bf16/fp32-acc)")
        self.config = SimpleNamespace(scaling_factor=cfg.scaling_factor, shift_factor=cfg.shift_factor,
                                      block_out_channels=cfg.block_out_channels,
                                      latent_channels=cfg.latent_channels)
        self.cfg = cfg
        self._dtype = dtype
        self.encoder = B200Encoder(sd, cfg, device, dtype)
        self.decoder = B200Decoder(sd, cfg, device, dtype)
        self.quant_conv = self.post_quant_conv = None
        self._folded = None
        if cfg.use_quant_conv:
            self.quant_conv = _Conv1x1(sd["quant_conv.weight"], sd["quant_conv.bias"], device, dtype)
            self.post_quant_conv = _Conv1x1(sd["post_quant_conv.weight"], sd["post_quant_conv.bias"], device, dtype)
            # quant_conv (1x1) after conv_out (3x3) is an exact linear composition: fold it into conv_out
            wq = sd["quant_conv.weight"].float().reshape(2 * cfg.latent_channels, 2 * cfg.latent_channels)
            wc = sd["encoder.conv_out.weight"].float()
            wf = torch.einsum("om,mikl->oikl", wq, wc)
            bf = wq @ sd["encoder.conv_out.bias"].float() + sd["quant_conv.bias"].float()
            self._folded = (_dev(pack_conv_weight(wf), device), _dev(bf, device, torch.float32))
        self.posterior_eps = None  # test hook: fixed noise for latent_dist.sample()

    @property
    def dtype(self):
        return self._dtype

    def encode(self, x):
        x = x.to(self._dtype)
        enc_patched = "forward" in self.encoder.__dict__  # vaehook monkey-patches the instance attribute
        if enc_patched:
            h = self.encoder.forward(x)  # tiled path returns NCHW pre-quant moments
            if self.quant_conv is not None:
                h = self.quant_conv(h.to(self._dtype))
            m = h.permute(0, 2, 3, 1).contiguous().float()
        elif self._folded is not None:
            m = self.encoder.moments_nhwc(x, *self._folded)
        else:
            m = self.encoder.moments_nhwc(x)
        return SimpleNamespace(latent_dist=_Posterior(m, self.cfg.latent_channels, self._dtype, self.posterior_eps))

    def encode_scaled(self, x):
        """(encode(x).latent_dist.sample() - shift_factor) * scaling_factor in one posterior kernel
        (infer/omgsr_f_infer_model.py:15-18 / infer/omgsr_s_infer_model.py:173)."""
        return self.encode(x).latent_dist.sample(shift=self.config.shift_factor, scale=self.config.scaling_factor)

    def decode(self, z, return_dict=True):
        z = z.to(self._dtype)
        if self.post_quant_conv is not None:
            z = self.post_quant_conv(z)
        img = self.decoder.forward(z)
        if not return_dict:
            return (img,)
        return SimpleNamespace(sample=img)
