"""ctypes binding of libomgsr_b200.so (the C ABI declared in include/omgsr_b200.h).

PyTorch is plumbing only: tensors provide device memory (`data_ptr()`), the current CUDA stream and
shapes. Every wrapper raises RuntimeError with the library's message on a non-zero return code; there is
no CPU fallback — if the shared library is missing, import of this module fails loudly.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libomgsr_b200.so")

ACT_NONE, ACT_SILU, ACT_GELU_ERF, ACT_GELU_TANH, ACT_GEGLU = 0, 1, 2, 3, 4
OUT_F32, B_MN_MAJOR, CONV_NO_HALO, RES_F32 = 1, 2, 4, 8

# every symbol include/omgsr_b200.h declares (tests/test_abi.py checks the two lists agree)
SYMBOLS = [
    "omgsr_version", "omgsr_last_error", "omgsr_destroy", "omgsr_launch_count", "omgsr_last_variant",
    "omgsr_gn_scratch_floats",
    "omgsr_gemm_bf16", "omgsr_gemm_batched_bf16", "omgsr_gemm_rope_bf16", "omgsr_bgemm_bf16", "omgsr_conv2d_nhwc_bf16",
    "omgsr_conv_in_nchw", "omgsr_conv_out_nchw", "omgsr_im2col3x3_nchw",
    "omgsr_groupnorm_stats", "omgsr_groupnorm_apply", "omgsr_groupnorm_combine", "omgsr_layernorm", "omgsr_softmax_rows",
    "omgsr_rmsnorm_rope", "omgsr_posterior_sample", "omgsr_latent_axpby", "omgsr_x0_unscale",
    "omgsr_pack_latents", "omgsr_unpack_latents", "omgsr_tile_blend",
    "omgsr_wavelet_fix", "omgsr_adain_fix", "omgsr_atrous_blur", "omgsr_nchw_to_nhwc", "omgsr_nhwc_to_nchw",
    "omgsr_flash_attn_bf16", "omgsr_flash_attn_set_variant", "omgsr_set_gemm_2cta", "omgsr_set_halo_rolling", "omgsr_conv_gn_wide_fusable", "omgsr_softmax_rows_causal", "omgsr_embed_tokens", "omgsr_softmax_rows_bias", "omgsr_rmsnorm", "omgsr_mul_bf16",
    "omgsr_groupnorm_coef", "omgsr_conv2d_gn_nhwc_bf16",
    "omgsr_resample_coeffs", "omgsr_resize_u8_workspace", "omgsr_resize_u8", "omgsr_u8_to_input", "omgsr_output_to_u8",
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m omgsr_b200.build` (or __graft_entry__.build()). "
            "omgsr_b200 has no CPU / PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.omgsr_last_error.restype = ctypes.c_char_p
    lib.omgsr_launch_count.restype = ctypes.c_longlong
    lib.omgsr_resize_u8_workspace.restype = ctypes.c_longlong
    lib.omgsr_gn_scratch_floats.restype = ctypes.c_longlong
    lib.omgsr_last_variant.restype = ctypes.c_char_p
    return lib


_lib = _load()

_vp, _i, _ll, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def lib():
    return _lib


def _ptr(t):
    if t is None:
        return _vp(0)
    return _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _check(rc, name):
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {_lib.omgsr_last_error().decode()}")


def version():
    return _lib.omgsr_version()


def launch_count():
    return int(_lib.omgsr_launch_count())


def last_conv_variant():
    """name of the kernel variant that served the most recent GEMM / convolution call (diagnostics)."""
    return (_lib.omgsr_last_variant() or b"").decode()


def _gn_scratch(B, groups, launches, device):
    """workspace for the deterministic (atomic-free) GroupNorm statistics: per-warp / per-block partial sums."""
    n = int(_lib.omgsr_gn_scratch_floats(_i(B), _i(groups), _i(launches)))
    return torch.empty(n, device=device, dtype=torch.float32), n


def _cuda_bf16(t, name):
    if not (t.is_cuda and t.dtype == torch.bfloat16):
        raise TypeError(f"{name} must be a CUDA bf16 tensor (got {t.device}, {t.dtype})")


def _cuda_f32(t, name):
    if t is not None and not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{name} must be a CUDA fp32 tensor (got {t.device}, {t.dtype})")


# ------------------------------------------------------------------------------------------------
# contractions
# ------------------------------------------------------------------------------------------------
def gemm(a, w, bias=None, res=None, gate=None, act=ACT_NONE, gate_rows=0, out=None, out_f32=False, force_bn=0,
         n_out=None):
    """out[M, N] = epi(a[M, K] @ w[N, K]^T). a may be a row-strided 2-D view (last dim contiguous)."""
    _cuda_bf16(a, "a"); _cuda_bf16(w, "w"); _cuda_f32(bias, "bias"); _cuda_f32(gate, "gate")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_cols = N // 2 if act == ACT_GEGLU else N
    if out is None:
        ld = (n_cols + 7) // 8 * 8
        out = torch.empty((M, ld), device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
        if ld != n_cols:
            out = out[:, :n_cols]
    assert out.stride(1) == 1
    ldr = res.stride(0) if res is not None else 0
    rc = _lib.omgsr_gemm_bf16(_ptr(a), _ptr(w), _ptr(bias), _ptr(res), _ptr(gate), _ptr(out), _ll(M), _i(N), _i(K),
                              _ll(a.stride(0)), _ll(out.stride(0)), _ll(ldr), _i(act), _i(gate_rows),
                              _i(OUT_F32 if out.dtype == torch.float32 else 0), _i(force_bn), _stream())
    _check(rc, "omgsr_gemm_bf16")
    return out


def bgemm_qk(q, k, out=None, force_bn=0):
    """scores[b,h,m,n] = sum_d q[b,m,h,d] k[b,n,h,d] (fp32). q/k: [B, T, H, D] views with D contiguous."""
    _cuda_bf16(q, "q"); _cuda_bf16(k, "k")
    B, M, H, D = q.shape
    N = k.shape[1]
    n_pad = (N + 63) // 64 * 64
    if out is None:
        out = torch.empty((B, H, M, n_pad), device=q.device, dtype=torch.float32)
    rc = _lib.omgsr_bgemm_bf16(_ptr(q), _ptr(k), _ptr(out), _i(B), _i(H), _i(M), _i(N), _i(D),
                               _ll(q.stride(1)), _ll(q.stride(2)), _ll(q.stride(0)),
                               _ll(k.stride(1)), _ll(k.stride(2)), _ll(k.stride(0)),
                               _ll(out.stride(2)), _ll(out.stride(1)), _ll(out.stride(0)), _i(N),
                               _i(OUT_F32), _i(force_bn), _stream())
    _check(rc, "omgsr_bgemm_bf16(qk)")
    return out


def bgemm_pv(p, v, out=None, n_kv=None, force_bn=0):
    """out[b,m,h,d] = sum_n p[b,h,m,n] v[b,n,h,d]. p: [B,H,M,Npad] bf16 (zero beyond n_kv); v: [B,N,H,D]."""
    _cuda_bf16(p, "p"); _cuda_bf16(v, "v")
    B, H, M, Np = p.shape
    D = v.shape[3]
    if out is None:
        out = torch.empty((B, M, H, D), device=p.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_bgemm_bf16(_ptr(p), _ptr(v), _ptr(out), _i(B), _i(H), _i(M), _i(D), _i(Np),
                               _ll(p.stride(2)), _ll(p.stride(1)), _ll(p.stride(0)),
                               _ll(v.stride(1)), _ll(v.stride(2)), _ll(v.stride(0)),
                               _ll(out.stride(1)), _ll(out.stride(2)), _ll(out.stride(0)), _i(v.shape[1]),
                               _i(B_MN_MAJOR), _i(force_bn), _stream())
    _check(rc, "omgsr_bgemm_bf16(pv)")
    return out


def conv2d(x, w, bias=None, res=None, ksize=3, stride=1, pad=(1, 1, 1, 1), upsample2x=False, act=ACT_NONE, out=None,
           out_f32=False, cout=None, force_bn=0, flags=0, gn_stats=None, gn_groups=32):
    """NHWC conv. x: [B,H,W,Cin] bf16 view (channel-contiguous, dense in B/H/W with a channel stride);
    w: packed [Cout, k*k*Cin] (or [4, Cout, 4*Cin] for upsample2x). pad = (top, left, bottom, right)."""
    _cuda_bf16(x, "x"); _cuda_bf16(w, "w"); _cuda_f32(bias, "bias")
    B, H, W, Cin = x.shape
    assert x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1)
    Cout = cout if cout is not None else (w.shape[1] if upsample2x else w.shape[0])
    if upsample2x:
        OH, OW = 2 * H, 2 * W
    else:
        OH = (H + pad[0] + pad[2] - ksize) // stride + 1
        OW = (W + pad[1] + pad[3] - ksize) // stride + 1
    if out is None:
        ld = (Cout + 7) // 8 * 8
        out = torch.empty((B, OH, OW, ld), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
        if ld != Cout:
            out = out[..., :Cout]
    assert out.shape[:3] == (B, OH, OW) and out.stride(3) == 1
    assert out.stride(1) == OW * out.stride(2) and out.stride(0) == OH * out.stride(1)
    ldr = 0
    if res is not None:
        _cuda_bf16(res, "res")
        assert res.shape[:3] == (B, OH, OW) and res.stride(3) == 1
        ldr = res.stride(2)
    scratch, n_scr = _gn_scratch(B, gn_groups, 4 if upsample2x else 1, x.device) if gn_stats is not None else (None, 0)
    rc = _lib.omgsr_conv2d_nhwc_bf16(_ptr(x), _ptr(w), _ptr(bias), _ptr(res), _ptr(out), _i(B), _i(H), _i(W),
                                     _i(Cin), _i(Cout), _i(ksize), _i(stride), _i(pad[0]), _i(pad[1]), _i(pad[2]),
                                     _i(pad[3]), _i(1 if upsample2x else 0), _i(act), _ll(x.stride(2)),
                                     _ll(out.stride(2)), _ll(ldr), _ptr(gn_stats),
                                     _i(gn_groups if gn_stats is not None else 0), _ptr(scratch), _ll(n_scr),
                                     _i((OUT_F32 if out.dtype == torch.float32 else 0) | flags), _i(force_bn),
                                     _stream())
    _check(rc, "omgsr_conv2d_nhwc_bf16")
    return out


def conv_in_nchw(x, w, bias, cout, ksize=3, out=None):
    """x NCHW (bf16|fp32) -> NHWC bf16 [B,H,W,cout]. w fp32 [k*k*Cin, cout8], bias fp32 [cout8] (zero-padded to a
    multiple of 8, omgsr_b200.pack.pack_conv_in_weight / pad8)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.bfloat16, torch.float32)
    _cuda_f32(w, "w"); _cuda_f32(bias, "bias")
    B, Cin, H, W = x.shape
    cout8 = w.shape[1]
    assert cout8 % 8 == 0 and cout8 >= cout and bias.numel() == cout8
    if out is None:
        out = torch.empty((B, H, W, cout8), device=x.device, dtype=torch.bfloat16)[..., :cout]
    assert out.stride(2) >= cout8 or cout8 == cout
    rc = _lib.omgsr_conv_in_nchw(_ptr(x), _i(int(x.dtype == torch.float32)), _ptr(w), _ptr(bias), _ptr(out), _i(B),
                                 _i(Cin), _i(H), _i(W), _i(cout8), _i(ksize), _ll(out.stride(2)), _stream())
    _check(rc, "omgsr_conv_in_nchw")
    return out


def conv_in_gemm(x, w_gemm, bias, out=None):
    """thin-input 3x3/p1 conv on the tensor cores: im2col (k padded to a multiple of 64) + GEMM.
    x NCHW (bf16|fp32); w_gemm bf16 [Cout, Kpad] (omgsr_b200.pack.pack_conv_in_gemm_weight); bias fp32 [Cout].
    -> NHWC bf16 [B,H,W,Cout] (optionally written into the NHWC view `out`, which may be a channel slice)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.bfloat16, torch.float32)
    _cuda_bf16(w_gemm, "w_gemm"); _cuda_f32(bias, "bias")
    B, Cin, H, W = x.shape
    cout, kpad = w_gemm.shape
    patches = torch.empty((B * H * W, kpad), device=x.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_im2col3x3_nchw(_ptr(x), _i(int(x.dtype == torch.float32)), _ptr(patches), _i(B), _i(Cin), _i(H),
                                   _i(W), _i(kpad), _stream())
    _check(rc, "omgsr_im2col3x3_nchw")
    if out is None:
        out = torch.empty((B, H, W, cout), device=x.device, dtype=torch.bfloat16)
    o2 = torch.as_strided(out, (B * H * W, cout), (out.stride(2), 1), out.storage_offset())
    gemm(patches, w_gemm, bias, out=o2)
    return out


def conv_out_nchw(x, w, bias, cout, clamp=False, out_dtype=torch.bfloat16):
    """x NHWC bf16 -> NCHW (bf16|fp32), 3x3 p1, Cout in {3,4}. w bf16 [Cout, 9*Cin]."""
    _cuda_bf16(x, "x"); _cuda_bf16(w, "w"); _cuda_f32(bias, "bias")
    B, H, W, Cin = x.shape
    out = torch.empty((B, cout, H, W), device=x.device, dtype=out_dtype)
    rc = _lib.omgsr_conv_out_nchw(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), _i(int(out_dtype == torch.float32)), _i(B),
                                  _i(H), _i(W), _i(Cin), _i(cout), _i(int(clamp)), _ll(x.stride(2)), _stream())
    _check(rc, "omgsr_conv_out_nchw")
    return out


# ------------------------------------------------------------------------------------------------
# glue
# ------------------------------------------------------------------------------------------------
def gn_fusable(B, H, W, C, groups=32):
    """can the conv producing an NHWC [B,H,W,C] tensor also produce its GroupNorm statistics (conv2d gn_stats=)?"""
    # cpg = 4 (C = 128, the 512^2 levels) is fused too since the epilogue was specialised: the statistics ride in the
    # epilogue's slack behind the K = 1152 MMAs and save a 1 GB re-read per layer (S-512: 86.9 -> 85.4 ms per step).
    cpg = C // groups
    return C % groups == 0 and H * W >= 128 and cpg in (4, 8, 16)


def groupnorm_stats(x, groups=32, stats=None):
    """(sum, sumsq) per (batch entry, group) of an NHWC bf16 view -> fp32 [B, groups, 2]."""
    _cuda_bf16(x, "x")
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    if stats is None:
        stats = torch.empty((B, groups, 2), device=x.device, dtype=torch.float32)
    scratch, n_scr = _gn_scratch(B, groups, 1, x.device)
    rc = _lib.omgsr_groupnorm_stats(_ptr(x), _ll(x.stride(-2)), _ptr(stats), _i(B), _ll(HW), _i(C), _i(groups),
                                    _ptr(scratch), _ll(n_scr), _stream())
    _check(rc, "omgsr_groupnorm_stats")
    return stats


def groupnorm_apply(x, stats, gamma, beta, groups=32, eps=1e-6, silu=False, out=None, shared_mod=0):
    """normalise with given statistics. shared_mod > 0: stats are (mean, var) rows shared by batch % shared_mod."""
    _cuda_bf16(x, "x"); _cuda_f32(gamma, "gamma"); _cuda_f32(beta, "beta"); _cuda_f32(stats, "stats")
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    coef = torch.empty((B, C, 2), device=x.device, dtype=torch.float32)
    rc = _lib.omgsr_groupnorm_apply(_ptr(x), _ll(x.stride(-2)), _ptr(out), _ll(out.stride(-2)), _ptr(stats),
                                    _ptr(gamma), _ptr(beta), _ptr(coef), _i(B), _ll(HW), _i(C), _i(groups), _f(eps),
                                    _i(int(silu)), _i(1 if shared_mod > 0 else 0), _i(shared_mod), _stream())
    _check(rc, "omgsr_groupnorm_apply")
    return out


def groupnorm(x, gamma, beta, groups=32, eps=1e-6, silu=False, out=None, stats=None):
    """x: [B, ..., C] NHWC bf16 (channel stride = x.stride(-2)). Two launches: stats, apply(+SiLU)."""
    stats = groupnorm_stats(x, groups, stats)
    return groupnorm_apply(x, stats, gamma, beta, groups, eps, silu, out)


def groupnorm_coef(stats, gamma, beta, B, HW, groups=32, eps=1e-6, shared_mod=0):
    """per-(image, channel) affine coefficients (rstd*gamma, beta - mean*rstd*gamma): fp32 [B, C, 2]."""
    _cuda_f32(stats, "stats"); _cuda_f32(gamma, "gamma"); _cuda_f32(beta, "beta")
    C = gamma.numel()
    coef = torch.empty((B, C, 2), device=stats.device, dtype=torch.float32)
    rc = _lib.omgsr_groupnorm_coef(_ptr(stats), _ptr(gamma), _ptr(beta), _ptr(coef), _i(B), _ll(HW), _i(C), _i(groups),
                                   _f(eps), _i(1 if shared_mod else 0), _i(max(shared_mod, 1)), _stream())
    _check(rc, "omgsr_groupnorm_coef")
    return coef


def conv_gn_in_fusable(B, H, W, Cin, Cout):
    """can omgsr_conv2d_gn_nhwc_bf16 serve this 3x3/s1/p1 conv (CTA-pair input-halo path)?"""
    if os.environ.get("OMGSR_NO_GN_IN_FUSION"):
        return False
    if Cout == 256:  # two 128-channel passes of the rolling-row kernel (large images only: the library decides)
        if os.environ.get("OMGSR_NO_WIDE"):   # A/B timing
            return False
        return bool(_lib.omgsr_conv_gn_wide_fusable(_i(B), _i(H), _i(W), _i(Cin), _i(Cout)))
    tiles = ((W + 127) // 128) * ((H + 1) // 2) * B
    return W >= 96 and Cin % 64 == 0 and 16 < Cout <= 128 and tiles >= 2


def conv2d_gn(x, coef, w, bias, res=None, out=None, gn_stats=None, gn_groups=32):
    """conv3x3(silu(groupnorm(x))) with the normalisation applied to the operand tile inside the kernel.
    x NHWC bf16 view (raw, un-normalised), coef fp32 [B, Cin, 2] from groupnorm_coef."""
    _cuda_bf16(x, "x"); _cuda_bf16(w, "w"); _cuda_f32(bias, "bias"); _cuda_f32(coef, "coef")
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1)
    assert coef.is_contiguous() and coef.shape == (B, Cin, 2)
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.bfloat16)
    assert out.stride(3) == 1 and out.stride(1) == W * out.stride(2) and out.stride(0) == H * out.stride(1)
    ldr = 0
    if res is not None:
        _cuda_bf16(res, "res")
        assert res.shape == out.shape and res.stride(3) == 1 and res.stride(1) == W * res.stride(2)
        ldr = res.stride(2)
    scratch, n_scr = _gn_scratch(B, gn_groups, 1, x.device) if gn_stats is not None else (None, 0)
    rc = _lib.omgsr_conv2d_gn_nhwc_bf16(_ptr(x), _ptr(coef), _ptr(w), _ptr(bias), _ptr(res), _ptr(out), _i(B), _i(H),
                                        _i(W), _i(Cin), _i(Cout), _ll(x.stride(2)), _ll(out.stride(2)), _ll(ldr),
                                        _ptr(gn_stats), _i(gn_groups), _ptr(scratch), _ll(n_scr), _stream())
    _check(rc, "omgsr_conv2d_gn_nhwc_bf16")
    return out


def groupnorm_combine(sums, pixels, n_img, C, groups=32, raw=False):
    """sums fp32 [entries, groups, 2], pixels fp32 [entries] -> (mean, var) fp32 [n_img, groups, 2].
    raw=True -> (numerators [n_img, groups, 2], weight sums [n_img]) for a cross-GPU all-reduce."""
    _cuda_f32(sums, "sums"); _cuda_f32(pixels, "pixels")
    entries = sums.shape[0]
    sums = sums.contiguous()
    out = torch.empty((n_img, groups, 2), device=sums.device, dtype=torch.float32)
    wsum = torch.empty((n_img,), device=sums.device, dtype=torch.float32) if raw else None
    rc = _lib.omgsr_groupnorm_combine(_ptr(sums), _ptr(pixels), _i(entries), _i(n_img), _i(C), _i(groups), _ptr(out),
                                      _ptr(wsum), _stream())
    _check(rc, "omgsr_groupnorm_combine")
    return (out, wsum) if raw else out


def layernorm(x, gamma=None, beta=None, eps=1e-5, rows_per_batch=0, gb_stride=0, out=None):
    """x: dense [..., C] or a 3-D view [B, rows, C] with free row / batch strides (bf16, or fp32 for an fp32 residual
    stream); output dense bf16."""
    if not (x.is_cuda and x.dtype in (torch.bfloat16, torch.float32)):
        raise TypeError(f"x must be a CUDA bf16 / fp32 tensor (got {x.device}, {x.dtype})")
    _cuda_f32(gamma, "gamma"); _cuda_f32(beta, "beta")
    C = x.shape[-1]
    rows = x.numel() // C
    ldx, xbs = C, 0
    if not x.is_contiguous():
        assert x.dim() == 3 and x.stride(2) == 1, "layernorm: strided input must be a [B, rows, C] view"
        ldx, xbs = x.stride(1), x.stride(0)
        if rows_per_batch == 0:
            rows_per_batch = x.shape[1]
        assert rows_per_batch == x.shape[1]
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_layernorm(_ptr(x), _i(int(x.dtype == torch.float32)), _ptr(out), _ptr(gamma), _ptr(beta), _ll(rows),
                              _i(C), _f(eps),
                              _ll(rows_per_batch), _ll(gb_stride), _ll(ldx), _ll(xbs), _stream())
    _check(rc, "omgsr_layernorm")
    return out


def softmax_rows(s, n, scale, out=None, causal=False):
    """s: fp32 [..., M, n_pad]; returns bf16 probabilities with columns >= n zeroed. causal: query i (the second to
    last axis) sees keys 0..i only."""
    _cuda_f32(s, "s")
    n_pad = s.shape[-1]
    rows = s.numel() // n_pad
    assert s.is_contiguous()
    if out is None:
        out = torch.empty(s.shape, device=s.device, dtype=torch.bfloat16)
    if causal:
        rc = _lib.omgsr_softmax_rows_causal(_ptr(s), _ll(n_pad), _ptr(out), _ll(n_pad), _ll(rows), _i(n), _i(n_pad),
                                            _f(scale), _i(s.shape[-2]), _stream())
        _check(rc, "omgsr_softmax_rows_causal")
        return out
    rc = _lib.omgsr_softmax_rows(_ptr(s), _ll(n_pad), _ptr(out), _ll(n_pad), _ll(rows), _i(n), _i(n_pad), _f(scale),
                                 _stream())
    _check(rc, "omgsr_softmax_rows")
    return out


def softmax_rows_bias(s, n, scale, bias, out=None):
    """s: fp32 [B, H, M, n_pad]; bias: fp32 [H, M, >= n] shared by the batch -> bf16 softmax(scale * s + bias)."""
    _cuda_f32(s, "s"); _cuda_f32(bias, "bias")
    assert s.is_contiguous() and bias.is_contiguous() and s.dim() == 4 and bias.shape[:2] == s.shape[1:3]
    n_pad = s.shape[-1]
    rows = s.numel() // n_pad
    if out is None:
        out = torch.empty(s.shape, device=s.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_softmax_rows_bias(_ptr(s), _ll(n_pad), _ptr(out), _ll(n_pad), _ll(rows), _i(n), _i(n_pad), _f(scale),
                                      _ptr(bias), _ll(bias.shape[0] * bias.shape[1]), _ll(bias.shape[2]), _stream())
    _check(rc, "omgsr_softmax_rows_bias")
    return out


def rmsnorm(x, weight, eps=1e-6, out=None):
    """T5LayerNorm over the last dim of a dense bf16 / fp32 tensor -> dense bf16."""
    if not (x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and x.is_contiguous()):
        raise TypeError("x must be a contiguous CUDA bf16 / fp32 tensor")
    _cuda_f32(weight, "weight")
    C = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_rmsnorm(_ptr(x), _i(int(x.dtype == torch.float32)), _ptr(out), _ptr(weight), _ll(x.numel() // C),
                            _i(C), _f(eps), _ll(C), _stream())
    _check(rc, "omgsr_rmsnorm")
    return out


def mul(a, b, out=None):
    _cuda_bf16(a, "a"); _cuda_bf16(b, "b")
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
    if out is None:
        out = torch.empty_like(a)
    rc = _lib.omgsr_mul_bf16(_ptr(a), _ptr(b), _ptr(out), _ll(a.numel()), _stream())
    _check(rc, "omgsr_mul_bf16")
    return out


def embed_tokens(ids, tok, pos=None):
    """ids: int32 [B, T]; tok: bf16 [vocab, C]; pos: bf16 [T, C] or None -> bf16 [B, T, C] = tok[ids] + pos."""
    if not (ids.is_cuda and ids.dtype == torch.int32 and ids.is_contiguous()):
        raise TypeError("ids must be a contiguous CUDA int32 tensor")
    _cuda_bf16(tok, "tok")
    if pos is not None:
        _cuda_bf16(pos, "pos")
    assert tok.is_contiguous() and (pos is None or (pos.is_contiguous() and pos.shape[0] >= ids.shape[-1]))
    B, T = ids.shape
    C = tok.shape[1]
    out = torch.empty((B, T, C), device=ids.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_embed_tokens(_ptr(ids), _ptr(tok), _ptr(pos), _ptr(out), _ll(B * T), _i(T), _i(C), _i(tok.shape[0]),
                                 _stream())
    _check(rc, "omgsr_embed_tokens")
    return out


def rmsnorm_rope_(x, weight, cos_t, sin_t, heads, heads_per_weight, pos_offset=0, eps=1e-6):
    """in place on x: [B, rows, >= heads*128] bf16 view (last dim contiguous). weight fp32 [heads/heads_per_weight, 128]."""
    _cuda_bf16(x, "x"); _cuda_f32(weight, "weight"); _cuda_f32(cos_t, "cos"); _cuda_f32(sin_t, "sin")
    B, rows = x.shape[0], x.shape[1]
    assert x.stride(2) == 1 and cos_t.is_contiguous() and sin_t.is_contiguous() and weight.is_contiguous()
    assert cos_t.shape[0] >= pos_offset + rows
    rc = _lib.omgsr_rmsnorm_rope(_ptr(x), _ll(x.stride(1)), _ll(x.stride(0)), _ptr(weight), _ptr(cos_t), _ptr(sin_t),
                                 _ll(rows), _i(B), _i(heads), _i(heads_per_weight), _i(pos_offset), _f(eps), _stream())
    _check(rc, "omgsr_rmsnorm_rope")
    return x


def gemm_batched(a, w, bias=None, res=None, gate=None, act=ACT_NONE, act_cols=None, out=None, out_f32=False,
                 force_bn=0):
    """out[b, m, :] = epi(a[b, m, :] @ w^T). a/out/res: 3-D views [B, M, *] with the last dim contiguous and
    arbitrary row / batch strides; gate: fp32 2-D view [B, N] (row stride free)."""
    _cuda_bf16(a, "a"); _cuda_bf16(w, "w"); _cuda_f32(bias, "bias"); _cuda_f32(gate, "gate")
    B, M, K = a.shape
    N = w.shape[0]
    assert a.stride(2) == 1 and w.is_contiguous() and w.shape[1] == K
    if out is None:
        out = torch.empty((B, M, N), device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.stride(2) == 1 and out.shape[:2] == (B, M)
    ldr = rb = 0
    res_f32 = False
    if res is not None:
        if res.dtype == torch.float32:  # fp32 residual stream: out must be fp32 too
            assert res.is_cuda and out.dtype == torch.float32, "an fp32 residual needs an fp32 output"
            res_f32 = True
        else:
            _cuda_bf16(res, "res")
        assert res.stride(2) == 1
        ldr, rb = res.stride(1), res.stride(0)
    gld = 0
    if gate is not None:
        assert gate.stride(1) == 1 and gate.shape[0] == B
        gld = gate.stride(0)
    c0, c1 = act_cols if act_cols is not None else (0, 0)
    rc = _lib.omgsr_gemm_batched_bf16(_ptr(a), _ptr(w), _ptr(bias), _ptr(res), _ptr(gate), _ptr(out), _i(B), _i(M),
                                      _i(N), _i(K), _ll(a.stride(1)), _ll(a.stride(0)), _ll(out.stride(1)),
                                      _ll(out.stride(0)), _ll(ldr), _ll(rb), _ll(gld), _i(act), _i(c0), _i(c1),
                                      _i((OUT_F32 if out.dtype == torch.float32 else 0) | (RES_F32 if res_f32 else 0)),
                                      _i(force_bn), _stream())
    _check(rc, "omgsr_gemm_batched_bf16")
    return out


def gemm_rope(a, w, bias, wq, wk, q_cols, k_cols, cos_t, sin_t, pos_offset=0, eps=1e-6, act=ACT_NONE, act_cols=None,
              out=None, force_bn=0):
    """out[b, m, :] = a[b, m, :] @ w^T + bias with per-head RMSNorm (weights wq / wk, fp32 [128]) + RoPE applied to the
    head columns q_cols = (q0, q1) and k_cols = (k0, k1) in the GEMM epilogue; position of row m = pos_offset + m."""
    _cuda_bf16(a, "a"); _cuda_bf16(w, "w"); _cuda_f32(bias, "bias")
    for t, nm in ((wq, "wq"), (wk, "wk"), (cos_t, "cos"), (sin_t, "sin")):
        _cuda_f32(t, nm)
        assert t.is_contiguous()
    B, M, K = a.shape
    N = w.shape[0]
    assert a.stride(2) == 1 and w.is_contiguous() and w.shape[1] == K and cos_t.shape[0] >= pos_offset + M
    if out is None:
        out = torch.empty((B, M, N), device=a.device, dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16 and out.stride(2) == 1 and out.shape[:2] == (B, M)
    c0, c1 = act_cols if act_cols is not None else (0, 0)
    rc = _lib.omgsr_gemm_rope_bf16(_ptr(a), _ptr(w), _ptr(bias), _ptr(out), _i(B), _i(M), _i(N), _i(K),
                                   _ll(a.stride(1)), _ll(a.stride(0)), _ll(out.stride(1)), _ll(out.stride(0)), _i(act),
                                   _i(c0), _i(c1), _ptr(wq), _ptr(wk), _i(q_cols[0]), _i(q_cols[1]), _i(k_cols[0]),
                                   _i(k_cols[1]), _ptr(cos_t), _ptr(sin_t), _i(pos_offset), _f(eps), _i(force_bn),
                                   _stream())
    _check(rc, "omgsr_gemm_rope_bf16")
    return out


def posterior_sample(moments, eps, cz, shift, scale, out_dtype=torch.bfloat16):
    """moments: NHWC fp32 [B,h,w,ld>=2cz]; eps: NCHW fp32 [B,cz,h,w] -> z NCHW."""
    _cuda_f32(moments, "moments"); _cuda_f32(eps, "eps")
    B, h, w, _ = moments.shape
    z = torch.empty((B, cz, h, w), device=moments.device, dtype=out_dtype)
    rc = _lib.omgsr_posterior_sample(_ptr(moments), _ll(moments.stride(2)), _ptr(eps.contiguous()), _ptr(z),
                                     _i(int(out_dtype == torch.float32)), _i(B), _i(cz), _i(h), _i(w), _f(shift),
                                     _f(scale), _stream())
    _check(rc, "omgsr_posterior_sample")
    return z


def latent_axpby(x, y, a, b, c=0.0, out=None):
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.bfloat16, torch.float32)
    if y is not None:
        assert y.dtype == x.dtype and y.is_contiguous() and y.numel() == x.numel()
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.omgsr_latent_axpby(_ptr(x), _ptr(y), _ptr(out), _i(int(x.dtype == torch.float32)), _ll(x.numel()),
                                 _f(a), _f(b), _f(c), _stream())
    _check(rc, "omgsr_latent_axpby")
    return out


def x0_unscale(z, eps_hat, s1, s2, scale, shift=0.0, wpq=None, bpq=None):
    """z NCHW (bf16|fp32), eps_hat NHWC fp32 [B,h,w,ld] -> NCHW bf16."""
    _cuda_f32(eps_hat, "eps_hat"); _cuda_f32(wpq, "wpq"); _cuda_f32(bpq, "bpq")
    assert z.is_contiguous()
    B, cz, h, w = z.shape
    out = torch.empty((B, cz, h, w), device=z.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_x0_unscale(_ptr(z), _i(int(z.dtype == torch.float32)), _ptr(eps_hat), _ll(eps_hat.stride(2)),
                               _ptr(out), _ptr(wpq), _ptr(bpq), _i(B), _i(cz), _i(h), _i(w), _f(s1), _f(s2), _f(scale),
                               _f(shift), _stream())
    _check(rc, "omgsr_x0_unscale")
    return out


def pack_latents(z, out=None):
    _cuda_bf16(z, "z")
    B, C, h, w = z.shape
    if out is None:
        out = torch.empty((B, (h // 2) * (w // 2), 4 * C), device=z.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_pack_latents(_ptr(z.contiguous()), _ptr(out), _i(B), _i(C), _i(h), _i(w), _ll(out.stride(1)),
                                 _stream())
    _check(rc, "omgsr_pack_latents")
    return out


def unpack_latents(tokens, C, h, w):
    _cuda_bf16(tokens, "tokens")
    B = tokens.shape[0]
    z = torch.empty((B, C, h, w), device=tokens.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_unpack_latents(_ptr(tokens), _ptr(z), _i(B), _i(C), _i(h), _i(w), _ll(tokens.stride(1)), _stream())
    _check(rc, "omgsr_unpack_latents")
    return z


def tile_blend(tiles, weights, ofs_xy, out, T):
    """tiles: logical [ntiles,B,C,T,T] view (any strides, bf16|fp32); out: fp32 logical [B,C,h,w] view (any strides);
    weights fp32 [T,T]; ofs_xy int32 device [ntiles,2]."""
    _cuda_f32(weights, "weights"); _cuda_f32(out, "out")
    nt, B, C = tiles.shape[:3]
    h, w = out.shape[2], out.shape[3]
    assert ofs_xy.is_cuda and ofs_xy.dtype == torch.int32 and tiles.dtype in (torch.float32, torch.bfloat16)
    ts = (ctypes.c_longlong * 5)(*tiles.stride())
    os_ = (ctypes.c_longlong * 4)(*out.stride())
    rc = _lib.omgsr_tile_blend(_ptr(tiles), _i(int(tiles.dtype == torch.float32)), _ptr(weights), _ptr(ofs_xy),
                               _i(nt), _ptr(out), _i(B), _i(C), _i(h), _i(w), _i(T), ts, os_, _stream())
    _check(rc, "omgsr_tile_blend")
    return out


def wavelet_fix(content, style):
    _cuda_f32(content, "content"); _cuda_f32(style, "style")
    assert content.shape == style.shape and content.is_contiguous() and style.is_contiguous()
    B, C, H, W = content.shape
    out = torch.empty_like(content)
    t0 = torch.empty_like(content)
    t1 = torch.empty_like(content)
    rc = _lib.omgsr_wavelet_fix(_ptr(content), _ptr(style), _ptr(out), _ptr(t0), _ptr(t1), _i(B), _i(C), _i(H), _i(W),
                                _stream())
    _check(rc, "omgsr_wavelet_fix")
    return out


def atrous_blur(x, radius):
    _cuda_f32(x, "x")
    assert x.is_contiguous() and x.dim() == 4
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    rc = _lib.omgsr_atrous_blur(_ptr(x), _ptr(out), _i(B), _i(C), _i(H), _i(W), _i(int(radius)), _stream())
    _check(rc, "omgsr_atrous_blur")
    return out


def adain_fix(content, style):
    _cuda_f32(content, "content"); _cuda_f32(style, "style")
    assert content.shape == style.shape and content.is_contiguous() and style.is_contiguous()
    B, C, H, W = content.shape
    out = torch.empty_like(content)
    scratch = torch.empty((B * C * 4,), device=content.device, dtype=torch.float64)
    rc = _lib.omgsr_adain_fix(_ptr(content), _ptr(style), _ptr(out), _ptr(scratch), _i(B), _i(C), _i(H), _i(W),
                              _stream())
    _check(rc, "omgsr_adain_fix")
    return out


def nchw_to_nhwc(x, out=None):
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.bfloat16, torch.float32)
    B, C, H, W = x.shape
    if out is None:
        out = torch.empty((B, H, W, (C + 7) // 8 * 8), device=x.device, dtype=torch.bfloat16)[..., :C]
    rc = _lib.omgsr_nchw_to_nhwc(_ptr(x), _i(int(x.dtype == torch.float32)), _ptr(out), _i(B), _i(C), _i(H), _i(W),
                                 _ll(out.stride(2)), _stream())
    _check(rc, "omgsr_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, out_dtype=torch.bfloat16, clamp=False):
    _cuda_bf16(x, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, C, H, W), device=x.device, dtype=out_dtype)
    rc = _lib.omgsr_nhwc_to_nchw(_ptr(x), _ll(x.stride(2)), _ptr(out), _i(int(out_dtype == torch.float32)), _i(B),
                                 _i(C), _i(H), _i(W), _i(int(clamp)), _stream())
    _check(rc, "omgsr_nhwc_to_nchw")
    return out


def flash_attn(q, k, v, scale, out=None):
    """out[b,m,h,:] = softmax(scale * q[b,m,h,:] . k[b,:,h,:]) v[b,:,h,:]. q/k/v: [B,T,H,D] bf16 views."""
    _cuda_bf16(q, "q"); _cuda_bf16(k, "k"); _cuda_bf16(v, "v")
    B, M, H, D = q.shape
    N = k.shape[1]
    if out is None:
        out = torch.empty((B, M, H, D), device=q.device, dtype=torch.bfloat16)
    rc = _lib.omgsr_flash_attn_bf16(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _i(B), _i(H), _i(M), _i(N), _i(D),
                                    _ll(q.stride(1)), _ll(q.stride(2)), _ll(q.stride(0)),
                                    _ll(k.stride(1)), _ll(k.stride(2)), _ll(k.stride(0)),
                                    _ll(v.stride(1)), _ll(v.stride(2)), _ll(v.stride(0)),
                                    _ll(out.stride(1)), _ll(out.stride(2)), _ll(out.stride(0)), _f(scale), _stream())
    _check(rc, "omgsr_flash_attn_bf16")
    return out


def attention_materialised(q, k, v, scale, max_score_bytes=2 << 30):
    """softmax(scale q k^T) v through the batched tcgen05 GEMMs + row softmax, for head sizes the fused kernels do not
    serve (the VAE mid block: one head, d = 512). The fp32 scores go through HBM, but only `max_score_bytes` of them at
    a time: query rows are processed in chunks (rows are independent under softmax), so a 65 536-token untiled 2048^2
    image needs 2 GB of scratch instead of 17 GB per image, and the chunk that was just written is re-read out of L2 /
    HBM while still hot.
    (A fused d = 512 kernel is not built, DESIGN.md §8: the 128 x 512 fp32 O tile alone fills TMEM, so O must be split
    over two passes that each recompute S = Q K^T with 32 K-steps: ~6.1 k tensor cycles per 128 x 128 block against 4.1 k
    for the two GEMMs here, whose remaining cost is the score round trip through HBM.)"""
    B, M, H, _ = q.shape
    n = k.shape[1]
    n_pad = (n + 63) // 64 * 64
    rows = max(128, (max_score_bytes // (4 * n_pad * B * H)) // 128 * 128)
    if rows >= M:
        s = bgemm_qk(q, k)
        p = softmax_rows(s, n, scale)
        return bgemm_pv(p, v)
    out = torch.empty((B, M, H, v.shape[3]), device=q.device, dtype=torch.bfloat16)
    for m0 in range(0, M, rows):
        qc = q[:, m0:m0 + rows]
        p = softmax_rows(bgemm_qk(qc, k), n, scale)
        bgemm_pv(p, v, out=out[:, m0:m0 + rows])
    return out


# ------------------------------------------------------------------------------------------------
# driver-side image pre/post-processing (bit-exact with PIL for 8-bit RGB)
# ------------------------------------------------------------------------------------------------
RESIZE_BICUBIC, RESIZE_LANCZOS = 0, 1


def resample_coeffs(in_size, out_size, flt):
    """host-only: Pillow's (bounds [out,2], coeffs [out,ksize]) tables as numpy int32 arrays."""
    import numpy as np
    k = ctypes.c_int(0)
    _check(_lib.omgsr_resample_coeffs(_i(in_size), _i(out_size), _i(flt), None, None, _i(0), ctypes.byref(k)),
           "omgsr_resample_coeffs")
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, k.value), dtype=np.int32)
    _check(_lib.omgsr_resample_coeffs(_i(in_size), _i(out_size), _i(flt), bounds.ctypes.data_as(ctypes.c_void_p),
                                      coeffs.ctypes.data_as(ctypes.c_void_p), _i(k.value), ctypes.byref(k)),
           "omgsr_resample_coeffs")
    return bounds, coeffs


def resize_u8(img, size_hw, flt=RESIZE_BICUBIC):
    """img: uint8 [H, W, 3] CUDA tensor -> uint8 [h, w, 3]; equals PIL.Image.resize((w, h), filter) bit for bit."""
    assert img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.is_contiguous()
    H, W = img.shape[:2]
    h, w = size_hw
    if (h, w) == (H, W):
        return img.clone()  # PIL returns a copy when the size does not change
    nbytes = int(_lib.omgsr_resize_u8_workspace(_i(H), _i(W), _i(h), _i(w), _i(flt)))
    ws = torch.empty(nbytes, device=img.device, dtype=torch.uint8)
    out = torch.empty((h, w, 3), device=img.device, dtype=torch.uint8)
    _check(_lib.omgsr_resize_u8(_ptr(img), _i(H), _i(W), _ptr(out), _i(h), _i(w), _i(flt), _ptr(ws), _ll(nbytes),
                                _stream()), "omgsr_resize_u8")
    return out


def u8_to_input(img, dtype=torch.bfloat16):
    """uint8 [H, W, 3] -> [1, 3, H, W] in [-1, 1]: to_tensor(img).to(dtype) * 2 - 1."""
    assert img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous() and dtype in (torch.bfloat16, torch.float32)
    H, W = img.shape[:2]
    out = torch.empty((1, 3, H, W), device=img.device, dtype=dtype)
    _check(_lib.omgsr_u8_to_input(_ptr(img), _ptr(out), _i(dtype == torch.float32), _i(H), _i(W), _stream()),
           "omgsr_u8_to_input")
    return out


def output_to_u8(x):
    """[3, H, W] (or [1, 3, H, W]) bf16|fp32 model output -> uint8 [H, W, 3] (clip(x*0.5+0.5,0,1)*255, truncated)."""
    if x.dim() == 4:
        assert x.shape[0] == 1
        x = x[0]
    assert x.is_cuda and x.is_contiguous() and x.shape[0] == 3 and x.dtype in (torch.bfloat16, torch.float32)
    H, W = x.shape[1:]
    out = torch.empty((H, W, 3), device=x.device, dtype=torch.uint8)
    _check(_lib.omgsr_output_to_u8(_ptr(x), _i(x.dtype == torch.float32), _ptr(out), _i(H), _i(W), _stream()),
           "omgsr_output_to_u8")
    return out
