"""GPU parity tests of the individual sm_100a kernels against the operator-level oracle (oracle/ops.py),
called through the C ABI (omgsr_b200.capi -> libomgsr_b200.so). Tolerances: bf16 storage with fp32
accumulation, max-abs error <= 1e-2 of the reference magnitude (north_star: latent rel-err <= 2e-2)."""
import math

import pytest
import torch

from tests.util import bf16r, report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from omgsr_b200 import capi as c
    return c


def _g(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 64, 64, 64), (128, 64, 256, 64), (256, 128, 128, 128), (384, 256, 192, 256), (1000, 320, 320, 160),
    (4096, 1280, 1280, 0), (77, 640, 1024, 0), (300, 2560, 320, 0), (128, 72, 64, 0),
])
def test_gemm(capi, M, N, K, bn):
    from oracle import ops
    g = _g(M + N + K)
    a = bf16r(torch.randn(M, K, generator=g))
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    ref = ops.linear(a, w, b)
    out = capi.gemm(a.cuda().bfloat16(), w.cuda().bfloat16(), b.cuda(), force_bn=bn)
    torch.cuda.synchronize()
    ok, _ = report(f"gemm {M}x{N}x{K} bn={bn}", out, ref)
    assert ok


def test_gemm_epilogues(capi):
    from oracle import ops
    g = _g(7)
    M, N, K = 512, 640, 320
    a = bf16r(torch.randn(M, K, generator=g))
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    res = bf16r(torch.randn(M, N, generator=g))
    gate = torch.randn(2, N, generator=g)
    lin = ops.linear(a, w, b)
    ac, wc, bc = a.cuda().bfloat16(), w.cuda().bfloat16(), b.cuda()
    for act, fn in [(capi.ACT_SILU, torch.nn.functional.silu), (capi.ACT_GELU_ERF, torch.nn.functional.gelu),
                    (capi.ACT_GELU_TANH, lambda t: torch.nn.functional.gelu(t, approximate="tanh"))]:
        out = capi.gemm(ac, wc, bc, act=act)
        assert report(f"gemm act={act}", out, fn(lin))[0]
    out = capi.gemm(ac, wc, bc, res=res.cuda().bfloat16(), gate=gate.cuda(), gate_rows=256)
    ref = res + gate.repeat_interleave(256, 0) * lin
    assert report("gemm gate+res", out, ref)[0]
    out = capi.gemm(ac, wc, bc, out_f32=True)
    assert out.dtype == torch.float32
    assert report("gemm f32 out", out, lin, tol=2e-3)[0]
    # strided A (column slice of a wider buffer) and strided output
    wide = torch.zeros(M, K + 64, device="cuda", dtype=torch.bfloat16)
    wide[:, 64:] = ac
    dst = torch.zeros(M, N + 128, device="cuda", dtype=torch.bfloat16)
    capi.gemm(wide[:, 64:], wc, bc, out=dst[:, 128:])
    assert report("gemm strided", dst[:, 128:], lin)[0]
    assert float(dst[:, :128].abs().max()) == 0.0


def test_gemm_geglu(capi):
    from oracle import ops
    from omgsr_b200.pack import pack_geglu_weight
    g = _g(11)
    M, K, F_ = 256, 320, 1280
    a = bf16r(torch.randn(M, K, generator=g))
    w = bf16r(torch.randn(2 * F_, K, generator=g) / math.sqrt(K))
    b = torch.randn(2 * F_, generator=g)
    ref = ops.geglu(a, w, b)
    for bn in (256, 128):
        wi, bi = pack_geglu_weight(w, b, bn)
        out = capi.gemm(a.cuda().bfloat16(), wi.cuda().bfloat16(), bi.cuda(), act=capi.ACT_GEGLU, force_bn=bn)
        assert out.shape == (M, F_)
        assert report(f"geglu bn={bn}", out, ref)[0]


# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad(t,l,b,r), upsample
    (1, 16, 16, 64, 64, 3, 1, (1, 1, 1, 1), False),
    (2, 32, 32, 128, 128, 3, 1, (1, 1, 1, 1), False),
    (1, 128, 128, 128, 256, 3, 1, (1, 1, 1, 1), False),
    (2, 8, 8, 320, 320, 3, 1, (1, 1, 1, 1), False),
    (1, 8, 8, 128, 128, 3, 1, (1, 1, 1, 1), False),
    (3, 20, 24, 64, 128, 3, 1, (1, 1, 1, 1), False),     # ragged tiles
    (1, 150, 150, 64, 64, 3, 1, (1, 1, 1, 1), False),    # vaehook-style odd tile
    (2, 32, 32, 256, 128, 1, 1, (0, 0, 0, 0), False),    # conv_shortcut
    (2, 32, 32, 128, 128, 3, 2, (0, 0, 1, 1), False),    # VAE downsample (asymmetric pad)
    (2, 16, 16, 320, 320, 3, 2, (1, 1, 1, 1), False),    # UNet downsample
    (1, 64, 64, 64, 64, 3, 2, (0, 0, 1, 1), False),
    (2, 16, 16, 128, 128, 3, 1, (1, 1, 1, 1), True),     # nearest-2x + conv, collapsed to 4 x (2x2)
    (1, 8, 8, 320, 320, 3, 1, (1, 1, 1, 1), True),
    (1, 64, 64, 256, 256, 3, 1, (1, 1, 1, 1), True),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,pad,up", CONV_CASES)
def test_conv(capi, B, H, W, Cin, Cout, k, stride, pad, up):
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight, pack_upsample_weight
    g = _g(B * 1000 + H + Cin + Cout + k + stride)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g))
    w = bf16r(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k))
    b = torch.randn(Cout, generator=g)
    ref = ops.conv2d_nhwc(x, w, b, stride=stride, pad=pad, upsample2x=up)
    pw = pack_upsample_weight(w) if up else pack_conv_weight(w)
    out = capi.conv2d(x.cuda().bfloat16(), pw.cuda(), b.cuda(), ksize=k, stride=stride, pad=pad, upsample2x=up)
    torch.cuda.synchronize()
    tol = 2e-2 if up else 1e-2   # collapsed taps are re-rounded to bf16 once
    assert report(f"conv {B}x{H}x{W} {Cin}->{Cout} k{k} s{stride} up={up}", out, ref, tol)[0]


HALO_CASES = [
    # B, H, W, Cin, Cout, upsample   (W >= 96 and Cout % 128 == 0 -> input-halo kernel)
    (1, 8, 128, 64, 128, False),
    (2, 40, 256, 128, 128, False),
    (1, 33, 200, 128, 256, False),      # ragged in x (masked tail) and odd row count
    (1, 64, 128, 256, 512, False),      # two n-tiles of 256
    (1, 130, 128, 64, 128, False),
    (1, 32, 128, 128, 128, True),       # nearest-2x + conv: four 2x2 parity convs through the halo kernel
    (2, 64, 96, 64, 256, True),
    (2, 40, 256, 128, 3, False),        # thin Cout (decoder conv_out) on one masked 128-wide tile
    (1, 6, 128, 64, 128, False),        # three halo tiles: the CTA pair's odd tail (one dummy tile)
    (3, 10, 300, 192, 128, False),      # 45 tiles, ragged x, three K chunks
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,up", HALO_CASES)
@pytest.mark.parametrize("mode", ["pair", "single"])
def test_conv_halo(capi, B, H, W, Cin, Cout, up, mode):
    from oracle import ops
    capi.lib().omgsr_set_gemm_2cta(1 if mode == "pair" else 0)
    from omgsr_b200.pack import pack_conv_weight, pack_upsample_weight
    g = _g(B * 7 + H + W + Cin + Cout)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g))
    w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    b = torch.randn(Cout, generator=g)
    ref = ops.conv2d_nhwc(x, w, b, upsample2x=up)
    rc = None
    if Cout % 8 == 0:
        res = bf16r(torch.randn(B, H * (2 if up else 1), W * (2 if up else 1), Cout, generator=g))
        ref = ref + res
        rc = res.cuda().bfloat16()
    pw = pack_upsample_weight(w) if up else pack_conv_weight(w)
    xc, wc, bc = x.cuda().bfloat16(), pw.cuda(), b.cuda()
    out = capi.conv2d(xc, wc, bc, res=rc, upsample2x=up, cout=Cout)
    torch.cuda.synchronize()
    ok, _ = report(f"conv-halo[{mode}] {B}x{H}x{W} {Cin}->{Cout} up={up}", out, ref, 2e-2 if up else 1e-2)
    gen = capi.conv2d(xc, wc, bc, res=rc, upsample2x=up, flags=capi.CONV_NO_HALO, cout=Cout)
    capi.lib().omgsr_set_gemm_2cta(1)
    assert report("generic kernel, same problem", gen, ref, 2e-2 if up else 1e-2)[0]
    assert ok


@pytest.mark.parametrize("B,H,W,Cin,Cout,up,stride", [(2, 32, 32, 128, 128, False, 1), (1, 16, 128, 64, 128, False, 1),
                                                      (2, 24, 40, 256, 256, False, 1), (1, 16, 16, 128, 512, False, 1),
                                                      (2, 16, 16, 128, 128, True, 1), (2, 32, 32, 128, 128, False, 2)])
def test_conv_fused_groupnorm_stats(capi, B, H, W, Cin, Cout, up, stride):
    """the conv epilogue also produces (sum, sumsq) per (image, group) of its stored output."""
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight, pack_upsample_weight
    g = _g(B + H + W + Cin + Cout + 3)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g))
    w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    b = torch.randn(Cout, generator=g)
    pad = (0, 0, 1, 1) if stride == 2 else (1, 1, 1, 1)
    OH, OW = (2 * H, 2 * W) if up else (H // stride, W // stride)
    res = bf16r(torch.randn(B, OH, OW, Cout, generator=g))
    pw = pack_upsample_weight(w) if up else pack_conv_weight(w)
    st = torch.full((B, 32, 2), 123.0, device="cuda")
    out = capi.conv2d(x.cuda().bfloat16(), pw.cuda(), b.cuda(), res=res.cuda().bfloat16(), stride=stride, pad=pad,
                      upsample2x=up, gn_stats=st)
    ref = ops.conv2d_nhwc(x, w, b, stride=stride, pad=pad, upsample2x=up) + res
    assert report("conv (with fused stats)", out, ref, 2e-2 if up else 1e-2)[0]
    o = out.float().cpu().reshape(B, OH * OW, 32, Cout // 32)
    sref = torch.stack([o.sum(dim=(1, 3)), o.pow(2).sum(dim=(1, 3))], dim=-1)
    assert report(f"fused gn stats C={Cout}", st, sref, 1e-3)[0]
    # and the GroupNorm built on them equals the two-pass one
    gam, bet = torch.randn(Cout, generator=g).cuda(), torch.randn(Cout, generator=g).cuda()
    a = capi.groupnorm_apply(out, st, gam, bet, 32, 1e-6, True)
    bb = capi.groupnorm(out, gam, bet, 32, 1e-6, True)
    assert report("gn from fused stats", a, bb, 5e-3)[0]


def test_conv_residual_silu_concat(capi):
    """residual epilogue + reading a channel slice of a concat buffer + writing into a slice."""
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight
    g = _g(21)
    B, H, W, C1, C2, Cout = 2, 16, 16, 128, 192, 128
    x = bf16r(torch.randn(B, H, W, C1 + C2, generator=g))
    w = bf16r(torch.randn(Cout, C1 + C2, 3, 3, generator=g) / math.sqrt(9 * (C1 + C2)))
    b = torch.randn(Cout, generator=g)
    res = bf16r(torch.randn(B, H, W, Cout, generator=g))
    ref = ops.conv2d_nhwc(x, w, b) + res
    dst = torch.zeros(B, H, W, Cout + 64, device="cuda", dtype=torch.bfloat16)
    capi.conv2d(x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda(), res=res.cuda().bfloat16(),
                out=dst[..., 64:])
    assert report("conv res+slice", dst[..., 64:], ref)[0]
    assert float(dst[..., :64].abs().max()) == 0.0
    out = capi.conv2d(x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda(), act=capi.ACT_SILU)
    assert report("conv silu", out, torch.nn.functional.silu(ops.conv2d_nhwc(x, w, b)))[0]
    # thin Cout through the tensor-core path with fp32 output (encoder conv_out / unet conv_out)
    w8 = bf16r(torch.randn(8, C1 + C2, 3, 3, generator=g) / math.sqrt(9 * (C1 + C2)))
    b8 = torch.randn(8, generator=g)
    out = capi.conv2d(x.cuda().bfloat16(), pack_conv_weight(w8).cuda(), b8.cuda(), out_f32=True)
    assert report("conv Cout=8 f32", out, ops.conv2d_nhwc(x, w8, b8), 2e-3)[0]


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,Hh,M,N,D", [(1, 1, 128, 128, 64), (2, 5, 256, 256, 64), (2, 5, 1024, 77, 64),
                                         (1, 2, 200, 300, 64), (1, 1, 256, 256, 512), (1, 3, 256, 320, 128),
                                         # rows of >= 1024 scores: register-resident softmax (2 / 4 / 8 float4 per thread)
                                         (1, 1, 256, 1024, 512), (1, 1, 128, 1990, 64), (1, 1, 128, 4096, 512),
                                         (1, 1, 128, 5000, 64),
                                         # 16 / 32 float4 per thread (1024^2 image: 16384 tokens in the VAE mid block)
                                         (1, 1, 128, 16384, 64), (1, 1, 128, 20000, 64)])
def test_attention_materialised(capi, B, Hh, M, N, D):
    from oracle import ops
    g = _g(B + Hh + M + N + D)
    q = bf16r(torch.randn(B, M, Hh, D, generator=g))
    k = bf16r(torch.randn(B, N, Hh, D, generator=g))
    v = bf16r(torch.randn(B, N, Hh, D, generator=g))
    scale = 1.0 / math.sqrt(D)
    s = capi.bgemm_qk(q.cuda().bfloat16(), k.cuda().bfloat16())
    sref = torch.einsum("bmhd,bnhd->bhmn", q, k)
    assert report(f"bgemm qk {B},{Hh},{M},{N},{D}", s[..., :N], sref, 2e-3)[0]
    p = capi.softmax_rows(s, N, scale)
    pref = (sref * scale).softmax(-1)
    assert report("softmax", p[..., :N], pref)[0]
    assert float(p[..., N:].abs().max()) == 0.0 if p.shape[-1] > N else True
    o = capi.bgemm_pv(p, v.cuda().bfloat16())
    assert report(f"bgemm pv (MN-major V) {B},{Hh},{M},{N},{D}", o, ops.attention(q, k, v, scale), 2e-2)[0]


def test_gemm_rope_epilogue(capi):
    """fused q|mlp|k|v projection with RMSNorm + RoPE of the q / k heads in the GEMM epilogue (FLUX single-block layout)
    against fp32 torch, and against the unfused pair (GEMM, then the in-place rmsnorm_rope kernel)."""
    import torch.nn.functional as F
    g = _g(777)
    B, M, K, Hh = 2, 300, 256, 2
    d, dm = Hh * 128, 384
    N = 3 * d + dm                                   # [q | mlp | k | v]
    a = bf16r(torch.randn(B, M, K, generator=g))
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g) * 0.1
    wq, wk = 1 + 0.1 * torch.randn(128, generator=g), 1 + 0.1 * torch.randn(128, generator=g)
    pos0 = 5
    ang = torch.rand(pos0 + M, 64, generator=g) * 6.28
    cos_t = ang.cos().repeat_interleave(2, dim=1).contiguous()
    sin_t = ang.sin().repeat_interleave(2, dim=1).contiguous()
    y = F.linear(a, w, bias)                          # fp32 reference

    def head_ref(x, wn):                              # x [B, M, heads, 128]
        xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * wn
        x2 = xn.reshape(*xn.shape[:-1], 64, 2)
        rot = torch.stack([-x2[..., 1], x2[..., 0]], -1).flatten(-2)
        c, s_ = cos_t[pos0:pos0 + M][None, :, None], sin_t[pos0:pos0 + M][None, :, None]
        return xn * c + rot * s_

    ref = y.clone()
    ref[..., :d] = head_ref(y[..., :d].reshape(B, M, Hh, 128), wq).reshape(B, M, d)
    ref[..., d + dm:2 * d + dm] = head_ref(y[..., d + dm:2 * d + dm].reshape(B, M, Hh, 128), wk).reshape(B, M, d)
    ref[..., d:d + dm] = F.gelu(y[..., d:d + dm], approximate="tanh")
    ac, wc, bc = a.cuda().bfloat16(), w.cuda().bfloat16(), bias.cuda()
    args = (wq.cuda(), wk.cuda(), (0, d), (d + dm, 2 * d + dm), cos_t.cuda(), sin_t.cuda())
    got = capi.gemm_rope(ac, wc, bc, *args, pos_offset=pos0, act=capi.ACT_GELU_TANH, act_cols=(d, d + dm))
    assert report("gemm + rmsnorm + rope epilogue", got, ref, 1e-2)[0]
    # the unfused pair of round 1
    un = capi.gemm_batched(ac, wc, bc, act=capi.ACT_GELU_TANH, act_cols=(d, d + dm))
    capi.rmsnorm_rope_(un[..., :d], wq.cuda().reshape(1, 128), cos_t.cuda(), sin_t.cuda(), Hh, Hh, pos_offset=pos0)
    capi.rmsnorm_rope_(un[..., d + dm:2 * d + dm], wk.cuda().reshape(1, 128), cos_t.cuda(), sin_t.cuda(), Hh, Hh, pos_offset=pos0)
    assert report("fused vs unfused", got, un.float(), 1.5e-2)[0]
    # into a strided destination (rows of a wider joint buffer), BN = 128 path (N = 128 * 11 is not a multiple of 256)
    wide = torch.zeros(B, M + 7, N + 64, device="cuda", dtype=torch.bfloat16)
    capi.gemm_rope(ac, wc, bc, *args, pos_offset=pos0, act=capi.ACT_GELU_TANH, act_cols=(d, d + dm), out=wide[:, 7:, :N])
    assert torch.equal(wide[:, 7:, :N], got) and float(wide[:, :7].abs().max()) == 0.0


def test_attention_materialised_query_chunks(capi):
    """the d = 512 path bounds its fp32 score scratch by chunking over query rows: same bits as the one-shot form."""
    from oracle import ops
    g = _g(4242)
    B, Hh, M, N, D = 2, 1, 1000, 1990, 512
    q = bf16r(torch.randn(B, M, Hh, D, generator=g))
    k = bf16r(torch.randn(B, N, Hh, D, generator=g))
    v = bf16r(torch.randn(B, N, Hh, D, generator=g))
    qc, kc, vc = q.cuda().bfloat16(), k.cuda().bfloat16(), v.cuda().bfloat16()
    scale = 1.0 / math.sqrt(D)
    whole = capi.attention_materialised(qc, kc, vc, scale)
    chunked = capi.attention_materialised(qc, kc, vc, scale, max_score_bytes=4 * 2048 * B * 256)  # 256-row chunks
    assert torch.equal(whole, chunked)
    assert report("materialised attention d=512", whole, ops.attention(q, k, v, scale), 2e-2)[0]


@pytest.mark.parametrize("B,Hh,M,N,D", [(1, 1, 128, 128, 64), (2, 5, 256, 384, 64), (2, 5, 1024, 77, 64),
                                         (1, 2, 200, 300, 64), (1, 10, 64, 64, 64), (1, 2, 4096, 4096, 64),
                                         (1, 1, 128, 64, 128), (1, 3, 256, 320, 128), (1, 2, 4608, 4608, 128)])
def test_flash_attention(capi, B, Hh, M, N, D):
    from oracle import ops
    g = _g(B + Hh + M + N + D + 1)
    q = bf16r(torch.randn(B, M, Hh, D, generator=g))
    k = bf16r(torch.randn(B, N, Hh, D, generator=g))
    v = bf16r(torch.randn(B, N, Hh, D, generator=g))
    scale = 1.0 / math.sqrt(D)
    o = capi.flash_attn(q.cuda().bfloat16(), k.cuda().bfloat16(), v.cuda().bfloat16(), scale)
    torch.cuda.synchronize()
    assert report(f"flash {B},{Hh},{M},{N},{D}", o, ops.attention(q, k, v, scale), 2e-2)[0]


def test_flash_attention_variants_agree(capi):
    """the single-tile kernel (forced) and the ping-pong kernel give the same result on a long sequence."""
    from oracle import ops
    g = _g(123)
    B, Hh, M, N, D = 1, 3, 1000, 1111, 128
    q = bf16r(torch.randn(B, M, Hh, D, generator=g) * 2)
    k = bf16r(torch.randn(B, N, Hh, D, generator=g) * 2)
    v = bf16r(torch.randn(B, N, Hh, D, generator=g))
    ref = ops.attention(q, k, v, 0.25)   # large logits: exercises the lazy O rescale of the ping-pong kernel
    qc, kc, vc = q.cuda().bfloat16(), k.cuda().bfloat16(), v.cuda().bfloat16()
    o2 = capi.flash_attn(qc, kc, vc, 0.25)
    capi.lib().omgsr_flash_attn_set_variant(1)
    try:
        o1 = capi.flash_attn(qc, kc, vc, 0.25)
    finally:
        capi.lib().omgsr_flash_attn_set_variant(0)
    assert report("flash v1 (single tile)", o1, ref, 2e-2)[0]
    assert report("flash auto (d=128: v3)", o2, ref, 2e-2)[0]
    for variant, name in ((2, "v2 (64-key ping-pong)"), (3, "v3 (128 keys)"), (4, "v4 (128 keys, transposed O)"),
                          (5, "v5 (128 keys, S released from registers)")):
        capi.lib().omgsr_flash_attn_set_variant(variant)
        try:
            ov = capi.flash_attn(qc, kc, vc, 0.25)
        finally:
            capi.lib().omgsr_flash_attn_set_variant(0)
        assert report("flash " + name, ov, ref, 2e-2)[0]
    # growing maxima along the key axis force the lazy O rescale (v4: per-group column rescale of O^T) in every block
    ramp = torch.linspace(0.0, 6.0, N)[None, :, None, None]
    k2 = bf16r(k * 0.2 + ramp * torch.sign(q.mean(dim=1, keepdim=True)))
    ref2 = ops.attention(q, k2, v, 0.25)
    for variant in (0, 4, 5):
        capi.lib().omgsr_flash_attn_set_variant(variant)
        try:
            o4 = capi.flash_attn(qc, k2.cuda().bfloat16(), vc, 0.25)
        finally:
            capi.lib().omgsr_flash_attn_set_variant(0)
        assert report(f"flash variant {variant} with rescales", o4, ref2, 2e-2)[0]


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("variant", [0, 4, 5])
@pytest.mark.parametrize("shape", ["ramp", "spike"])
def test_flash_attention_reference_rebase(capi, D, variant, shape):
    """One-pass softmax of v3 / v4: the row's reference exponent moves when a block's row sum exceeds 2^10.
    'ramp': scores grow along the key axis, so the reference moves in most blocks; 'spike': a few keys in a late
    block score up to ~40 (log2 units) above everything before them (p ~ 2^40 for one block, rescaled away in the next).
    Envelope of the one-pass form (stated in attn_tc.cu): within ONE 128-key block a row's scaled scores may exceed the
    running reference by < 127 log2 units (88 nats); the 64-key kernel (variant 2) has no such limit."""
    from oracle import ops
    g = _g(7 + D)
    B, M, N, Hh = 1, 384, 1024, 2
    q = bf16r(torch.randn(B, M, Hh, D, generator=g))
    k = bf16r(torch.randn(B, N, Hh, D, generator=g) * 0.3)
    v = bf16r(torch.randn(B, N, Hh, D, generator=g))
    qdir = q.mean(dim=1, keepdim=True)
    qdir = qdir / qdir.norm(dim=-1, keepdim=True)
    if shape == "ramp":
        k = bf16r(k + torch.linspace(0.0, 8.0, N)[None, :, None, None] * torch.sign(qdir))
    else:
        k[:, 700:704] = bf16r(k[:, 700:704] + 8.0 * torch.sign(qdir))   # |q . sign| <~ 3.5 sqrt(D): jumps of <~ 40 (log2 units)
    scale = D ** -0.5
    ref = ops.attention(q, k, v, scale)
    capi.lib().omgsr_flash_attn_set_variant(variant)
    try:
        o = capi.flash_attn(q.cuda().bfloat16(), k.cuda().bfloat16(), v.cuda().bfloat16(), scale)
    finally:
        capi.lib().omgsr_flash_attn_set_variant(0)
    assert torch.isfinite(o.float()).all()
    assert report(f"flash d={D} variant={variant} {shape}", o, ref, 2e-2)[0]


def test_flash_attention_fused_qkv_layout(capi):
    """q/k/v as column slices of one fused [B, T, 3*H*D] projection output (strided views)."""
    from oracle import ops
    g = _g(99)
    B, T, Hh, D = 2, 256, 5, 64
    qkv = bf16r(torch.randn(B, T, 3 * Hh * D, generator=g))
    q, k, v = [qkv[..., i * Hh * D:(i + 1) * Hh * D].reshape(B, T, Hh, D) for i in range(3)]
    qc = qkv.cuda().bfloat16()
    qv, kv, vv = [qc[..., i * Hh * D:(i + 1) * Hh * D].unflatten(-1, (Hh, D)) for i in range(3)]
    o = capi.flash_attn(qv, kv, vv, 0.125)
    assert report("flash fused-qkv", o, ops.attention(q, k, v, 0.125), 2e-2)[0]


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,HW,C,eps,silu", [(2, 64 * 64, 128, 1e-6, True), (1, 32 * 32, 512, 1e-6, False),
                                              (3, 16 * 16, 320, 1e-5, True), (2, 8 * 8, 2560, 1e-5, True),
                                              (1, 150 * 150, 256, 1e-6, True), (2, 64, 960, 1e-5, True)])
def test_groupnorm(capi, B, HW, C, eps, silu):
    from oracle import ops
    g = _g(B + HW + C)
    x = bf16r(torch.randn(B, HW, C, generator=g) * 2 + 0.5)
    gamma = torch.randn(C, generator=g)
    beta = torch.randn(C, generator=g)
    ref = ops.groupnorm_nhwc(x, gamma, beta, 32, eps, silu)
    out = capi.groupnorm(x.cuda().bfloat16(), gamma.cuda(), beta.cuda(), 32, eps, silu)
    assert report(f"groupnorm {B},{HW},{C}", out, ref)[0]


@pytest.mark.parametrize("rows,C", [(512, 320), (100, 640), (64, 1280), (4608, 3072)])
def test_layernorm(capi, rows, C):
    from oracle import ops
    g = _g(rows + C)
    x = bf16r(torch.randn(rows, C, generator=g) * 3 + 1)
    gamma = torch.randn(C, generator=g)
    beta = torch.randn(C, generator=g)
    out = capi.layernorm(x.cuda().bfloat16(), gamma.cuda(), beta.cuda(), 1e-5)
    assert report(f"layernorm {rows}x{C}", out, ops.layernorm(x, gamma, beta, 1e-5))[0]
    # AdaLN modulation: per-image (1+scale, shift), no learned affine
    nb = 2
    sc = torch.randn(nb, C, generator=g)
    sh = torch.randn(nb, C, generator=g)
    rpb = rows // nb
    ref = ops.layernorm(x, None, None, 1e-6).reshape(nb, rpb, C) * (1 + sc[:, None]) + sh[:, None]
    out = capi.layernorm(x.cuda().bfloat16(), (1 + sc).cuda().contiguous(), sh.cuda(), 1e-6, rows_per_batch=rpb,
                         gb_stride=C)
    assert report("layernorm modulate", out, ref.reshape(rows, C))[0]
    # batch-strided input view (image tokens inside a joint [text; image] sequence)
    joint = torch.zeros(nb, rpb + 16, C, device="cuda", dtype=torch.bfloat16)
    joint[:, 16:] = x.cuda().bfloat16().view(nb, rpb, C)
    out = capi.layernorm(joint[:, 16:], (1 + sc).cuda().contiguous(), sh.cuda(), 1e-6, gb_stride=C)
    assert report("layernorm strided view", out.reshape(rows, C), ref.reshape(rows, C))[0]


def test_thin_convs_and_layout(capi):
    from oracle import ops
    from omgsr_b200.pack import pack_conv_in_weight, pack_conv_weight, pad8
    g = _g(5)
    for (B, Cin, H, W, Cout, k) in [(2, 3, 40, 48, 128, 3), (1, 4, 16, 16, 320, 3), (2, 16, 8, 8, 512, 3),
                                     (1, 4, 16, 16, 8, 1), (2, 4, 12, 20, 4, 1)]:
        x = bf16r(torch.randn(B, Cin, H, W, generator=g))
        w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
        b = torch.randn(Cout, generator=g)
        ref = torch.nn.functional.conv2d(x, w, b, padding=k // 2).permute(0, 2, 3, 1)
        for dt in (torch.bfloat16, torch.float32):
            out = capi.conv_in_nchw(x.cuda().to(dt), pack_conv_in_weight(w).cuda(), pad8(b).cuda(), Cout, k)
            assert report(f"conv_in {Cin}->{Cout} {dt}", out, ref)[0]
    from omgsr_b200.pack import pack_conv_in_gemm_weight
    for (B, Cin, H, W, Cout) in [(2, 3, 40, 48, 128), (1, 4, 16, 16, 320), (1, 16, 12, 20, 512)]:
        x = bf16r(torch.randn(B, Cin, H, W, generator=g))
        w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
        b = torch.randn(Cout, generator=g)
        ref = torch.nn.functional.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)
        out = capi.conv_in_gemm(x.cuda().bfloat16(), pack_conv_in_gemm_weight(w).cuda(), b.cuda())
        assert report(f"conv_in im2col+gemm {Cin}->{Cout}", out, ref)[0]
    for (B, H, W, Cin, Cout) in [(2, 24, 40, 128, 3), (1, 16, 16, 320, 4)]:
        x = bf16r(torch.randn(B, H, W, Cin, generator=g))
        w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
        b = torch.randn(Cout, generator=g)
        ref = ops.conv2d_nhwc(x, w, b).permute(0, 3, 1, 2)
        out = capi.conv_out_nchw(x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda(), Cout, clamp=False,
                                 out_dtype=torch.float32)
        assert report(f"conv_out {Cin}->{Cout}", out, ref, 2e-3)[0]
        out = capi.conv_out_nchw(x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda(), Cout, clamp=True)
        assert report("conv_out clamp", out, ref.clamp(-1, 1))[0]
    x = bf16r(torch.randn(2, 20, 9, 11, generator=g))
    nhwc = capi.nchw_to_nhwc(x.cuda().bfloat16())
    assert report("nchw->nhwc", nhwc, x.permute(0, 2, 3, 1), 1e-6)[0]
    back = capi.nhwc_to_nchw(nhwc, torch.float32)
    assert report("nhwc->nchw", back, x, 1e-6)[0]


def test_latent_steps(capi):
    g = _g(31)
    B, Cz, h, w = 2, 4, 16, 24
    mom = torch.randn(B, h, w, 8, generator=g)
    eps = torch.randn(B, Cz, h, w, generator=g)
    mean = mom[..., :Cz].permute(0, 3, 1, 2)
    logvar = mom[..., Cz:].permute(0, 3, 1, 2).clamp(-30, 20)
    ref = ((mean + torch.exp(0.5 * logvar) * eps) - 0.1159) * 0.3611
    z = capi.posterior_sample(mom.cuda(), eps.cuda(), Cz, 0.1159, 0.3611, torch.float32)
    assert report("posterior", z, ref, 2e-4)[0]   # fp32 with the fast exp intrinsic
    x = torch.randn(1000, generator=g)
    y = torch.randn(1000, generator=g)
    out = capi.latent_axpby(x.cuda(), y.cuda(), 0.495, 0.505, 0.0)   # z_t = (1-sigma) z + sigma eps
    assert report("axpby f32", out, 0.495 * x + 0.505 * y, 1e-6)[0]
    out = capi.latent_axpby(x.cuda().bfloat16(), y.cuda().bfloat16(), 1.0, -0.505112)
    assert report("axpby bf16", out, bf16r(x) - 0.505112 * bf16r(y))[0]
    # x0 prediction + post_quant_conv
    zz = torch.randn(B, Cz, h, w, generator=g)
    eh = torch.randn(B, h, w, 8, generator=g)
    Wpq = torch.randn(Cz, Cz, generator=g)
    bpq = torch.randn(Cz, generator=g)
    a = 0.6357423
    z0 = (zz - math.sqrt(1 - a) * eh[..., :Cz].permute(0, 3, 1, 2)) / math.sqrt(a) / 0.18215
    ref = torch.einsum("oc,bchw->bohw", Wpq, z0) + bpq[None, :, None, None]
    out = capi.x0_unscale(zz.cuda(), eh.cuda(), math.sqrt(1 - a), math.sqrt(a), 0.18215, 0.0, Wpq.cuda(), bpq.cuda())
    assert report("x0+postquant", out, ref)[0]
    # pack / unpack (infer/omgsr_f_infer_model.py:21-41)
    lat = bf16r(torch.randn(2, 16, 8, 12, generator=g))
    ref = lat.view(2, 16, 4, 2, 6, 2).permute(0, 2, 4, 1, 3, 5).reshape(2, 24, 64)
    tok = capi.pack_latents(lat.cuda().bfloat16())
    assert report("pack", tok, ref, 1e-6)[0]
    assert report("unpack", capi.unpack_latents(tok, 16, 8, 12), lat, 1e-6)[0]


def test_rmsnorm_rope(capi):
    g = _g(41)
    rows, nb, Hh, off = 96, 2, 3, 16
    # fused [q | k | v] rows: q and k normalised with their own weights in one launch, v untouched
    x = bf16r(torch.randn(nb, rows, 3 * Hh * 128, generator=g))
    wt = torch.rand(2, 128, generator=g) + 0.5
    ang = torch.randn(off + rows, 64, generator=g)
    cos = torch.cos(ang).repeat_interleave(2, -1)
    sin = torch.sin(ang).repeat_interleave(2, -1)
    qk = x[..., :2 * Hh * 128].reshape(nb, rows, 2, Hh, 128)
    n = qk * torch.rsqrt(qk.pow(2).mean(-1, keepdim=True) + 1e-6) * wt[None, None, :, None, :]
    n2 = n.reshape(nb, rows, 2, Hh, 64, 2)
    rot = torch.stack([-n2[..., 1], n2[..., 0]], -1).reshape(nb, rows, 2, Hh, 128)
    c, s_ = cos[off:][None, :, None, None], sin[off:][None, :, None, None]
    ref = x.clone()
    ref[..., :2 * Hh * 128] = (n * c + rot * s_).reshape(nb, rows, 2 * Hh * 128)
    xc = x.cuda().bfloat16().contiguous()
    capi.rmsnorm_rope_(xc, wt.cuda(), cos.cuda().contiguous(), sin.cuda().contiguous(), 2 * Hh, Hh, pos_offset=off)
    assert report("rmsnorm+rope", xc, ref)[0]


def test_gemm_batched(capi):
    from oracle import ops
    g = _g(43)
    B, S, T0, M, K, N = 2, 384, 128, 256, 128, 320
    joint = bf16r(torch.randn(B, S, K, generator=g))          # [text(128) ; image(256)] rows
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    gate = torch.randn(B, 3 * N, generator=g)
    res = bf16r(torch.randn(B, M, N, generator=g))
    a = joint[:, T0:T0 + M]
    lin = ops.linear(a, w, b)
    ref = res + gate[:, N:2 * N, None].transpose(1, 2) * lin
    jc = joint.cuda().bfloat16()
    out = capi.gemm_batched(jc[:, T0:T0 + M], w.cuda().bfloat16(), b.cuda(), res=res.cuda().bfloat16(),
                            gate=gate.cuda()[:, N:2 * N])
    assert report("gemm_batched gate+res", out, ref)[0]
    # activation on a column range only, written into a slice of a wider joint buffer
    dst = torch.zeros(B, S, N + 64, device="cuda", dtype=torch.bfloat16)
    capi.gemm_batched(jc[:, T0:T0 + M], w.cuda().bfloat16(), b.cuda(), act=capi.ACT_GELU_TANH, act_cols=(64, 192),
                      out=dst[:, T0:T0 + M, 64:])
    ref2 = lin.clone()
    ref2[..., 64:192] = torch.nn.functional.gelu(lin[..., 64:192], approximate="tanh")
    assert report("gemm_batched act range", dst[:, T0:T0 + M, 64:], ref2)[0]
    assert float(dst[:, :T0].abs().max()) == 0.0 and float(dst[..., :64].abs().max()) == 0.0


def test_tile_blend(capi):
    g = _g(51)
    B, C, h, w, T = 1, 4, 24, 24, 16
    offs = [(0, 0), (8, 0), (0, 8), (8, 8)]
    tiles = torch.randn(len(offs), B, C, T, T, generator=g)
    wts = torch.rand(T, T, generator=g) + 0.1
    num = torch.zeros(B, C, h, w)
    den = torch.zeros(B, C, h, w)
    for t, (ox, oy) in enumerate(offs):
        num[:, :, oy:oy + T, ox:ox + T] += tiles[t] * wts
        den[:, :, oy:oy + T, ox:ox + T] += wts
    out = torch.empty(B, C, h, w, device="cuda")
    capi.tile_blend(tiles.cuda(), wts.cuda(), torch.tensor(offs, dtype=torch.int32).cuda(), out, T)
    assert report("tile blend", out, num / den, 1e-5)[0]
    # NHWC-strided tiles and output (the layout the UNet path produces)
    tn = tiles.permute(0, 1, 3, 4, 2).contiguous().cuda()           # [nt,B,T,T,C]
    on = torch.empty(B, h, w, 8, device="cuda")
    capi.tile_blend(tn.permute(0, 1, 4, 2, 3), wts.cuda(), torch.tensor(offs, dtype=torch.int32).cuda(),
                    on[..., :C].permute(0, 3, 1, 2), T)
    assert report("tile blend nhwc", on[..., :C].permute(0, 3, 1, 2), num / den, 1e-5)[0]


def test_colorfix(capi):
    g = _g(61)
    c = torch.rand(2, 3, 70, 90, generator=g)
    s = torch.rand(2, 3, 70, 90, generator=g)
    # oracle restatement lives in oracle/colorfix.py (pinned to the reference's own functions on CPU)
    from oracle import colorfix
    assert report("wavelet", capi.wavelet_fix(c.cuda(), s.cuda()), colorfix.wavelet_reconstruction(c, s), 1e-5)[0]
    assert report("adain", capi.adain_fix(c.cuda(), s.cuda()), colorfix.adaptive_instance_normalization(c, s), 1e-4)[0]


# ------------------------------------------------------------------------------------------------
# CTA-pair (cta_group::2) kernel
# ------------------------------------------------------------------------------------------------
@pytest.fixture(params=["pair", "single"])
def two_cta(capi, request):
    """the CTA-pair kernel is the default; 'single' forces the one-CTA kernel so both stay covered."""
    capi.lib().omgsr_set_gemm_2cta(1 if request.param == "pair" else 0)
    yield request.param
    capi.lib().omgsr_set_gemm_2cta(1)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 256, 256), (384, 512, 192), (4096, 1280, 1280), (1000, 256, 320),
                                   (700, 320, 320), (515, 640, 128), (300, 128, 256),
                                   # K not a multiple of the 64-deep k-block: TMA zero-fills both operands
                                   (300, 128, 32), (515, 256, 160), (256, 320, 40)])
def test_gemm_two_cta(capi, two_cta, M, N, K):
    from oracle import ops
    g = _g(M + N + K + 2)
    a = bf16r(torch.randn(M, K, generator=g))
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    res = bf16r(torch.randn(M, N, generator=g))
    out = capi.gemm(a.cuda().bfloat16(), w.cuda().bfloat16(), b.cuda(), res=res.cuda().bfloat16(), act=capi.ACT_SILU)
    torch.cuda.synchronize()
    assert report(f"gemm 2cta {M}x{N}x{K}", out, torch.nn.functional.silu(ops.linear(a, w, b)) + res)[0]


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,up", [(2, 32, 32, 256, 256, 1, False), (1, 24, 40, 128, 512, 1, False),
                                                      (3, 16, 16, 128, 256, 2, False), (1, 16, 16, 256, 256, 1, True),
                                                      (1, 8, 8, 320, 1280, 1, False), (2, 16, 16, 320, 320, 1, False),
                                                      (1, 32, 32, 256, 128, 1, False)])
def test_conv_two_cta(capi, two_cta, B, H, W, Cin, Cout, stride, up):
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight, pack_upsample_weight
    g = _g(B + H + W + Cin + Cout + 5)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g))
    w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    b = torch.randn(Cout, generator=g)
    pad = (0, 0, 1, 1) if stride == 2 else (1, 1, 1, 1)
    ref = ops.conv2d_nhwc(x, w, b, stride=stride, pad=pad, upsample2x=up)
    pw = pack_upsample_weight(w) if up else pack_conv_weight(w)
    st = torch.zeros(B, 32, 2, device="cuda") if (Cout // 32 in (4, 8, 16) and ref.shape[1] * ref.shape[2] >= 128) else None
    out = capi.conv2d(x.cuda().bfloat16(), pw.cuda(), b.cuda(), stride=stride, pad=pad, upsample2x=up, gn_stats=st)
    torch.cuda.synchronize()
    assert report(f"conv 2cta {B}x{H}x{W} {Cin}->{Cout} s{stride} up={up}", out, ref, 2e-2 if up else 1e-2)[0]
    if st is not None:
        o = out.float().cpu().reshape(B, -1, 32, Cout // 32)
        sref = torch.stack([o.sum(dim=(1, 3)), o.pow(2).sum(dim=(1, 3))], dim=-1)
        assert report("2cta fused gn stats", st, sref, 1e-3)[0]


def test_geglu_two_cta(capi, two_cta):
    from oracle import ops
    from omgsr_b200.pack import pack_geglu_weight
    g = _g(77)
    M, K, F_ = 512, 320, 1280
    a = bf16r(torch.randn(M, K, generator=g))
    w = bf16r(torch.randn(2 * F_, K, generator=g) / math.sqrt(K))
    b = torch.randn(2 * F_, generator=g)
    wi, bi = pack_geglu_weight(w, b, 256)
    out = capi.gemm(a.cuda().bfloat16(), wi.cuda().bfloat16(), bi.cuda(), act=capi.ACT_GEGLU, force_bn=256)
    assert report(f"geglu [{two_cta}]", out, ops.geglu(a, w, b))[0]


def test_groupnorm_combine_zero_weight_entries(capi):
    """tile-parallel shadow ranks contribute duplicate tiles with weight 0 (pixels = 0): they must drop out of the
    combine instead of producing 0/0 (found at 8 GPUs: 4 decoder tiles, 4 shadow ranks -> NaN image)."""
    g = _g(77)
    n_img, G, C = 2, 32, 128
    x = [bf16r(torch.randn(n_img, 12, 10 + 3 * i, C, generator=g)).cuda().bfloat16() for i in range(3)]
    sums = torch.cat([capi.groupnorm_stats(t, G) for t in x], 0)
    pix = torch.cat([torch.full((n_img,), float(t.shape[1] * t.shape[2]), device="cuda") for t in x], 0)
    ref = capi.groupnorm_combine(sums[:2 * n_img].contiguous(), pix[:2 * n_img].contiguous(), n_img, C)
    pix0 = pix.clone()
    pix0[2 * n_img:] = 0.0  # third tile group: weight 0
    got = capi.groupnorm_combine(sums, pix0, n_img, C)
    assert torch.isfinite(got).all() and torch.equal(got, ref)
    num, wsum = capi.groupnorm_combine(sums, pix0, n_img, C, raw=True)
    assert torch.isfinite(num).all() and torch.equal(wsum, pix[:n_img] + pix[n_img:2 * n_img])
    zero = torch.zeros_like(pix)  # a rank that holds ONLY shadows: all-zero numerators, not NaN
    num, wsum = capi.groupnorm_combine(sums, zero, n_img, C, raw=True)
    assert float(num.abs().max()) == 0.0 and float(wsum.abs().max()) == 0.0


ROLLING_CASES = [
    # B, H, W, Cin, Cout: B * ceil(ceil(W / 128) / 2) * H >= 74 * 48 pair-rows -> conv_halo_rr_kernel is eligible
    (2, 900, 384, 128, 128),    # three strips: the pair's second CTA idles on a dummy strip; H not a multiple of anything
    (4, 1024, 200, 128, 128),   # ragged x
    (4, 1000, 256, 64, 128),    # one K chunk: the A / B rings advance one unit per input row
    (2, 930, 500, 192, 64),     # three K chunks, thin Cout (zero-filled weight rows, masked stores)
]


@pytest.mark.parametrize("B,H,W,Cin,Cout", ROLLING_CASES)
@pytest.mark.parametrize("variant", ["plain", "res_stats", "gn_in", "gn_in_res_stats"])
def test_conv_halo_rolling_rows(capi, B, H, W, Cin, Cout, variant):
    """conv_halo_rr_kernel (each input row loaded / transformed once, four-slot accumulator ring) against the
    row-shared kernel on the same problem (bit-identical sums are not expected: the tap order differs) and against the
    fp32 oracle on a row band that includes a segment boundary, the image top and the image bottom."""
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight
    g = _g(B * 13 + H + W + Cin + Cout)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g) * 1.5 + 0.3)
    gamma, beta = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.2
    w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    b = torch.randn(Cout, generator=g)
    with_res = "res" in variant
    with_stats = "stats" in variant and Cout % 32 == 0 and Cout // 32 in (4, 8, 16)
    with_gn = variant.startswith("gn_in")
    r = bf16r(torch.randn(B, H, W, Cout, generator=g)) if with_res else None
    xc, wc, bc = x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda()
    rc = r.cuda().bfloat16() if with_res else None
    coef = None
    if with_gn:
        st = capi.groupnorm_stats(xc, 32)
        coef = capi.groupnorm_coef(st, gamma.cuda(), beta.cuda(), B, H * W, 32, 1e-6)

    def run(mode):
        capi.lib().omgsr_set_halo_rolling(mode)
        try:
            so = torch.empty(B, 32, 2, device="cuda") if with_stats else None
            if with_gn:
                y = capi.conv2d_gn(xc, coef, wc, bc, res=rc, gn_stats=so)
            else:
                y = capi.conv2d(xc, wc, bc, res=rc, gn_stats=so, cout=Cout)
            torch.cuda.synchronize()
            return y, so, capi.last_conv_variant()
        finally:
            capi.lib().omgsr_set_halo_rolling(-2)

    y_rr, s_rr, name_rr = run(1)
    y_rs, s_rs, name_rs = run(0)
    assert "conv_halo_rr_kernel" in name_rr and "conv_halo_rr_kernel" not in name_rs, (name_rr, name_rs)
    assert report(f"rolling vs row-shared [{variant}]", y_rr, y_rs.float().cpu(), 8e-3)[0]   # two bf16 steps at the largest value
    if with_stats:
        assert report("output statistics, rolling vs row-shared", s_rr, s_rs.cpu(), 1e-3)[0]
        f32 = y_rr.float()
        s_ref = torch.stack([f32.reshape(B, H * W, 32, Cout // 32).sum((1, 3)),
                             (f32 * f32).reshape(B, H * W, 32, Cout // 32).sum((1, 3))], -1)
        assert report("output statistics vs the stored tensor", s_rr, s_ref.cpu(), 1e-3)[0]
    # fp32 oracle on bands of rows (image top, bottom, and around the first cluster boundary G / 74)
    xin = bf16r(torch.nn.functional.silu(ops.groupnorm_nhwc(x, gamma, beta, 32, 1e-6))) if with_gn else x
    xp = ((W + 127) // 128 + 1) // 2
    cut = (B * xp * H // 74) % H
    for y0, y1 in ((0, 6), (H - 6, H), (max(cut - 4, 0), min(cut + 4, H))):
        lo, hi = max(y0 - 1, 0), min(y1 + 1, H)
        band = ops.conv2d_nhwc(xin[:, lo:hi], w, b)[:, y0 - lo:y0 - lo + (y1 - y0)]
        if with_res:
            band = band + r[:, y0:y1]
        assert report(f"rolling rows {y0}..{y1} vs oracle [{variant}]", y_rr[:, y0:y1], band)[0]


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(8, 256, 256, 256, 256), (4, 300, 384, 128, 256), (4, 512, 256, 512, 256)])
@pytest.mark.parametrize("variant", ["gn_in", "gn_in_res_stats"])
def test_conv_halo_rolling_rows_cout256(capi, B, H, W, Cin, Cout, variant):
    """Cout = 256 with the input GroupNorm fused: two 128-channel passes of conv_halo_rr_kernel over the same rows (the
    VAE's 256-channel level at 256^2 and above). Against the two-pass form (normalise, then the generic convolution) on
    the whole tensor and against the fp32 oracle on row bands; odd strip count, Cin = 128 / 256 / 512."""
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight
    assert capi.conv_gn_in_fusable(B, H, W, Cin, Cout)
    g = _g(B * 17 + H + W + Cin + Cout)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g) * 1.5 + 0.3)
    gamma, beta = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.2
    w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    b = torch.randn(Cout, generator=g)
    with_res = "res" in variant
    r = bf16r(torch.randn(B, H, W, Cout, generator=g)) if with_res else None
    xc, wc, bc = x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda()
    rc = r.cuda().bfloat16() if with_res else None
    st = capi.groupnorm_stats(xc, 32)
    coef = capi.groupnorm_coef(st, gamma.cuda(), beta.cuda(), B, H * W, 32, 1e-6)
    so = torch.empty(B, 32, 2, device="cuda") if "stats" in variant else None
    y = capi.conv2d_gn(xc, coef, wc, bc, res=rc, gn_stats=so)
    torch.cuda.synchronize()
    assert "conv_halo_rr_kernel<1>" in capi.last_conv_variant(), capi.last_conv_variant()
    two_pass = capi.conv2d(capi.groupnorm_apply(xc, st, gamma.cuda(), beta.cuda(), 32, 1e-6, True), wc, bc, res=rc)
    assert report(f"cout256 fused vs two-pass [{variant}]", y, two_pass.float().cpu(), 8e-3)[0]
    if so is not None:
        f32 = y.float()
        s_ref = torch.stack([f32.reshape(B, H * W, 32, Cout // 32).sum((1, 3)),
                             (f32 * f32).reshape(B, H * W, 32, Cout // 32).sum((1, 3))], -1)
        assert report("cout256 output statistics", so, s_ref.cpu(), 1e-3)[0]
    xin = bf16r(torch.nn.functional.silu(ops.groupnorm_nhwc(x, gamma, beta, 32, 1e-6)))
    for y0, y1 in ((0, 4), (H - 4, H), (H // 2 - 2, H // 2 + 2)):
        lo, hi = max(y0 - 1, 0), min(y1 + 1, H)
        band = ops.conv2d_nhwc(xin[:, lo:hi], w, b)[:, y0 - lo:y0 - lo + (y1 - y0)]
        if with_res:
            band = band + r[:, y0:y1]
        assert report(f"cout256 rows {y0}..{y1} vs oracle [{variant}]", y[:, y0:y1], band)[0]


@pytest.mark.parametrize("B,H,W,Cin,Cout,res", [(2, 8, 128, 64, 128, False), (1, 33, 200, 128, 128, True),
                                                (2, 40, 256, 256, 128, True), (1, 7, 130, 64, 64, False),
                                                (3, 10, 300, 192, 128, False)])
def test_conv_fused_input_groupnorm(capi, B, H, W, Cin, Cout, res):
    """conv3x3(silu(groupnorm(x))) with the normalisation applied to the operand tile in shared memory
    (omgsr_conv2d_gn_nhwc_bf16) == the two-pass form, and == the fp32 oracle. Covers ragged x, odd row counts, an odd
    number of tiles (dummy CTA of the pair) and image borders (halo pixels must stay zero AFTER the transform)."""
    from oracle import ops
    from omgsr_b200.pack import pack_conv_weight
    g = _g(B * 11 + H + W + Cin + Cout)
    x = bf16r(torch.randn(B, H, W, Cin, generator=g) * 1.5 + 0.3)
    gamma, beta = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.2
    w = bf16r(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    b = torch.randn(Cout, generator=g)
    r = bf16r(torch.randn(B, H, W, Cout, generator=g)) if res else None
    assert capi.conv_gn_in_fusable(B, H, W, Cin, Cout)
    xc, wc, bc = x.cuda().bfloat16(), pack_conv_weight(w).cuda(), b.cuda()
    gc, btc = gamma.cuda(), beta.cuda()
    rc = r.cuda().bfloat16() if res else None
    st = capi.groupnorm_stats(xc, 32)
    two_pass = capi.conv2d(capi.groupnorm_apply(xc, st, gc, btc, 32, 1e-6, True), wc, bc, res=rc)
    coef = capi.groupnorm_coef(st, gc, btc, B, H * W, 32, 1e-6)
    stats_out = torch.empty(B, 32, 2, device="cuda") if Cout % 32 == 0 and Cout // 32 in (4, 8, 16) else None
    fused = capi.conv2d_gn(xc, coef, wc, bc, res=rc, gn_stats=stats_out)
    torch.cuda.synchronize()
    assert report("fused vs two-pass", fused, two_pass.float().cpu(), 2e-3)[0]
    ref = ops.conv2d_nhwc(bf16r(torch.nn.functional.silu(ops.groupnorm_nhwc(x, gamma, beta, 32, 1e-6))), w, b)
    if res:
        ref = ref + r
    assert report(f"conv(silu(gn(x))) {B}x{H}x{W} {Cin}->{Cout}", fused, ref)[0]
    if stats_out is not None:
        f32 = fused.float()
        s_ref = torch.stack([f32.reshape(B, H * W, 32, Cout // 32).sum((1, 3)),
                             (f32 * f32).reshape(B, H * W, 32, Cout // 32).sum((1, 3))], -1)
        assert report("fused output statistics", stats_out, s_ref.cpu(), 1e-3)[0]
