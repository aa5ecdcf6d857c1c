"""Weight repacking from the diffusers state-dict layouts into the layouts the sm_100a kernels read.

All functions are pure tensor reshapes done once at load time (after the LoRA merge,
infer/omgsr_s_infer_model.py:16-25): no arithmetic of the hot path happens here except the exact algebraic
folds documented per function.
"""
import torch


def pack_conv_weight(w, dtype=torch.bfloat16):
    """OIHW [Cout, Cin, R, S] -> [Cout, R*S*Cin] with k = (r*S + s)*Cin + cin (K-major B operand)."""
    co, ci, r, s = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, r * s * ci).contiguous().to(dtype)


def pack_upsample_weight(w, dtype=torch.bfloat16):
    """Collapse nearest-2x-upsample + 3x3/p1 conv into four 2x2 convs on the un-upsampled input.

    Output pixel (2i+py, 2j+px) reads upsampled rows 2i+py+r-1 (r = 0..2) = source rows floor((2i+py+r-1)/2):
      py = 0: r=0 -> i-1, r=1,2 -> i        => taps (W0, W1+W2) at source rows (i-1, i), pad_top = 1
      py = 1: r=0,1 -> i, r=2 -> i+1        => taps (W0+W1, W2) at source rows (i, i+1), pad_top = 0
    and likewise along x. Sums are formed in fp32 and rounded once. Returns [4, Cout, 4*Cin] with
    parity index py*2+px and k = (a*2 + b)*Cin + cin. 2.25x fewer MACs than convolving the 4x larger tensor.
    """
    co, ci, r, s = w.shape
    assert r == 3 and s == 3
    w = w.float()
    rows = {0: [w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 1: [w[:, :, 0] + w[:, :, 1], w[:, :, 2]]}  # [Cout,Cin,3(s)]
    out = torch.empty(4, co, 4 * ci, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for a in (0, 1):
                ra = rows[py][a]
                cols = [ra[:, :, 0], ra[:, :, 1] + ra[:, :, 2]] if px == 0 else [ra[:, :, 0] + ra[:, :, 1], ra[:, :, 2]]
                taps.extend(cols)  # (a,0), (a,1)
            out[py * 2 + px] = torch.stack(taps, dim=1).reshape(co, 4 * ci)
    return out.contiguous().to(dtype)


def pack_geglu_weight(w, b, bn=256):
    """GEGLU projection [2*F, K] (first F rows = value, last F = gate; diffusers GEGLU.chunk(2)) ->
    rows interleaved per `bn`-wide output tile: [bn/2 value rows | bn/2 gate rows] so the GEMM epilogue
    finds v and g of the same channel in one accumulator tile."""
    f2, k = w.shape
    f = f2 // 2
    half = bn // 2
    assert f % half == 0
    v = w[:f].reshape(f // half, half, k)
    g = w[f:].reshape(f // half, half, k)
    wi = torch.cat([v, g], dim=1).reshape(f2, k).contiguous()
    bi = None
    if b is not None:
        bv = b[:f].reshape(f // half, half)
        bg = b[f:].reshape(f // half, half)
        bi = torch.cat([bv, bg], dim=1).reshape(f2).contiguous()
    return wi, bi


def pack_conv_in_weight(w):
    """OIHW [Cout, Cin, R, S] -> fp32 [R*S*Cin, Cout8] for the thin conv_in kernel (Cout zero-padded to a
    multiple of 8: the kernel computes 8 output channels per thread)."""
    co, ci, r, s = w.shape
    co8 = (co + 7) // 8 * 8
    out = torch.zeros(r * s * ci, co8, dtype=torch.float32)
    out[:, :co] = w.permute(2, 3, 1, 0).reshape(r * s * ci, co).float()
    return out.contiguous()


def pad8(v):
    """zero-pad a 1-D fp32 vector to a multiple of 8 entries."""
    n8 = (v.numel() + 7) // 8 * 8
    out = torch.zeros(n8, dtype=torch.float32)
    out[: v.numel()] = v.float()
    return out


def pack_conv_in_gemm_weight(w, dtype=torch.bfloat16):
    """OIHW [Cout, Cin, 3, 3] -> [Cout, Kpad] with k = (r*3+s)*Cin + ci, zero padded to a multiple of 32
    (B operand of the im2col + GEMM form of a thin-input conv; the GEMM's TMA boxes zero-fill up to the next 64, so
    a 3-channel conv_in streams 32 instead of 64 patch columns per pixel)."""
    co, ci, r, s = w.shape
    assert r == 3 and s == 3
    k = 9 * ci
    kpad = (k + 31) // 32 * 32
    out = torch.zeros(co, kpad, dtype=torch.float32)
    out[:, :k] = w.permute(0, 2, 3, 1).reshape(co, k).float()
    return out.contiguous().to(dtype)
