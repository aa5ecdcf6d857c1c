"""Operator-level oracle: fp32 CPU restatements of the primitive ops the nets are made of.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import math

import torch
import torch.nn.functional as F


def conv2d_nhwc(x, w, bias=None, stride=1, pad=(1, 1, 1, 1), upsample2x=False):
    """x [B,H,W,Cin], w OIHW; pad = (top, left, bottom, right). Mirrors F.pad + Conv2d in diffusers'
    ResnetBlock2D / Downsample2D (asymmetric (0,1,0,1) pad) / Upsample2D (nearest 2x then conv)."""
    xn = x.float().permute(0, 3, 1, 2)
    if upsample2x:
        xn = F.interpolate(xn, scale_factor=2.0, mode="nearest")
    xn = F.pad(xn, (pad[1], pad[3], pad[0], pad[2]))
    y = F.conv2d(xn, w.float(), None if bias is None else bias.float(), stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def linear(x, w, bias=None):
    return F.linear(x.float(), w.float(), None if bias is None else bias.float())


def groupnorm_nhwc(x, gamma, beta, groups=32, eps=1e-6, silu=False):
    xn = x.float().permute(0, 3, 1, 2) if x.dim() == 4 else x.float().transpose(1, 2)
    y = F.group_norm(xn, groups, gamma.float(), beta.float(), eps)
    if silu:
        y = F.silu(y)
    return (y.permute(0, 2, 3, 1) if x.dim() == 4 else y.transpose(1, 2)).contiguous()


def layernorm(x, gamma=None, beta=None, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), None if gamma is None else gamma.float(),
                        None if beta is None else beta.float(), eps)


def attention(q, k, v, scale=None):
    """q [B,M,H,D], k/v [B,N,H,D] -> [B,M,H,D] (softmax(q k^T * scale) v in fp32)."""
    q, k, v = q.float(), k.float(), v.float()
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    s = torch.einsum("bmhd,bnhd->bhmn", q, k) * scale
    p = s.softmax(dim=-1)
    return torch.einsum("bhmn,bnhd->bmhd", p, v)


def geglu(x, w, bias):
    """diffusers GEGLU: hidden, gate = proj(x).chunk(2); hidden * gelu(gate) (erf form)."""
    h = linear(x, w, bias)
    a, g = h.chunk(2, dim=-1)
    return a * F.gelu(g)
