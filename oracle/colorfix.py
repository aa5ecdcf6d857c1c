"""Oracle for the colour fix (TEST INFRASTRUCTURE ONLY).

Restates infer/wavelet_color_fix.py of the reference: AdaIN (:44-74) and the 5-level a-trous wavelet
reconstruction (:77-125), plus the PIL wrappers (:12-41). Pinned against the reference's own functions by
tests/golden/colorfix_*.pt (tests/test_oracle_cpu.py)."""
import numpy as np
import torch
import torch.nn.functional as F

_BINOMIAL = torch.tensor([1.0, 2.0, 1.0]) / 4.0


def _mean_std(feat, eps=1e-5):
    # infer/wavelet_color_fix.py:44-57 — unbiased variance over H*W, + eps, sqrt
    b, c = feat.shape[:2]
    flat = feat.reshape(b, c, -1)
    std = (flat.var(dim=2) + eps).sqrt().reshape(b, c, 1, 1)
    mean = flat.mean(dim=2).reshape(b, c, 1, 1)
    return mean, std


def adaptive_instance_normalization(content, style):
    # infer/wavelet_color_fix.py:60-74
    ms, ss = _mean_std(style)
    mc, sc = _mean_std(content)
    return (content - mc) / sc * ss + ms


def atrous_blur(img, radius):
    # infer/wavelet_color_fix.py:77-96 — depthwise [1,2,1]x[1,2,1]/16, dilation = radius, replicate pad
    c = img.shape[1]
    k2 = torch.outer(_BINOMIAL, _BINOMIAL).to(img.dtype)
    k = k2[None, None].repeat(c, 1, 1, 1)
    padded = F.pad(img, (radius,) * 4, mode="replicate")
    return F.conv2d(padded, k, groups=c, dilation=radius)


def wavelet_decomposition(img, levels=5):
    # infer/wavelet_color_fix.py:99-111
    high = torch.zeros_like(img)
    low = img
    for lvl in range(levels):
        nxt = atrous_blur(low, 2 ** lvl)
        high = high + (low - nxt)
        low = nxt
    return high, low


def wavelet_reconstruction(content, style):
    # infer/wavelet_color_fix.py:114-125 — content detail + style base
    high, _ = wavelet_decomposition(content)
    _, low = wavelet_decomposition(style)
    return high + low


def _pil_to_tensor(img):
    # torchvision ToTensor: uint8 HWC -> float CHW / 255
    arr = np.asarray(img.convert("RGB"), dtype=np.uint8)
    return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div(255.0).unsqueeze(0)


def _tensor_to_pil(t):
    # torchvision ToPILImage on a float tensor: mul(255).byte() (truncation), CHW -> HWC
    from PIL import Image
    arr = t.squeeze(0).clamp(0.0, 1.0).mul(255).byte().permute(1, 2, 0).numpy()
    return Image.fromarray(arr, mode="RGB")


def adain_color_fix(target, source):
    # infer/wavelet_color_fix.py:12-25
    return _tensor_to_pil(adaptive_instance_normalization(_pil_to_tensor(target), _pil_to_tensor(source)))


def wavelet_color_fix(target, source):
    # infer/wavelet_color_fix.py:28-41
    return _tensor_to_pil(wavelet_reconstruction(_pil_to_tensor(target), _pil_to_tensor(source)))
