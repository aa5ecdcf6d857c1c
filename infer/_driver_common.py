"""Shared pieces of the two CLI drivers (infer_omgsr_s.py / infer_omgsr_f.py): file discovery, the PIL
pre-processing chain of the reference (infer/infer_omgsr_s.py:69-88), colour alignment and saving (:96-107)."""
import glob
import os

import numpy as np
import torch
from PIL import Image


def list_inputs(spec):
    """a .txt list, a directory of png/jpg/jpeg (sorted per extension like the reference), or one file."""
    if ".txt" in spec:
        with open(spec, "r") as f:
            return [line.strip() for line in f.readlines()]
    if os.path.isdir(spec):
        return sorted(glob.glob(f"{spec}/*.png") + glob.glob(f"{spec}/*.jpg") + glob.glob(f"{spec}/*.jpeg"))
    return [spec]


def synthetic_images(n, size, seed=1234):
    """seeded low-passed noise LQ images (no dataset is reachable offline)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        small = rng.integers(0, 256, (size // 4, size // 4, 3), dtype=np.uint8)
        out.append((f"synthetic_{i:03d}.png", Image.fromarray(small, "RGB").resize((size, size), Image.BICUBIC)))
    return out


def preprocess(img, process_size, upscale):
    """-> (LQ image upsampled to the output grid and cropped to a multiple of 8, resize_flag, original size).
    Same chain as the reference: optional min-side upscale, x`upscale` (PIL default filter), LANCZOS to /8."""
    ow, oh = img.size
    resize_flag = False
    if ow < process_size // upscale or oh < process_size // upscale:
        s = (process_size // upscale) / min(ow, oh)
        img = img.resize((int(s * ow), int(s * oh)))
        resize_flag = True
    img = img.resize((img.size[0] * upscale, img.size[1] * upscale))
    nw, nh = img.width - img.width % 8, img.height - img.height % 8
    return img.resize((nw, nh), Image.LANCZOS), resize_flag, (ow, oh)


def pil_to_model_input(img, device, dtype):
    arr = np.asarray(img.convert("RGB"), dtype=np.uint8)
    x = torch.from_numpy(arr.copy()).to(device).permute(2, 0, 1).unsqueeze(0)
    return (x.float() / 255.0).to(dtype) * 2 - 1


def finish(output, lq_pil, align_method, resize_flag, orig_size, upscale, save_path):
    """de-normalise, clip, colour-align against the LQ image and save. The colour fix runs on the GPU
    (infer/wavelet_color_fix.py) with the reference's uint8 truncation semantics."""
    from .wavelet_color_fix import _cuda_to_pil, adain_color_fix, wavelet_color_fix
    out = torch.clip(output * 0.5 + 0.5, 0, 1).float()
    pil = _cuda_to_pil(out[:1].clone())
    if align_method == "adain":
        pil = adain_color_fix(target=pil, source=lq_pil)
    elif align_method == "wavelet":
        pil = wavelet_color_fix(target=pil, source=lq_pil)
    if resize_flag:
        pil = pil.resize((int(upscale * orig_size[0]), int(upscale * orig_size[1])))
    if save_path:
        pil.save(save_path)
    return pil
