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


def preprocess_gpu(img, process_size, upscale, device, dtype):
    """the same chain as `preprocess` + `pil_to_model_input`, on the device: the decoded image is uploaded once as
    uint8 and resampled by kernels that reproduce Pillow's fixed-point resampler bit for bit (omgsr_resize_u8).
    -> (lq_u8 [H, W, 3] uint8 CUDA tensor on the output grid, model input [1, 3, H, W], resize_flag, original size)"""
    from omgsr_b200 import capi
    ow, oh = img.size
    x = torch.from_numpy(np.asarray(img.convert("RGB"), dtype=np.uint8).copy()).to(device)
    resize_flag = False
    w, h = ow, oh
    if ow < process_size // upscale or oh < process_size // upscale:
        s = (process_size // upscale) / min(ow, oh)
        w, h = int(s * ow), int(s * oh)
        x = capi.resize_u8(x, (h, w), capi.RESIZE_BICUBIC)
        resize_flag = True
    w, h = w * upscale, h * upscale
    x = capi.resize_u8(x, (h, w), capi.RESIZE_BICUBIC)
    nw, nh = w - w % 8, h - h % 8
    x = capi.resize_u8(x, (nh, nw), capi.RESIZE_LANCZOS)
    return x, capi.u8_to_input(x, dtype), resize_flag, (ow, oh)


def pil_to_model_input(img, device, dtype):
    arr = np.asarray(img.convert("RGB"), dtype=np.uint8)
    x = torch.from_numpy(arr.copy()).to(device).permute(2, 0, 1).unsqueeze(0)
    return (x.float() / 255.0).to(dtype) * 2 - 1


def finish(output, lq_pil, align_method, resize_flag, orig_size, upscale, save_path):
    """de-normalise, clip, colour-align against the LQ image and save. The colour fix runs on the GPU
    (infer/wavelet_color_fix.py) with the reference's uint8 truncation semantics."""
    from .wavelet_color_fix import _cuda_to_pil, adain_color_fix, wavelet_color_fix
    if torch.is_tensor(lq_pil):  # GPU pre-processing path: the LQ image is a uint8 [H, W, 3] device tensor
        lq_pil = Image.fromarray(lq_pil.cpu().numpy(), mode="RGB")
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


def finish_gpu(output, lq_u8, align_method, resize_flag, orig_size, upscale, save_path):
    """`finish` without the host round trips: the model output is quantised to uint8 on the device with the
    reference's truncation (omgsr_output_to_u8), colour-aligned against the uint8 LQ tensor from `preprocess_gpu`
    (same values as to_tensor of the PIL images the reference feeds its colour fix), and crosses PCIe once as uint8."""
    from omgsr_b200 import capi
    from .wavelet_color_fix import adaptive_instance_normalization, wavelet_reconstruction
    out_u8 = capi.output_to_u8(output[:1].contiguous())
    if align_method in ("adain", "wavelet"):
        tgt = out_u8.permute(2, 0, 1).float().div_(255.0).unsqueeze(0).contiguous()
        src = lq_u8.permute(2, 0, 1).float().div_(255.0).unsqueeze(0).contiguous()
        fixed = adaptive_instance_normalization(tgt, src) if align_method == "adain" else wavelet_reconstruction(tgt, src)
        out_u8 = fixed.squeeze(0).clamp_(0.0, 1.0).mul(255).byte().permute(1, 2, 0).contiguous()
    if resize_flag:
        out_u8 = capi.resize_u8(out_u8, (int(upscale * orig_size[1]), int(upscale * orig_size[0])), capi.RESIZE_BICUBIC)
    pil = Image.fromarray(out_u8.cpu().numpy(), mode="RGB")
    if save_path:
        pil.save(save_path)
    return pil


def run_batched(items, run_model, args, device, dtype):
    """Driver loop with batching: images whose model-input size agrees are stacked
    up to --batch_size and go through ONE forward; pre/post-processing stays per image on the device and the PNG
    encodes run on a worker thread so that they overlap the next batch. `run_model(batch [B,3,H,W]) -> (out, secs)`.
    Results are identical to the one-image-at-a-time loop (same kernels per image; batch-invariant arithmetic
    except for the GroupNorm atomics' summation order)."""
    import concurrent.futures as cf
    total_time, n_done = 0.0, 0
    pool = cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2))  # PNG encoding releases the GIL in zlib
    saves = []

    def flush(batch):
        nonlocal total_time, n_done
        if not batch:
            return
        x = torch.cat([b[2] for b in batch], dim=0)
        with torch.no_grad():
            out, t = run_model(x)
        total_time += t
        for i, (bname, lq_u8, _x, flag, orig) in enumerate(batch):
            pil = finish_gpu(out[i:i + 1], lq_u8, args.align_method, flag, orig, args.upscale, None)
            saves.append(pool.submit(pil.save, os.path.join(args.output_dir, bname)))
            n_done += 1

    batch = []
    for bname, img in items:
        lq_u8, x, flag, orig = preprocess_gpu(img, args.process_size, args.upscale, device, dtype)
        if batch and (batch[0][2].shape != x.shape or len(batch) >= args.batch_size):
            flush(batch)
            batch = []
        batch.append((bname, lq_u8, x, flag, orig))
    flush(batch)
    for f in saves:
        f.result()
    pool.shutdown()
    return total_time, n_done
