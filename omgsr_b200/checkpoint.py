"""On-disk formats at load time: diffusers checkpoint directories (`*.safetensors`, one or
several shards per sub-folder) and PEFT adapter directories (`adapter_config.json` + `adapter_model.safetensors`,
keys `base_model.model.<module path>.lora_{A,B}[.<adapter>].weight`) are read WITHOUT diffusers / peft /
safetensors being installed, straight into the diffusers-keyed state dicts the engines repack.

Reference call sites replaced: AutoencoderKL / UNet2DConditionModel / FluxTransformer2DModel.from_pretrained and
PeftModel.from_pretrained + merge_and_unload (infer/omgsr_s_infer_model.py:11-23, infer/omgsr_f_infer_model.py:99-118);
writer side of the adapter format: train/train_omgsr_s.py:516-517.
"""
import json
import os
import struct

import numpy as np
import torch

from . import weights as W

_DTYPES = {"F32": (torch.float32, np.float32), "F16": (torch.float16, np.float16), "BF16": (torch.bfloat16, np.uint16),
           "F64": (torch.float64, np.float64), "I64": (torch.int64, np.int64), "I32": (torch.int32, np.int32),
           "U8": (torch.uint8, np.uint8), "BOOL": (torch.bool, np.bool_)}


def read_safetensors(path):
    """minimal safetensors reader: 8-byte little-endian header length, JSON header, raw little-endian data."""
    out = {}
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n).decode("utf-8"))
        base = 8 + n
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            tdt, ndt = _DTYPES[meta["dtype"]]
            lo, hi = meta["data_offsets"]
            f.seek(base + lo)
            raw = np.frombuffer(f.read(hi - lo), dtype=ndt).copy()
            t = torch.from_numpy(raw)
            if meta["dtype"] == "BF16":
                t = t.view(torch.bfloat16)
            out[name] = t.reshape(meta["shape"])
    return out


def write_safetensors(path, tensors):
    """writer counterpart (used by the tests and to export synthetic checkpoints)."""
    rev = {v[0]: k for k, v in _DTYPES.items()}
    header, blobs, off = {}, [], 0
    for name, t in tensors.items():
        t = t.detach().cpu().contiguous()
        raw = (t.view(torch.uint16) if t.dtype == torch.bfloat16 else t).numpy().tobytes()
        header[name] = {"dtype": rev[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def _load_folder(folder):
    """all *.safetensors shards of a diffusers sub-folder merged into one dict (+ its config.json)."""
    sd = {}
    files = sorted(fn for fn in os.listdir(folder) if fn.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors file in {folder}")
    for fn in files:
        sd.update(read_safetensors(os.path.join(folder, fn)))
    cfg_path = os.path.join(folder, "config.json")
    cfg = json.load(open(cfg_path)) if os.path.exists(cfg_path) else {}
    return sd, cfg


def _vae_cfg(c, default):
    if not c:
        return default
    return W.VAEConfig(in_channels=c.get("in_channels", 3), out_channels=c.get("out_channels", 3),
                       latent_channels=c.get("latent_channels", default.latent_channels),
                       block_out_channels=tuple(c.get("block_out_channels", default.block_out_channels)),
                       layers_per_block=c.get("layers_per_block", 2),
                       scaling_factor=c.get("scaling_factor", default.scaling_factor),
                       shift_factor=c.get("shift_factor") or 0.0,
                       use_quant_conv=c.get("use_quant_conv", default.use_quant_conv))


def _normalise_vae_keys(sd):
    """older SD VAE checkpoints name the mid-attention projections query/key/value/proj_attn (as 1x1 convs)."""
    ren = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts and parts[-2] in ren:
            parts[-2] = ren[parts[-2]]
            k = ".".join(parts)
            if v.dim() == 4:
                v = v.reshape(v.shape[0], v.shape[1])
        out[k] = v
    return out


def load_sd_checkpoint(sd_path):
    """-> (vae_sd, unet_sd, vae_cfg, unet_cfg) from a diffusers SD-2.1 directory (vae/, unet/)."""
    vae_sd, vc = _load_folder(os.path.join(sd_path, "vae"))
    unet_sd, uc = _load_folder(os.path.join(sd_path, "unet"))
    vae_cfg = _vae_cfg(vc, W.SD21_VAE)
    unet_cfg = W.SD21_UNET
    if uc:
        heads = uc.get("attention_head_dim", list(W.SD21_UNET.num_heads))
        ch = tuple(uc.get("block_out_channels", W.SD21_UNET.block_out_channels))
        heads = tuple(heads) if isinstance(heads, (list, tuple)) else (heads,) * len(ch)
        unet_cfg = W.UNetConfig(in_channels=uc.get("in_channels", 4), out_channels=uc.get("out_channels", 4),
                                block_out_channels=ch, num_heads=heads,
                                down_has_attn=tuple("CrossAttn" in t for t in uc.get("down_block_types", [])) or
                                W.SD21_UNET.down_has_attn,
                                layers_per_block=uc.get("layers_per_block", 2),
                                cross_attention_dim=uc.get("cross_attention_dim", 1024))
    return _normalise_vae_keys(vae_sd), unet_sd, vae_cfg, unet_cfg


def load_flux_checkpoint(flux_path):
    """-> (vae_sd, flux_sd, vae_cfg, flux_cfg) from a diffusers FLUX.1-dev directory (vae/, transformer/)."""
    vae_sd, vc = _load_folder(os.path.join(flux_path, "vae"))
    flux_sd, fc = _load_folder(os.path.join(flux_path, "transformer"))
    flux_cfg = W.FLUX_DEV
    if fc:
        flux_cfg = W.FluxConfig(in_channels=fc.get("in_channels", 64), num_layers=fc.get("num_layers", 19),
                                num_single_layers=fc.get("num_single_layers", 38),
                                num_heads=fc.get("num_attention_heads", 24), head_dim=fc.get("attention_head_dim", 128),
                                joint_attention_dim=fc.get("joint_attention_dim", 4096),
                                pooled_projection_dim=fc.get("pooled_projection_dim", 768),
                                axes_dims_rope=tuple(fc.get("axes_dims_rope", (16, 56, 56))),
                                guidance_embeds=fc.get("guidance_embeds", True))
    return _normalise_vae_keys(vae_sd), flux_sd, _vae_cfg(vc, W.FLUX_VAE), flux_cfg


def load_peft_adapter(adapter_dir, prefix=""):
    """-> ({module path: (A, B)}, scale) folded to {module: (A, scale * B)} so that merge_lora(scale=1) applies
    W <- W + (lora_alpha / r) * B A. `prefix` is prepended to the module paths (the VAE adapter is saved relative
    to vae.encoder, infer/omgsr_s_infer_model.py:16-18)."""
    cfg = json.load(open(os.path.join(adapter_dir, "adapter_config.json")))
    sd = read_safetensors(os.path.join(adapter_dir, "adapter_model.safetensors"))
    alpha, r = cfg.get("lora_alpha", cfg.get("r", 1)), cfg.get("r", 1)
    scale = alpha / r
    a_map, b_map = {}, {}
    for k, v in sd.items():
        parts = k.split(".")
        if parts[:2] == ["base_model", "model"]:
            parts = parts[2:]
        if "lora_A" in parts:
            i = parts.index("lora_A")
            a_map[prefix + ".".join(parts[:i])] = v.float()
        elif "lora_B" in parts:
            i = parts.index("lora_B")
            b_map[prefix + ".".join(parts[:i])] = v.float()
    return {m: (a_map[m], scale * b_map[m]) for m in a_map if m in b_map}
