"""Generates tests/golden/*.pt by running the REFERENCE's own importable code on CPU.

Run in the authoring container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
What is importable without diffusers/peft: infer/wavelet_color_fix.py (everything),
infer/vaehook.py (split_tiles, get_best_tile_size, crop_valid_region, get_var_mean, custom_group_norm),
and the pure-math blocks of infer/omgsr_f_infer_model.py:21-76 and infer/omgsr_s_infer_model.py:56-71,
infer/infer_omgsr_f.py:17-28 (taken by source slice because the module imports peft at the top).
The fixtures pin the oracle (tests/test_oracle_cpu.py) and, through it, the CUDA path."""
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _slice_exec(path, start, end, extra_globals=None):
    lines = open(path).read().split("\n")[start - 1:end]
    g = {"torch": torch, "math": math, "np": np, "Callable": __import__("typing").Callable}
    if extra_globals:
        g.update(extra_globals)
    exec("\n".join(lines), g)
    return g


def main():
    # bind `infer` to the reference directory (it is a namespace package there; this repo's own `infer`
    # package would otherwise shadow it)
    pkg = types.ModuleType("infer")
    pkg.__path__ = [os.path.join(REF, "infer")]
    sys.modules["infer"] = pkg
    import infer.wavelet_color_fix as wcf
    import infer.vaehook as vaehook

    torch.manual_seed(1234)
    gold = {}
    # ---- colour fix ----
    c = torch.rand(1, 3, 48, 56)
    s = torch.rand(1, 3, 48, 56)
    gold["colorfix"] = {
        "content": c, "style": s,
        "wavelet": wcf.wavelet_reconstruction(c.clone(), s.clone()),
        "adain": wcf.adaptive_instance_normalization(c.clone(), s.clone()),
        "blur_r4": wcf.wavelet_blur(c.clone(), 4),
    }
    from PIL import Image
    rng = np.random.default_rng(1234)
    tgt = Image.fromarray(rng.integers(0, 256, (40, 36, 3), dtype=np.uint8), "RGB")
    src = Image.fromarray(rng.integers(0, 256, (40, 36, 3), dtype=np.uint8), "RGB")
    gold["colorfix_pil"] = {
        "target": torch.from_numpy(np.asarray(tgt).copy()), "source": torch.from_numpy(np.asarray(src).copy()),
        "wavelet": torch.from_numpy(np.asarray(wcf.wavelet_color_fix(tgt, src)).copy()),
        "adain": torch.from_numpy(np.asarray(wcf.adain_color_fix(tgt, src)).copy()),
    }
    # ---- vaehook tile planner + GN helpers ----
    plans = {}
    for (is_dec, tile, h, w) in [(False, 256, 1024, 1024), (True, 256, 128, 128), (True, 224, 256, 256),
                                 (False, 1024, 2048, 2048), (True, 64, 128, 128), (False, 512, 1000, 1512),
                                 (True, 64, 96, 160), (False, 256, 320, 704)]:
        hook = vaehook.VAEHook(None, tile, is_dec, False, False, False)
        ib, ob = hook.split_tiles(h, w)
        plans[(is_dec, tile, h, w)] = (ib, ob)
    gold["split_tiles"] = plans
    hook = vaehook.VAEHook(None, 256, False, False, False, False)
    gold["best_tile"] = {(lo, hi): hook.get_best_tile_size(lo, hi) for lo in (1, 31, 100, 150, 224, 240, 255, 256)
                         for hi in (224, 256, 512)}
    x = torch.randn(2, 64, 10, 12)
    var, mean = vaehook.get_var_mean(x, 32)
    wgt, bia = torch.randn(64), torch.randn(64)
    gold["groupnorm"] = {"x": x, "var": var, "mean": mean, "w": wgt, "b": bia,
                         "out": vaehook.custom_group_norm(x.clone(), 32, mean, var, wgt, bia)}
    t = torch.randn(1, 3, 40, 48)
    gold["crop"] = {"x": t, "dec": vaehook.crop_valid_region(t, [0, 6, 0, 5], [0, 37, 0, 29], True),
                    "enc": vaehook.crop_valid_region(t, [64, 448, 0, 320], [12, 52, 0, 36], False)}
    # ---- FLUX helpers (source slices: the module itself imports peft/diffusers) ----
    gf = _slice_exec(os.path.join(REF, "infer/omgsr_f_infer_model.py"), 21, 76)
    sched = gf["get_flux_setting_timesteps"]()
    lat = torch.randn(2, 16, 8, 12)
    packed = gf["_pack_latents"](lat, 2, 16, 8, 12)
    gold["flux"] = {"schedule": torch.tensor(sched, dtype=torch.float64), "t_curr_244": sched[-(244 + 1)],
                    "lat": lat, "packed": packed, "unpacked": gf["_unpack_latents"](packed, 64, 96, 8)}
    gi = _slice_exec(os.path.join(REF, "infer/infer_omgsr_f.py"), 17, 28)
    gold["flux"]["img_ids"] = gi["_prepare_latent_image_ids"](6, 4, "cpu", torch.float32)
    # ---- S helpers: _gaussian_weights as a free function over a stub self ----
    src = open(os.path.join(REF, "infer/omgsr_s_infer_model.py")).read().split("\n")[55:71]
    body = "\n".join(l[4:] for l in src)
    g = {"torch": torch}
    exec(body, g)
    stub = types.SimpleNamespace(device="cpu", unet=types.SimpleNamespace(config=types.SimpleNamespace(in_channels=4)))
    gold["gauss"] = {(tw, th): g["_gaussian_weights"](stub, tw, th, 1) for (tw, th) in [(64, 64), (8, 8), (16, 12)]}
    torch.save(gold, os.path.join(OUT, "reference_fixtures.pt"))
    print("wrote", os.path.join(OUT, "reference_fixtures.pt"), os.path.getsize(os.path.join(OUT, "reference_fixtures.pt")))


if __name__ == "__main__":
    main()
