"""BASELINE config 4 on one GPU: OMGSR-F-1024 (FLUX.1-dev shapes, random-init + merged LoRA generated on the
device), 256->1024, batch B per GPU. Prints stage timings and images/s (CUDA events)."""
import contextlib, io, json, sys, time
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from infer.omgsr_f_infer_model import OMGSR_F_Infer, encode_images, _pack_latents
from infer.infer_omgsr_f import _prepare_latent_image_ids
from omgsr_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
t0 = time.time()
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_F_Infer("random://flux_dev?seed=0", "random://lora?seed=1&rank=64", dev)
torch.cuda.synchronize()
print(f"model build {time.time()-t0:.1f}s, mem {torch.cuda.memory_allocated()/2**30:.1f} GiB", flush=True)
g = torch.Generator().manual_seed(1234)
lq = (torch.rand(B, 3, 1024, 1024, generator=g) * 2 - 1).to(dev).bfloat16()
pe = torch.randn(1, 512, 4096, generator=g).to(dev).bfloat16()
pp = torch.randn(1, 768, generator=g).to(dev).bfloat16()
tid = torch.zeros(512, 3, device=dev, dtype=torch.bfloat16)
ids = _prepare_latent_image_ids(64, 64, dev, torch.bfloat16)
def ev(): return torch.cuda.Event(enable_timing=True)
res = {}
with torch.no_grad():
    for it in range(1 + steps):
        e = [ev() for _ in range(4)]
        l0 = capi.launch_count()
        e[0].record()
        z = encode_images(lq, m.vae, torch.bfloat16)
        e[1].record()
        packed = _pack_latents(z, B, 16, 128, 128)
        v = m._dit(packed, pe, pp, tid, ids)
        e[2].record()
        packed = capi.latent_axpby(packed.contiguous(), v.contiguous(), 1.0, m.t_prev - m.t_curr)
        from infer.omgsr_f_infer_model import _unpack_latents
        lat = _unpack_latents(packed, 1024, 1024, 8)
        lat = capi.latent_axpby(lat, None, 1.0 / 0.3611, 0.0, 0.1159)
        img = m.vae.decode(lat, return_dict=False)[0]
        e[3].record()
        torch.cuda.synchronize()
        if it > 0:
            res.setdefault("enc", []).append(e[0].elapsed_time(e[1]))
            res.setdefault("dit", []).append(e[1].elapsed_time(e[2]))
            res.setdefault("dec", []).append(e[2].elapsed_time(e[3]))
            res.setdefault("launches", []).append(capi.launch_count() - l0)
enc, dit, dec = (sum(res[k]) / len(res[k]) for k in ("enc", "dit", "dec"))
tot = enc + dit + dec
out = {"config": f"OMGSR-F-1024 bf16 batch {B} (1 GPU)", "ms": {"vae_enc": enc, "dit": dit, "vae_dec": dec, "total": tot},
       "images_per_s": B / (tot * 1e-3), "tflops_algorithmic": {"vae_enc": 4.8826 * B / enc * 1e3, "dit": 74.3847 * B / dit * 1e3,
       "vae_dec": 10.4722 * B / dec * 1e3, "total": 89.74 * B / tot * 1e3}, "launches_per_step": res["launches"][-1],
       "finite": bool(torch.isfinite(img.float()).all()), "img_std": float(img.float().std())}
print(json.dumps(out))
