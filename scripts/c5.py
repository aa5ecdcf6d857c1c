"""BASELINE config 5: OMGSR-F 512->2048, ONE image, tiles spread over the ranks (torchrun). Reference tile semantics
(VAEHook encoder tile 1024 / decoder tile 224 as in the reference's commented default, 3x3 latent tiles of 128)."""
import contextlib, io, json, os, sys, time
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from omgsr_b200 import dist as D
ws, rank, lr = D.env_world()
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
D.init(device=dev)
from infer.omgsr_f_infer_model import OMGSR_F_Infer
from infer.infer_omgsr_f import _prepare_latent_image_ids
dec_tile = int(sys.argv[1]) if len(sys.argv) > 1 else 224
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_F_Infer("random://flux_dev?seed=0", "random://lora?seed=1&rank=64", dev)
    m._init_tiled_vae(encoder_tile_size=1024, decoder_tile_size=dec_tile)
g = torch.Generator().manual_seed(1234)
lq = (torch.rand(1, 3, 2048, 2048, generator=g) * 2 - 1).to(dev).bfloat16()
pe = torch.randn(1, 512, 4096, generator=g).to(dev).bfloat16()
pp = torch.randn(1, 768, generator=g).to(dev).bfloat16()
tid = torch.zeros(512, 3, device=dev, dtype=torch.bfloat16)
ids = _prepare_latent_image_ids(64, 64, dev, torch.bfloat16)
m.vae.posterior_eps = torch.randn(1, 16, 256, 256, generator=g)
D.enable_tile_parallel(ws > 1)
times = []
with torch.no_grad():
    for it in range(3):
        D.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out, _ = m(lq, pe, pp, tid, ids, 128, 64)
        torch.cuda.synchronize(); D.barrier()
        times.append(time.perf_counter() - t0)
t = D.max_over_ranks(min(times[1:]), device=dev)
if rank == 0:
    print(json.dumps({"config": f"OMGSR-F 512->2048, 1 image, tile-parallel x{ws}, dec_tile {dec_tile}", "seconds": t,
                      "tflops_algorithmic": 744.1 / t, "finite": bool(torch.isfinite(out.float()).all()),
                      "std": float(out.float().std()), "mem_GiB": torch.cuda.max_memory_allocated() / 2**30}))
if ws > 1:
    torch.distributed.destroy_process_group()
