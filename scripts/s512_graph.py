"""OMGSR-S-512 batch 16: eager launch sequence vs CUDA-graph replay of the same sequence (module forward)."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infer.omgsr_s_infer_model import OMGSR_S_Infer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, "cuda", torch.bfloat16)
x = (torch.rand(B, 3, 512, 512) * 2 - 1).cuda().bfloat16()
ctx = torch.randn(1, 77, 1024).cuda().bfloat16()
def run(n):
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for _ in range(3): m(x, ctx, 64, 32)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): m(x, ctx, 64, 32)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_eager = run(8)
m.enable_cuda_graph()
t_graph = run(8)
print(f"batch {B}: eager {t_eager:.2f} ms/step ({B / t_eager * 1e3:.1f} images/s), graph replay {t_graph:.2f} ms/step "
      f"({B / t_graph * 1e3:.1f} images/s), mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
