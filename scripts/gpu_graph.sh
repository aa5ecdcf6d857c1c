#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s -p no:cacheprovider -k "graph or nonsquare or pil_wrappers or driver_script" > gpurun_out/pipeline2.log 2>&1
echo "== new pipeline tests exit $?" >> gpurun_out/summary.txt
grep -E "relerr|psnr|passed|failed|Error|error|latency|level diff" gpurun_out/pipeline2.log | tail -30 >> gpurun_out/summary.txt
timeout 600 python scripts/s512_stages.py 16 > gpurun_out/stages16.log 2>&1; tail -1 gpurun_out/stages16.log >> gpurun_out/summary.txt
timeout 600 python scripts/s512_stages.py 1 > gpurun_out/stages1.log 2>&1; tail -1 gpurun_out/stages1.log >> gpurun_out/summary.txt
timeout 600 python - > gpurun_out/graph_b1.log 2>&1 <<'PY'
import contextlib, io, time, torch, sys
sys.path.insert(0, ".")
from infer.omgsr_s_infer_model import OMGSR_S_Infer
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, "cuda", torch.bfloat16)
x = (torch.rand(1, 3, 512, 512) * 2 - 1).cuda().bfloat16(); ctx = torch.randn(1, 77, 1024).cuda().bfloat16()
def t(n=10):
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for _ in range(3): m(x, ctx, 64, 32)
        t0 = time.perf_counter()
        for _ in range(n): m(x, ctx, 64, 32)
        return (time.perf_counter() - t0) / n * 1e3
e = t(); m.enable_cuda_graph(); g = t()
print(f"OMGSR-S-512 batch 1 latency through the module call: eager {e:.2f} ms, CUDA-graph replay {g:.2f} ms")
PY
tail -1 gpurun_out/graph_b1.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
