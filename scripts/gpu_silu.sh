#!/bin/bash
mkdir -p gpurun_out
timeout 300 python /dev/stdin > gpurun_out/summary.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from omgsr_b200 import capi
# silu accuracy through the GroupNorm-apply kernel with identity coefficients: stats (mean 0, var 1 - eps) per group
C = 128
xs = torch.linspace(-14, 14, 4096 * C).reshape(1, 64, 64, C)[..., torch.randperm(C)].contiguous()
x = xs.bfloat16().cuda()
stats = torch.zeros(1, 32, 2, device="cuda"); stats[..., 1] = 1.0 - 1e-6
y = capi.groupnorm_apply(x, stats, torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), 32, 1e-6, True, shared_mod=1)
xf = x.float()
ref = torch.nn.functional.silu(xf)
err = (y.float() - ref)
refb = ref.bfloat16().float()
print(f"silu via tanh.approx: max |err| {err.abs().max():.3e}, mean signed err {err.mean():.3e}, "
      f"bf16-rounding-only max |err| {(refb - ref).abs().max():.3e}, mean {(refb - ref).mean():.3e}")
for lo, hi in ((-14, -8), (-8, -4), (-4, -1), (-1, 1), (1, 4), (4, 14)):
    m = (xf >= lo) & (xf < hi)
    print(f"  x in [{lo},{hi}): max |err| {err[m].abs().max():.3e} mean {err[m].mean():.3e}  (bf16-only max {(refb - ref)[m].abs().max():.3e})")
PY
timeout 200 python scripts/glue_time.py 2>&1 | grep gn_apply >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 >> gpurun_out/summary.txt
timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
cat gpurun_out/summary.txt
