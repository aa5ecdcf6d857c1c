#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s -p no:cacheprovider -k "fused_input_groupnorm" > gpurun_out/k_gnin.log 2>&1
echo "== fused-input-gn tests exit $?" >> gpurun_out/summary.txt; grep -E "FAIL|first bad|dim[0-9]|passed|failed|Error|error" gpurun_out/k_gnin.log | tail -30 >> gpurun_out/summary.txt
timeout 300 python /dev/stdin >> gpurun_out/summary.txt 2>&1 <<'PY'
import sys, math, torch
sys.path.insert(0, ".")
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
B,H,W,C=16,512,512,128
x=torch.randn(B,H,W,C,device='cuda',dtype=torch.bfloat16); w=pack_conv_weight(torch.randn(C,C,3,3)/math.sqrt(9*C)).cuda(); b=torch.zeros(C,device='cuda')
g=torch.ones(C,device='cuda'); bt=torch.zeros(C,device='cuda'); out=torch.empty_like(x); z=torch.empty_like(x)
st=capi.groupnorm_stats(x,32)
def timeit(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
def two():
    capi.groupnorm_apply(x,st,g,bt,32,1e-6,True,out=z); capi.conv2d(z,w,b,out=out)
def fused():
    coef=capi.groupnorm_coef(st,g,bt,B,H*W,32,1e-6); capi.conv2d_gn(x,coef,w,b,out=out)
print(f"gn_apply + conv 128->128 @512^2 x16: {timeit(two):.3f} ms; fused: {timeit(fused):.3f} ms; conv alone: {timeit(lambda: capi.conv2d(z,w,b,out=out)):.3f} ms")
PY
cat gpurun_out/summary.txt
