#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for d in 0 8 2 10; do OMGSR_HALO_DEBUG=$d timeout 120 python /dev/stdin >> gpurun_out/summary.txt 2>&1 <<'PY'
import sys, math, torch, os
sys.path.insert(0, ".")
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
def tconv(B,H,W,Cin,Cout,n=10):
    x=torch.randn(B,H,W,Cin,device='cuda',dtype=torch.bfloat16); w=pack_conv_weight(torch.randn(Cout,Cin,3,3)/math.sqrt(9*Cin)).cuda(); b=torch.zeros(Cout,device='cuda'); out=torch.empty(B,H,W,Cout,device='cuda',dtype=torch.bfloat16)
    for _ in range(3): capi.conv2d(x,w,b,out=out)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.conv2d(x,w,b,out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 2*Cin*Cout*9*B*H*W/ms/1e9
ms,tf=tconv(16,512,512,128,128); print("dbg",os.environ.get("OMGSR_HALO_DEBUG"),f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
PY
done
cat gpurun_out/summary.txt
