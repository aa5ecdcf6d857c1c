#!/bin/bash
mkdir -p gpurun_out
timeout 300 python /dev/stdin > gpurun_out/summary.txt 2>&1 <<'PY'
import sys, math, torch, os
sys.path.insert(0, ".")
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
def tconv(B,H,W,Cin,Cout,n=10):
    x=torch.randn(B,H,W,Cin,device='cuda',dtype=torch.bfloat16); w=pack_conv_weight(torch.randn(Cout,Cin,3,3)/math.sqrt(9*Cin)).cuda(); b=torch.zeros(Cout,device='cuda'); out=torch.empty(B,H,W,max(Cout,8),device='cuda',dtype=torch.bfloat16)
    for _ in range(3): capi.conv2d(x,w,b,out=out,cout=Cout)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.conv2d(x,w,b,out=out,cout=Cout)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
print(f"conv 128->3 @512^2 x16 (BN=16 halo): {tconv(16,512,512,128,3):.3f} ms")
print(f"conv 128->128 @512^2 x16: {tconv(16,512,512,128,128):.3f} ms")
PY
OMGSR_TRACE=1 timeout 600 python scripts/s512_trace.py 16 2> gpurun_out/trace_s512.log
python scripts/summarise_trace.py gpurun_out/trace_s512.log > gpurun_out/trace_s512_summary.txt
grep -E "Cout=3 |Cout=8 |Cout=4 |Cin=32 |Cin=64 " gpurun_out/trace_s512_summary.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
