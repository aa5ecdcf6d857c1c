#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
OMGSR_TRACE=1 timeout 600 python scripts/s512_trace.py 16 2> gpurun_out/trace_s512.log
python scripts/summarise_trace.py gpurun_out/trace_s512.log > gpurun_out/trace_s512_summary.txt
head -70 gpurun_out/trace_s512_summary.txt >> gpurun_out/summary.txt
for d in 0 2 4 6; do OMGSR_HALO_DEBUG=$d timeout 120 python /dev/stdin >> gpurun_out/summary.txt 2>&1 <<'PY'
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
    ms=e0.elapsed_time(e1)/n; return ms, 2*Cin*Cout*9*B*H*W/ms/1e9
for shp in [(16,512,512,128,128),(16,256,256,256,128)]:
    ms,tf=tconv(*shp); print("dbg",os.environ.get("OMGSR_HALO_DEBUG"),"pair conv",shp,f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
PY
done
cat gpurun_out/summary.txt
