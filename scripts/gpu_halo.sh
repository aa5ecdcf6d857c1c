#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -s -k "conv_halo" > gpurun_out/k_halo.log 2>&1
echo "== halo exit $?" > gpurun_out/summary.txt; grep -E "FAIL|ok rel|passed|failed" gpurun_out/k_halo.log | tail -60 >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "materialised or test_conv[" > gpurun_out/k4.log 2>&1
echo "== others exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/k4.log >> gpurun_out/summary.txt
timeout 600 python - > gpurun_out/halo_time.log 2>&1 <<'PY'
import torch, math
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
def t(B,H,W,Cin,Cout,flags,n=10):
    x=torch.randn(B,H,W,Cin,device='cuda',dtype=torch.bfloat16); w=pack_conv_weight(torch.randn(Cout,Cin,3,3)/math.sqrt(9*Cin)).cuda(); b=torch.zeros(Cout,device='cuda'); out=torch.empty(B,H,W,Cout,device='cuda',dtype=torch.bfloat16)
    for _ in range(3): capi.conv2d(x,w,b,out=out,flags=flags)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.conv2d(x,w,b,out=out,flags=flags)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 2*Cin*Cout*9*B*H*W/ms/1e9
for shp in [(16,512,512,128,128),(16,256,256,256,256),(16,128,128,512,512),(16,256,256,128,256)]:
    for name,fl in [("generic",capi.CONV_NO_HALO),("halo",0),("halo_desc1",capi.CONV_HALO_DESC1)]:
        ms,tf=t(*shp,fl); print(shp,name,f"{ms:.3f} ms {tf:.0f} TFLOP/s", flush=True)
PY
cat gpurun_out/halo_time.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
