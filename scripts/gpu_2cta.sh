#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -p no:cacheprovider -k "two_cta" > gpurun_out/k_2cta.log 2>&1
echo "== 2cta tests exit $?" >> gpurun_out/summary.txt; grep -E "FAIL|ok rel|passed|failed|Error|error" gpurun_out/k_2cta.log | tail -30 >> gpurun_out/summary.txt
timeout 300 python - > gpurun_out/time_2cta.log 2>&1 <<'PY'
import sys, math, torch
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
def tg(M,N,K,n=10):
    a=torch.randn(M,K,device='cuda',dtype=torch.bfloat16); w=torch.randn(N,K,device='cuda',dtype=torch.bfloat16); out=torch.empty(M,N,device='cuda',dtype=torch.bfloat16)
    for _ in range(3): capi.gemm(a,w,out=out)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.gemm(a,w,out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 2*M*N*K/ms/1e9
for on in (0,1):
    capi.lib().omgsr_set_gemm_2cta(on)
    for shp in [(16,256,256,256,256),(16,128,128,512,512),(16,256,256,128,256),(16,64,64,512,512)]:
        ms,tf=tconv(*shp); print("2cta" if on else "1cta","conv",shp,f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
    for shp in [(16384,3072,3072),(18432,12288,3072),(18432,3072,15360),(8192,8192,8192),(65536,2560,320)]:
        ms,tf=tg(*shp); print("2cta" if on else "1cta","gemm",shp,f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
PY
cat gpurun_out/time_2cta.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
