#!/bin/bash
mkdir -p gpurun_out
timeout 300 python /dev/stdin > gpurun_out/summary.txt 2>&1 <<'PY'
import sys, math, torch, os
sys.path.insert(0, ".")
from omgsr_b200 import capi
def tg(M,N,K,n=20,bias=True,force_bn=0,res=False):
    a=torch.randn(M,K,device='cuda',dtype=torch.bfloat16); w=torch.randn(N,K,device='cuda',dtype=torch.bfloat16)/math.sqrt(K)
    b=torch.zeros(N,device='cuda') if bias else None
    out=torch.empty(M,N,device='cuda',dtype=torch.bfloat16); r=torch.randn(M,N,device='cuda',dtype=torch.bfloat16) if res else None
    for _ in range(3): capi.gemm(a,w,b,res=r,out=out,force_bn=force_bn)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.gemm(a,w,b,res=r,out=out,force_bn=force_bn)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms
for (M,N,K) in [(65536,160,320),(65536,320,320),(65536,960,320),(65536,1280,320),(65536,1280,64),(65536,1280,1280),(65536,256,320),(65536,1024,320)]:
    for bias in (True,False):
        for fb in (0,64,128,256):
            if fb and N%fb: continue
            ms=tg(M,N,K,bias=bias,force_bn=fb)
            tiles=(M//128)*((N+ (fb or 160)-1)//(fb or 160))
            print(f"M={M} N={N} K={K} bias={int(bias)} force_bn={fb}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s  out-write {M*N*2/ms/1e6:.0f} GB/s",flush=True)
PY
cat gpurun_out/summary.txt
