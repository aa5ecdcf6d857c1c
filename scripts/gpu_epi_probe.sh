#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
for d in 0 4 5; do OMGSR_HALO_DEBUG=$d timeout 120 python /dev/stdin >> gpurun_out/summary.txt 2>&1 <<'PY'
import sys, math, torch, os
sys.path.insert(0, ".")
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
def tconv(B,H,W,Cin,Cout,n=10,res=False):
    x=torch.randn(B,H,W,Cin,device='cuda',dtype=torch.bfloat16); w=pack_conv_weight(torch.randn(Cout,Cin,3,3)/math.sqrt(9*Cin)).cuda(); b=torch.zeros(Cout,device='cuda'); out=torch.empty(B,H,W,max(Cout,8),device='cuda',dtype=torch.bfloat16)
    r=torch.randn(B,H,W,Cout,device='cuda',dtype=torch.bfloat16) if res else None
    for _ in range(3): capi.conv2d(x,w,b,res=r,out=out,cout=Cout)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.conv2d(x,w,b,res=r,out=out,cout=Cout)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 2*Cin*Cout*9*B*H*W/ms/1e9
for res in (False, True):
    ms,tf=tconv(16,512,512,128,128,res=res); print("dbg",os.environ.get("OMGSR_HALO_DEBUG"),"res",res,f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
PY
done
timeout 200 python /dev/stdin >> gpurun_out/summary.txt 2>&1 <<'PY'
import sys, math, torch, os
sys.path.insert(0, ".")
from omgsr_b200 import capi
def tg(M,N,K,n=20,res=False,act=0):
    a=torch.randn(M,K,device='cuda',dtype=torch.bfloat16); w=torch.randn(N,K,device='cuda',dtype=torch.bfloat16)/math.sqrt(K); bias=torch.zeros(N,device='cuda')
    No = N//2 if act==capi.ACT_GEGLU else N
    out=torch.empty(M,No,device='cuda',dtype=torch.bfloat16); r=torch.randn(M,No,device='cuda',dtype=torch.bfloat16) if res else None
    for _ in range(3): capi.gemm(a,w,bias,res=r,out=out,act=act)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.gemm(a,w,bias,res=r,out=out,act=act)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 2*M*N*K/ms/1e9
for on in (1,0):
    capi.lib().omgsr_set_gemm_2cta(on)
    for (M,N,K,res,act) in [(65536,320,320,False,0),(65536,320,320,True,0),(65536,960,320,False,0),(65536,2560,320,False,capi.ACT_GEGLU),(16384,640,640,True,0),(16384,5120,640,False,capi.ACT_GEGLU),(65536,320,1280,True,0),(4096,1280,1280,True,0),(4096,10240,1280,False,capi.ACT_GEGLU),(4096,1280,5120,True,0)]:
        ms,tf=tg(M,N,K,res=res,act=act)
        mb=(M*K+M*(N//2 if act else N)*(2 if res else 1))*2/1e6
        print("pair" if on else "single",(M,N,K),"res",res,"act",act,f"{ms*1e3:.1f} us {tf:.0f} TFLOP/s  hbm-floor {mb/6.55e3*1e3:.1f} us  mma-floor {2*M*N*K/1.5e15*1e6:.1f} us",flush=True)
PY
cat gpurun_out/summary.txt
