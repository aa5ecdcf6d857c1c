import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
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
for shp in [(16,512,512,128,128),(16,256,256,256,256),(16,128,128,512,512),(16,256,256,128,256),(16,64,64,512,512),(16,64,64,320,320),(16,32,32,640,640),(16,64,64,960,320)]:
    for name,fl in [("generic",capi.CONV_NO_HALO),("auto",0)]:
        ms,tf=t(*shp,fl); print(shp,name,f"{ms:.3f} ms {tf:.0f} TFLOP/s", flush=True)
def tg(M,N,K,n=10):
    a=torch.randn(M,K,device='cuda',dtype=torch.bfloat16); w=torch.randn(N,K,device='cuda',dtype=torch.bfloat16); out=torch.empty(M,N,device='cuda',dtype=torch.bfloat16)
    for _ in range(3): capi.gemm(a,w,out=out)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.gemm(a,w,out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 2*M*N*K/ms/1e9
for shp in [(65536,320,320),(65536,2560,320),(65536,320,1280),(16384,3072,3072),(18432,12288,3072),(18432,3072,15360),(8192,8192,8192)]:
    ms,tf=tg(*shp); print("gemm",shp,f"{ms:.3f} ms {tf:.0f} TFLOP/s", flush=True)
