#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -p no:cacheprovider -k "halo or groupnorm_stats" > gpurun_out/k_halo2.log 2>&1
echo "== halo tests exit $?" >> gpurun_out/summary.txt; grep -E "FAIL|passed|failed|Error|error" gpurun_out/k_halo2.log | tail -30 >> gpurun_out/summary.txt
timeout 300 python - > gpurun_out/time_halo2.log 2>&1 <<'PY'
import sys, math, torch
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
for on in (0,1):
    capi.lib().omgsr_set_gemm_2cta(on)
    for shp in [(16,512,512,128,128),(16,256,256,256,128),(16,512,512,128,3),(4,1024,1024,128,128)]:
        ms,tf=tconv(*shp); print("2cta" if on else "1cta","conv",shp,f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
PY
cat gpurun_out/time_halo2.log >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt; cat gpurun_out/bench.json >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
