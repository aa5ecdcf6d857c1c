#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -p no:cacheprovider -k "flash" > gpurun_out/k_flash.log 2>&1
echo "== flash exit $?" >> gpurun_out/summary.txt; grep -E "FAIL|passed|failed" gpurun_out/k_flash.log | tail -8 >> gpurun_out/summary.txt
timeout 300 python - > gpurun_out/attn_time.log 2>&1 <<'PY'
import torch, math, sys
sys.path.insert(0, ".")
from omgsr_b200 import capi
def t(B,H,M,N,D,n=5):
    q=torch.randn(B,M,H,D,device='cuda',dtype=torch.bfloat16); k=torch.randn(B,N,H,D,device='cuda',dtype=torch.bfloat16); v=torch.randn(B,N,H,D,device='cuda',dtype=torch.bfloat16); o=torch.empty_like(q)
    for _ in range(2): capi.flash_attn(q,k,v,1/math.sqrt(D),out=o)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): capi.flash_attn(q,k,v,1/math.sqrt(D),out=o)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n; return ms, 4*B*H*M*N*D/ms/1e9
for shp in [(4,24,4608,4608,128),(16,5,4096,4096,64),(16,10,1024,1024,64),(16,5,4096,77,64)]:
    ms,tf=t(*shp); print(shp,"v2-lean",f"{ms:.3f} ms {tf:.0f} TFLOP/s",flush=True)
PY
cat gpurun_out/attn_time.log >> gpurun_out/summary.txt
timeout 900 python scripts/f1024.py 4 3 > gpurun_out/f1024.log 2>&1
echo "== f1024 exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/f1024.log >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pipeline.log 2>&1
echo "== pipeline exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/pipeline.log >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt
cut -c1-330 gpurun_out/bench.json >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
