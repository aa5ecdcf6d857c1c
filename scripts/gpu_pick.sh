#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "== kernel tests exit $?" > gpurun_out/summary.txt; tail -3 gpurun_out/k_all.log >> gpurun_out/summary.txt
timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
OMGSR_TRACE=1 timeout 600 python scripts/s512_trace.py 16 2> gpurun_out/trace_s512.log
python scripts/summarise_trace.py gpurun_out/trace_s512.log > gpurun_out/trace_s512_summary.txt
grep -E "H=8 W=8|H=16 W=16|W=1024 |W=4096 " gpurun_out/trace_s512_summary.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
