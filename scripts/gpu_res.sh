#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "== kernel tests exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/k_all.log >> gpurun_out/summary.txt
OMGSR_TRACE=1 timeout 600 python scripts/s512_trace.py 16 2> gpurun_out/trace_s512.log
python scripts/summarise_trace.py gpurun_out/trace_s512.log > gpurun_out/trace_s512_summary.txt
head -${TRACE_LINES:-40} gpurun_out/trace_s512_summary.txt >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt; cat gpurun_out/bench.json >> gpurun_out/summary.txt
timeout 300 python scripts/s512_stages.py >> gpurun_out/summary.txt 2>&1
cat gpurun_out/summary.txt
