#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pipeline.log 2>&1
echo "== pipeline exit $?" > gpurun_out/summary.txt
grep -E "relerr|psnr|passed|failed|Error|error" gpurun_out/pipeline.log | tail -40 >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt; tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
# launch list of one step (cold-cache, serialised: SHARES only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 16 > gpurun_out/ncu_bench.log 2>&1
echo "== ncu exit $?" >> gpurun_out/summary.txt
wc -l gpurun_out/launches.csv >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
