#!/bin/bash
# what the driver runs at round end, in the same order: gpu tests (-x), smoke, bench (+ reference arm)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt
timeout 900 python scripts/f1024.py 4 3 > gpurun_out/f1024.log 2>&1; tail -1 gpurun_out/f1024.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
