#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/multigpu.log 2>&1
echo "== multigpu test exit $?" >> gpurun_out/summary.txt; grep -E "tile-parallel|passed|failed|Error" gpurun_out/multigpu.log | tail -8 >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "thin" > gpurun_out/k5.log 2>&1
echo "== thin exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/k5.log >> gpurun_out/summary.txt
timeout 900 python scripts/c5.py 224 > gpurun_out/c5_1gpu.log 2>&1
echo "== c5 1gpu exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/c5_1gpu.log >> gpurun_out/summary.txt
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/c5.py 224 > gpurun_out/c5_ngpu.log 2>&1
echo "== c5 ${N}gpu exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/c5_ngpu.log >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
