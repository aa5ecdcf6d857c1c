#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "combine_zero" > gpurun_out/k_comb.log 2>&1
echo "== combine test exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/k_comb.log >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/multigpu.log 2>&1
echo "== multigpu test exit $?" >> gpurun_out/summary.txt; grep -E "tile-parallel|passed|failed|Error" gpurun_out/multigpu.log | tail -8 >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/c5.py 224 > gpurun_out/c5_ngpu.log 2>&1
echo "== c5 ${N}gpu exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/c5_ngpu.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
