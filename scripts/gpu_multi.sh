#!/bin/bash
# 2-GPU validation of the torchrun / image-parallel path + reference arm + F-1024 single-GPU timing
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi -L >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "== bench 2gpu exit $?" >> gpurun_out/summary.txt; cat gpurun_out/bench_2gpu.json >> gpurun_out/summary.txt; tail -3 gpurun_out/bench_2gpu.err >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "== bench reference exit $?" >> gpurun_out/summary.txt; cat gpurun_out/bench_ref.json >> gpurun_out/summary.txt
timeout 900 python scripts/f1024.py 4 3 > gpurun_out/f1024.log 2>&1
echo "== f1024 exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/f1024.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
