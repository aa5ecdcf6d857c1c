#!/bin/bash
# N-GPU evidence: multi-GPU parity tests, S-512 image-parallel bench, F-1024 bench, F 512->2048 tile-parallel
mkdir -p gpurun_out
: > gpurun_out/summary.txt
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/multigpu.log 2>&1
echo "== multigpu test exit $?" >> gpurun_out/summary.txt; grep -E "tile-parallel|passed|failed|Error" gpurun_out/multigpu.log | tail -8 >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_s512_n$N.json 2> gpurun_out/bench_n.err
echo "== bench s512 x$N exit $?" >> gpurun_out/summary.txt; cat gpurun_out/bench_s512_n$N.json >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload f1024 --steps 3 --warmup 3 > gpurun_out/bench_f1024_n$N.json 2> gpurun_out/bench_f.err
echo "== bench f1024 x$N exit $?" >> gpurun_out/summary.txt; cat gpurun_out/bench_f1024_n$N.json >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/c5.py 224 > gpurun_out/c5_ngpu.log 2>&1
echo "== c5 ${N}gpu exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/c5_ngpu.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
