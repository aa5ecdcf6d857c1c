#!/bin/bash
# round-2 final evidence on N GPUs (gpurun --gpus N): tile-parallel parity test, S-512 image-parallel line, config-5 line
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/final_multi${N}_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|Error" gpurun_out/final_multi${N}_tests.log | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload s512 --steps 10 --warmup 3 > gpurun_out/final_s512_n${N}.json 2> gpurun_out/final_s512_n${N}.err
echo "s512 rc=$? stdout lines: $(wc -l < gpurun_out/final_s512_n${N}.json)"; cut -c1-420 gpurun_out/final_s512_n${N}.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload f1024 --steps 5 --warmup 3 > gpurun_out/final_f1024_n${N}.json 2> gpurun_out/final_f1024_n${N}.err
echo "f1024 rc=$?"; cut -c1-300 gpurun_out/final_f1024_n${N}.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --workload c5 --steps 3 --warmup 3 > gpurun_out/final_c5_n${N}.json 2> gpurun_out/final_c5_n${N}.err
echo "c5 rc=$?"; cut -c1-300 gpurun_out/final_c5_n${N}.json; tail -3 gpurun_out/final_c5_n${N}.err
