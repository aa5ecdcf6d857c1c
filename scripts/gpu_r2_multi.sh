#!/bin/bash
# N-GPU run (gpurun --gpus N): tile-parallel parity test + the config-5 bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2_multi${N}_tests.log 2>&1
echo "pytest rc=$?"; grep -n "tile-parallel\|passed\|failed\|Error" gpurun_out/r2_multi${N}_tests.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_c5_n${N}.json 2> gpurun_out/r2_c5_n${N}.err
echo "c5 rc=$?"; tail -c 1500 gpurun_out/r2_c5_n${N}.json; tail -5 gpurun_out/r2_c5_n${N}.err
