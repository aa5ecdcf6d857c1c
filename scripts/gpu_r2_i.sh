#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv or halo or gn or group" > gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2i_tests.log
timeout 120 python scripts/halo_ab.py > gpurun_out/r2i_halo.txt 2>&1; cat gpurun_out/r2i_halo.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('S512', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'])
print('F1024', d['f1024']['value'], d['f1024']['ms_per_step'], d['f1024']['roofline']['frac'], d['f1024']['roofline']['attention'])
P
