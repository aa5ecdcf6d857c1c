#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2k_tests.log
for dbg in 0 8; do echo "== OMGSR_HALO_DEBUG=$dbg (8 = no L2 prefetch of the next tile's rows)"; OMGSR_HALO_DEBUG=$dbg timeout 120 python scripts/halo_ab.py; done > gpurun_out/r2k_halo_prefetch_ab.txt 2>&1; cat gpurun_out/r2k_halo_prefetch_ab.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --workload s512 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('S512', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'])
P
