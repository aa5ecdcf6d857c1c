#!/bin/bash
# round 2, call B: row-shared halo kernel + fp32 residual stream: parity first, then A/B timing and the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2b_tests.log
grep -n "full size\|batch 16\|C3 S\|FAILED\|passed\|failed" gpurun_out/r2b_tests.log | tail -12
OMGSR_HALO_ROW_SHARED=0 timeout 300 python scripts/halo_ab.py > gpurun_out/r2b_halo_old.txt 2>&1
timeout 300 python scripts/halo_ab.py > gpurun_out/r2b_halo_new.txt 2>&1
cat gpurun_out/r2b_halo_old.txt gpurun_out/r2b_halo_new.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print('S512', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'])
print('F1024', d['f1024']['value'], d['f1024']['ms_per_step'], d['f1024']['roofline']['frac'], d['f1024']['roofline']['attention'])
P
