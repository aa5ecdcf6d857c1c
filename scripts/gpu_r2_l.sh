#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2l_tests.log
timeout 120 python scripts/halo_ab.py > gpurun_out/r2l_halo.txt 2>&1; cat gpurun_out/r2l_halo.txt
timeout 300 python scripts/s512_stages.py > gpurun_out/r2l_s512_stages.txt 2>&1; tail -3 gpurun_out/r2l_s512_stages.txt
timeout 300 python scripts/f1024.py 4 4 > gpurun_out/r2l_f1024_stages.txt 2>&1; tail -1 gpurun_out/r2l_f1024_stages.txt
