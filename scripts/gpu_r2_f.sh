#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/mma_probe > gpurun_out/mma_probe4.txt 2>&1; tail -32 gpurun_out/mma_probe4.txt | grep "single"
timeout 120 python scripts/halo_ab.py > gpurun_out/r2f_halo.txt 2>&1; cat gpurun_out/r2f_halo.txt
