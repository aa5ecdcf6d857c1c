#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/l2_chain_probe.py 512 128 > gpurun_out/summary.txt 2>&1
timeout 300 python scripts/l2_chain_probe.py 256 256 >> gpurun_out/summary.txt 2>&1
cat gpurun_out/summary.txt
