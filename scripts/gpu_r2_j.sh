#!/bin/bash
mkdir -p gpurun_out
for v in 2 0 2 0; do echo "== OMGSR_ATTN_VARIANT=$v"; OMGSR_ATTN_VARIANT=$v timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1; done > gpurun_out/r2j_f1024_attn_ab.txt 2>&1
cat gpurun_out/r2j_f1024_attn_ab.txt
