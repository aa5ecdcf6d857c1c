#!/bin/bash
# programmatic dependent launch: full GPU test suite, then same-box A/B of the step timings (OMGSR_PDL=0|1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2x_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2x_tests.log | cut -c1-200
for rep in 1 2; do for pdl in 0 1; do
  OMGSR_PDL=$pdl timeout 300 python scripts/s512_stages.py 2>&1 | tail -1 | sed "s/^/PDL=$pdl S: /"
  OMGSR_PDL=$pdl timeout 300 python scripts/f1024.py 4 4 2>&1 | tail -1 | cut -c1-200 | sed "s/^/PDL=$pdl F: /"
done; done > gpurun_out/r2x_pdl_ab.txt 2>&1; cat gpurun_out/r2x_pdl_ab.txt
timeout 200 python scripts/halo_ab.py > gpurun_out/r2x_halo.txt 2>&1; head -6 gpurun_out/r2x_halo.txt
