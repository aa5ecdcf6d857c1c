#!/bin/bash
# round 2, call A: the whole GPU suite (incl. the new full-size parity / determinism tests), the default bench line
# (S-512 + f1024 sub-object + eager / CPU baselines) and the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --durations=12 -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
