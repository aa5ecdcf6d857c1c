#!/bin/bash
# round-2 check of the default bench line (S-512 + f1024 sub-object), both arms, as the driver runs them
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err
echo "ref exit $?"; tail -c 1500 gpurun_out/r2n_bench_reference.json
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "ours exit $?"; tail -c 6000 gpurun_out/r2n_bench.json; tail -5 gpurun_out/r2n_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
