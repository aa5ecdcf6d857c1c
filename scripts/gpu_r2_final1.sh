#!/bin/bash
# round-2 final evidence on ONE GPU: full test suite, both bench arms as the driver runs them, smoke, launch lists of an
# S-512 and an F-1024 step, kernel timings, stage timings, config 5 on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/final_gputests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/final_gputests.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "ref exit $?"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "ours exit $?"; tail -c 7000 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:omgsr --csv --log-file gpurun_out/launches_s512_b16_final.csv python scripts/s512_two_steps.py > gpurun_out/launches_final.log 2>&1
python profiles/summarise_launches.py gpurun_out/launches_s512_b16_final.csv 2 > gpurun_out/launches_s512_b16_final_summary.txt 2>&1; head -16 gpurun_out/launches_s512_b16_final_summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:omgsr --csv --log-file gpurun_out/launches_f1024_b4_final.csv python scripts/f1024.py 4 1 > gpurun_out/launches_f_final.log 2>&1
python profiles/summarise_launches.py gpurun_out/launches_f1024_b4_final.csv 2 > gpurun_out/launches_f1024_b4_final_summary.txt 2>&1; head -10 gpurun_out/launches_f1024_b4_final_summary.txt
timeout 200 python scripts/halo_ab.py > gpurun_out/final_halo.txt 2>&1
timeout 300 python scripts/attn_ab.py > gpurun_out/final_attn.txt 2>&1
timeout 300 python scripts/conv_time.py > gpurun_out/final_conv_time.txt 2>&1
timeout 300 python scripts/s512_stages.py > gpurun_out/final_s512_stages.txt 2>&1; tail -1 gpurun_out/final_s512_stages.txt
timeout 300 python scripts/f1024.py 4 4 > gpurun_out/final_f1024_stages.txt 2>&1; tail -1 gpurun_out/final_f1024_stages.txt | cut -c1-260
timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/final_c5_n1.json 2> gpurun_out/final_c5_n1.err; tail -c 600 gpurun_out/final_c5_n1.json
