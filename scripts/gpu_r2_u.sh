#!/bin/bash
# after the warp-uniform issue change: attention tests, timings, launch lists of one S-512 and one F-1024 step
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:omgsr --csv --log-file gpurun_out/launches_s512_b16_r02b.csv python scripts/s512_two_steps.py > gpurun_out/launches_r02b.log 2>&1
python profiles/summarise_launches.py gpurun_out/launches_s512_b16_r02b.csv 2 > gpurun_out/launches_s512_b16_r02b_summary.txt 2>&1; cat gpurun_out/launches_s512_b16_r02b_summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:omgsr --csv --log-file gpurun_out/launches_f1024_b4_r02.csv python scripts/f1024.py 4 1 > gpurun_out/launches_f_r02.log 2>&1
python profiles/summarise_launches.py gpurun_out/launches_f1024_b4_r02.csv 2 > gpurun_out/launches_f1024_b4_r02_summary.txt 2>&1; cat gpurun_out/launches_f1024_b4_r02_summary.txt
