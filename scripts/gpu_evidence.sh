#!/bin/bash
# evidence run for profiles/: ncu --set full on the top kernels (one launch each), launch list of one bench step
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_pair128 python scripts/conv_probe.py halo > gpurun_out/ncu1.log 2>&1
echo "== ncu halo exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_pair128_gn python scripts/conv_probe.py halo_gn > gpurun_out/ncu1b.log 2>&1
echo "== ncu halo_gn exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/ncu_gemm_pair256 python scripts/conv_probe.py generic 16 256 256 256 256 > gpurun_out/ncu2.log 2>&1
echo "== ncu gemm256 exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_apply|flash_attn2" -c 40 -f -o gpurun_out/ncu_glue python scripts/glue_probe.py > gpurun_out/ncu3.log 2>&1
echo "== ncu glue exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn2 -s 2 -c 1 -f -o gpurun_out/ncu_flash64 python scripts/attn_probe.py 16 5 4096 4096 64 > gpurun_out/ncu4.log 2>&1
echo "== ncu flash64 exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== launch list exit $?" >> gpurun_out/summary.txt
python profiles/summarise_launches.py gpurun_out/launches.csv 4 > gpurun_out/launches_summary.txt 2>&1
for r in ncu_halo_pair128 ncu_halo_pair128_gn ncu_gemm_pair256 ncu_glue ncu_flash64; do python profiles/extract_ncu.py gpurun_out/$r.ncu-rep > gpurun_out/$r.txt 2>&1; done
cat gpurun_out/summary.txt; cat gpurun_out/ncu_halo_pair128_gn.txt; head -20 gpurun_out/launches_summary.txt
