#!/bin/bash
# round-2 evidence: ncu --set full of the dominant kernels (-> profiles/r02), config-5 on one GPU, launch list of a step
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/extract_ncu.py gpurun_out/ncu_$name.ncu-rep > gpurun_out/ncu_$name.txt 2>&1
  cat gpurun_out/ncu_$name.txt
}
cap r02_halo_rs_gn_res conv_halo 2 python scripts/conv_probe.py halo_gn_res
cap r02_halo_rs_gn conv_halo 2 python scripts/conv_probe.py halo_gn
cap r02_halo_rs_plain conv_halo 2 python scripts/conv_probe.py halo
cap r02_attn3_d128 flash_attn 2 python scripts/attn_probe.py 4 24 4608 4608 128
cap r02_attn3_d64 flash_attn 2 python scripts/attn_probe.py 16 5 4096 4096 64
cap r02_f_gemm_3072_12288 conv_gemm 2 python scripts/gemm_probe.py 18432 12288 3072 0 3
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_c5_n1.json 2> gpurun_out/r2_c5_n1.err; tail -c 400 gpurun_out/r2_c5_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_s512_b16_r02.csv python scripts/s512_two_steps.py > gpurun_out/launches_r02.log 2>&1
python profiles/summarise_launches.py gpurun_out/launches_s512_b16_r02.csv 2 > gpurun_out/launches_s512_b16_r02_summary.txt 2>&1; cat gpurun_out/launches_s512_b16_r02_summary.txt
