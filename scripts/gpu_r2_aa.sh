#!/bin/bash
# ncu --set full: DiT GEMM 18432 x 12288 x 3072 (+GELU), 3x3 256->256 @256^2 x16 conv, UNet 3x3 320->320 @64^2 x16 conv
mkdir -p gpurun_out
cap() {  # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/extract_ncu.py gpurun_out/ncu_$name.ncu-rep > gpurun_out/ncu_$name.txt 2>&1
  ncu -i gpurun_out/ncu_$name.ncu-rep --page source --csv > gpurun_out/ncu_${name}_source.csv 2>/dev/null
  python profiles/source_hotspots.py gpurun_out/ncu_${name}_source.csv samples 14 >> gpurun_out/ncu_$name.txt
  cut -c1-160 gpurun_out/ncu_$name.txt
}
cap r02b_f_gemm conv_gemm python scripts/gemm_probe.py 18432 12288 3072 0 3
cap r02b_conv256 conv_gemm python scripts/conv_probe.py generic 16 256 256 256 256
cap r02b_conv320 conv_gemm python scripts/conv_probe.py generic 16 64 64 320 320
