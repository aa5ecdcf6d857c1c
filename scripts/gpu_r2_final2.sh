#!/bin/bash
# last check of the bench line after the roofline kernels were re-ranked + ncu capture of the `second` probe shape
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/ncu_r02c_conv512 python scripts/conv_probe.py generic 16 128 128 512 512 > gpurun_out/ncu_c512.log 2>&1
python profiles/extract_ncu.py gpurun_out/ncu_r02c_conv512.ncu-rep > gpurun_out/ncu_r02c_conv512.txt 2>&1; head -16 gpurun_out/ncu_r02c_conv512.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
echo "ours exit $?"; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/final2_bench.json') if l.startswith('{')][-1])
print("S", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roofline", d["roofline"]["kernel"][:40], round(d["roofline"]["frac"],3), "second", d["roofline"]["second"]["kernel"][:40], round(d["roofline"]["second"]["frac"],3), "clocks", d["clocks"]["sm_mhz"])
f=d["f1024"]; print("F", round(f["value"],2), "e2e", round(f["e2e"]["value"],2), "gemm frac", round(f["roofline"]["frac"],3), "attn", round(f["roofline"]["attention"]["frac"],3), "clocks", f["clocks"]["sm_mhz"])
P
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
