#!/bin/bash
# ncu --set full of the heaviest 128-channel conv variant (fused input GroupNorm + residual + output statistics)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_rs_gn_res python scripts/conv_probe.py halo_gn_res > gpurun_out/ncu_h.log 2>&1
echo "== ncu exit $?" > gpurun_out/ncu_halo_summary.txt
python profiles/extract_ncu.py gpurun_out/ncu_halo_rs_gn_res.ncu-rep > gpurun_out/ncu_halo_rs_gn_res.txt 2>&1
ncu -i gpurun_out/ncu_halo_rs_gn_res.ncu-rep --page source --csv > gpurun_out/ncu_halo_rs_gn_res_source.csv 2>/dev/null
cat gpurun_out/ncu_halo_rs_gn_res.txt >> gpurun_out/ncu_halo_summary.txt
python profiles/source_hotspots.py gpurun_out/ncu_halo_rs_gn_res_source.csv samples 45 >> gpurun_out/ncu_halo_summary.txt
ncu -i gpurun_out/ncu_halo_rs_gn_res.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for k,x in zip(h,v):
    if any(t in k for t in ('sm__inst_executed_pipe_','sm__pipe_','smsp__inst_executed_pipe','l1tex__data_pipe','sm__mio','shared','smsp__warps_issue_stalled')) and 'pct' in k: print(k,x)
" >> gpurun_out/ncu_halo_summary.txt
tail -120 gpurun_out/ncu_halo_summary.txt
