#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 8; do echo "== OMGSR_HALO_DEBUG=$dbg"; OMGSR_HALO_DEBUG=$dbg timeout 200 python scripts/halo_ab.py 2>&1 | head -6; done > gpurun_out/r2s_halo_relaxed.txt 2>&1; cat gpurun_out/r2s_halo_relaxed.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_rr_gn python scripts/conv_probe.py halo_gn > gpurun_out/ncu_h.log 2>&1
echo "== ncu exit $?" > gpurun_out/ncu_halo_rr_gn_summary.txt
python profiles/extract_ncu.py gpurun_out/ncu_halo_rr_gn.ncu-rep >> gpurun_out/ncu_halo_rr_gn_summary.txt 2>&1
ncu -i gpurun_out/ncu_halo_rr_gn.ncu-rep --page source --csv > gpurun_out/ncu_halo_rr_gn_source.csv 2>/dev/null
python profiles/source_hotspots.py gpurun_out/ncu_halo_rr_gn_source.csv samples 40 >> gpurun_out/ncu_halo_rr_gn_summary.txt
cat gpurun_out/ncu_halo_rr_gn_summary.txt | cut -c1-170
