#!/bin/bash
# rolling-row kernel on by default: stage timings (S, F), then ncu --set full of the heaviest variant
mkdir -p gpurun_out
timeout 300 python scripts/s512_stages.py > gpurun_out/r2q_s512_stages.txt 2>&1; tail -1 gpurun_out/r2q_s512_stages.txt
OMGSR_HALO_ROLLING=0 timeout 300 python scripts/s512_stages.py > gpurun_out/r2q_s512_stages_rr0.txt 2>&1; tail -1 gpurun_out/r2q_s512_stages_rr0.txt
timeout 300 python scripts/f1024.py 4 4 > gpurun_out/r2q_f1024_stages.txt 2>&1; tail -1 gpurun_out/r2q_f1024_stages.txt
OMGSR_HALO_ROLLING=0 timeout 300 python scripts/f1024.py 4 4 > gpurun_out/r2q_f1024_stages_rr0.txt 2>&1; tail -1 gpurun_out/r2q_f1024_stages_rr0.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/ncu_halo_rr_gn_res python scripts/conv_probe.py halo_gn_res > gpurun_out/ncu_h.log 2>&1
echo "== ncu exit $?" > gpurun_out/ncu_halo_rr_summary.txt
python profiles/extract_ncu.py gpurun_out/ncu_halo_rr_gn_res.ncu-rep > gpurun_out/ncu_halo_rr_gn_res.txt 2>&1
ncu -i gpurun_out/ncu_halo_rr_gn_res.ncu-rep --page source --csv > gpurun_out/ncu_halo_rr_gn_res_source.csv 2>/dev/null
cat gpurun_out/ncu_halo_rr_gn_res.txt >> gpurun_out/ncu_halo_rr_summary.txt
python profiles/source_hotspots.py gpurun_out/ncu_halo_rr_gn_res_source.csv samples 60 >> gpurun_out/ncu_halo_rr_summary.txt
tail -90 gpurun_out/ncu_halo_rr_summary.txt
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2q_tests2.log 2>&1
echo "pytest2 rc=$?"; tail -4 gpurun_out/r2q_tests2.log | cut -c1-200
