#!/bin/bash
# last full check with flash_attn5 as the default: whole GPU suite, smoke, bench both arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/final3_gputests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/final3_gputests.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final3_bench_reference.json 2> gpurun_out/final3_bench_reference.err; echo "ref exit $?"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err
echo "ours exit $?"; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/final3_bench.json') if l.startswith('{')][-1])
print("S", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],3), "second", round(d["roofline"]["second"]["frac"],3), "clocks", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], "eager", round(d["gpu_eager_baseline"]["value"],1))
f=d["f1024"]; print("F", round(f["value"],2), "ms", round(f["ms_per_step"],1), "e2e", round(f["e2e"]["value"],2), "gemm frac", round(f["roofline"]["frac"],3), "attn", round(f["roofline"]["attention"]["achieved"],0), round(f["roofline"]["attention"]["frac"],3), "clocks", f["clocks"]["sm_mhz"], "eager", round(f["gpu_eager_baseline"]["value"],2))
P
timeout 300 python scripts/s512_stages.py 2>&1 | tail -1
