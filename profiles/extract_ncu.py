"""ncu -i <rep> --page raw --csv  ->  compact key-metric summary (committed under profiles/rNN/)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "?")[:120])
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {u.get(k, '')}")
        stalls = {k: float(v) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_")
                  and not k.endswith("_not_issued") and v not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            print(f"  stall {k[len('smsp__pcsamp_warps_issue_stalled_'):]:28s} {100 * v / tot:5.1f} %")


if __name__ == "__main__":
    for r in sys.argv[1:]:
        print("==", r)
        main(r)
