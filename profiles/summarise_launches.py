"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals of the LAST step.
usage: summarise_launches.py launches.csv <steps in the capture (warm-up + timed)>"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = len(rows) // steps
last = rows[-n:]
def ms(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
agg = collections.defaultdict(lambda: [0, 0.0])
for r in last:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("omgsr::", "")
    agg[k][0] += 1; agg[k][1] += ms(r)
tot = sum(v[1] for v in agg.values())
print(f"{len(last)} launches in the last of {steps} steps, {tot:.2f} ms summed (cold-cache, serialised ncu timings)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:8.2f} ms {100 * v[1] / tot:5.1f}% n={v[0]:5d} {k[:90]}")
