"""summarise an OMGSR_TRACE log (last pass): per stage, distinct (variant, shape) rows with count, total ms, TFLOP/s."""
import collections, re, sys
lines = open(sys.argv[1]).read().splitlines()
last = max(int(l.split()[2]) for l in lines if l.startswith("MARK"))
stage = None
agg = collections.OrderedDict()
for l in lines:
    if l.startswith("MARK"):
        _, _, p, st = l.split()
        stage = st if int(p) == last else None
    elif l.startswith("TRACE") and stage:
        m = re.match(r"TRACE (\S+) (.*) ([\d.]+) ms (\d+) TFLOP/s", l)
        key = (stage, m.group(1), m.group(2))
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += float(m.group(3)); a[2] += float(m.group(3)) * float(m.group(4))
tot = sum(a[1] for a in agg.values())
print(f"total traced GEMM/conv time {tot:.2f} ms")
for (st, var, shape), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:7.3f} ms {100 * a[1] / tot:5.1f}% n={a[0]:3d} {a[2] / a[1]:6.0f} TFLOP/s  {st:4s} {var:14s} {shape}")
