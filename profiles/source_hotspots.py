"""ncu --page source --csv -> instructions ranked by stall samples / executed count (SASS level)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    ins.append((n, r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), r))
tot_s = sum(i[2] for i in ins); tot_e = sum(i[3] for i in ins)
print(f"instructions {len(ins)} samples {tot_s} warp-instr executed {tot_e}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mode = sys.argv[2] if len(sys.argv) > 2 else "samples"
key = (lambda i: -i[2]) if mode == "samples" else (lambda i: -i[3])
for n, src, s, e, r in sorted(ins, key=key)[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{n:5d} samp {s:6d} ({100*s/tot_s:4.1f}%) exec {e:8d}  {src[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
