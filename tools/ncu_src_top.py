"""Top stall sites of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
h = r[1]
rows = r[2:]
ci = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(x[ci["# Samples"]] or 0) for x in rows)
print("total samples", tot)
agg = {}
for x in rows:
    for n in stall_cols:
        agg[n] = agg.get(n, 0) + int(x[ci[n]] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(rows, key=lambda x: -int(x[ci["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for x in top:
    s = int(x[ci["# Samples"]] or 0)
    st = {n[6:]: int(x[ci[n]] or 0) for n in stall_cols if int(x[ci[n]] or 0)}
    print(f"{s:6d} {100*s/tot:5.1f}% {x[ci['Source']].strip()[:70]:70s} exec={x[ci['Instructions Executed']]} {st}")
