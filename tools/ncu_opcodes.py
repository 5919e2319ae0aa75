"""Opcode histogram (executed warp instructions + stall samples) of an ncu source-page CSV, and the
source lines with the most stall samples.   ncu -i X.ncu-rep --page source --csv > src.csv"""
import csv, re, collections, sys
r = list(csv.reader(open(sys.argv[1])))
h = r[1]; rows = r[2:]
ci = {n: i for i, n in enumerate(h)}
tot = sum(int(x[ci['Instructions Executed']] or 0) for x in rows)
agg = collections.Counter(); samp = collections.Counter()
for x in rows:
    src = x[ci['Source']].strip()
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src[:10]
    agg[op] += int(x[ci['Instructions Executed']] or 0)
    samp[op] += int(x[ci['# Samples']] or 0)
ts = sum(samp.values()) or 1
print('total warp instr', tot, 'samples', ts)
for op, c in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print(f"{op:12s} {c:10d} {100*c/tot:5.1f}%  samples {100*samp[op]/ts:5.1f}%")
