"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv) -> profiles/<tag>_ncu_launches_summary.txt

    python tools/ncu_launch_summary.py profiles/r01b_ncu_launches.csv r01b "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ingest"
"""
import csv, collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]; ci = {n: i for i, n in enumerate(h)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])
    name = re.sub(r"^void (bn::)?", "", name)
    v = float(r[ci["Metric Value"]].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ci["Metric Unit"]], 1.0)
    tot[name] += v; cnt[name] += 1
allt = sum(tot.values())
dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_summary.txt")
with open(dst, "w") as o:
    o.write(f"# ncu launch list summary ({tag}): first {sum(cnt.values())} launches of `{cmd}`\n")
    o.write("# gpu__time_duration.sum, --clock-control none; cold-cache serialised replays: compare SHARES\n# unit: ns\n\n")
    for k, v in tot.most_common():
        o.write(f"{v:14.1f} {100*v/allt:5.1f}%  n={cnt[k]:4d}  {k}\n")
print("wrote", dst)
