"""Summarise `ncu --set full ... --page raw --csv` of ONE device-resident batch (tools/run_once.py,
REPS=1) into profiles/: a per-launch table and a {stage: dram bytes per launch} JSON that bench.py
reports as roofline.traffic.

    python tools/ncu_summarize.py gpurun_out/r01_ncu_full_raw.csv gpurun_out/stagesNN.log r01

The stage log (PROFILE=1 tools/run_once.py) gives the stage order; launches map to stages in order
(normalize = 2 launches with the fused front-end, 3 with the round-1 path; every other stage = 1)."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, stage_log, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
h, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(h)}


def f(row, name):
    try:
        return float(row[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


stages = []
for line in open(stage_log):
    p = line.split()
    if len(p) == 2 and p[0] not in ("TOTAL", "ok") and p[0] != "d2h":
        try:
            float(p[1])
        except ValueError:
            continue
        stages.append(p[0])
launch_stage = []
n_norm = 2 if "spectrogram" in stages else 3      # fused front-end: min/max init + pass; round-1 path: + k_normalize_emit
for s in stages:
    launch_stage += [s] * (n_norm if s == "normalize" else 1)

out_rows, traffic = [], {}
t_unit = units[col["gpu__time_duration.sum"]]
t_scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(t_unit, 1.0)      # -> microseconds
ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
for i, r in enumerate(data):
    st = launch_stage[i] if i < len(launch_stage) else "?"
    rd, wr = to_bytes(f(r, "dram__bytes_read.sum"), ur), to_bytes(f(r, "dram__bytes_write.sum"), uw)
    traffic[st] = traffic.get(st, 0.0) + rd + wr
    out_rows.append((i, st, r[col["Kernel Name"]].split("(")[0][:34], f(r, "gpu__time_duration.sum") * t_scale, rd / 1e6, wr / 1e6,
                     f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                     f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     int(f(r, "launch__grid_size")), int(f(r, "launch__registers_per_thread"))))
tot = sum(x[3] for x in out_rows)
dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt")
with open(dst, "w") as o:
    o.write(f"# ncu --set full --clock-control none, one device-resident batch of 256 (REPS=1 python tools/run_once.py), {len(out_rows)} launches\n")
    o.write("# durations are cold-cache, serialised replays: compare SHARES with the live CUDA-event stage times, not absolutes\n")
    o.write(f"# {'#':>3} {'stage':22s} {'kernel':34s} {'us':>8s} {'share':>6s} {'dramRdMB':>9s} {'dramWrMB':>9s} {'dram%':>6s} {'tensor%':>7s} {'L2%':>6s} {'warps%':>6s} {'grid':>6s} {'regs':>4s}\n")
    for x in out_rows:
        o.write(f"  {x[0]:3d} {x[1]:22s} {x[2]:34s} {x[3]:8.1f} {100*x[3]/tot:5.1f}% {x[4]:9.1f} {x[5]:9.1f} {x[6]:6.1f} {x[7]:7.1f} {x[8]:6.1f} {x[9]:6.1f} {x[10]:6d} {x[11]:4d}\n")
    o.write(f"# total {tot:.1f} us\n")
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w") as o:
    json.dump({"source": os.path.basename(raw), "batch": 256, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
               "traffic": {k: int(v) for k, v in traffic.items()}}, o, indent=1)
print("wrote", dst)
