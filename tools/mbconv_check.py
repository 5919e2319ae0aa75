#!/usr/bin/env python
"""GPU check of the fused MBConv kernel: same batch through the fused path and through the layer-by-layer path
(BN_DISABLE_MBCONV=1, separate process), block outputs and logits compared, stage times of both printed.

    python tools/mbconv_check.py [--family birdnet_v24] [--batch 8] [--times-batch 256]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def child(args):
    import birdnet_b200 as bb
    from birdnet_b200.modelgen import get_spec, synth, parse_model
    from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
    spec = get_spec(args.family)
    path = ensure_model(args.family)
    fe = spec.frontend
    clf = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build())
    B = args.batch
    audio = synth.batch(0, B, fe.sample_count, fe.sample_rate)
    ctx = clf.create_batch_context(B, allow_perch=True) if args.family == "perch_v2" else clf.create_batch_context(B)
    res = clf.predict_batch_with_context(ctx, list(audio))
    out = {"logits": np.stack([r.raw_scores for r in res])}
    # block outputs: the value the next block reads (projection output, or the residual add after it)
    with open(path, "rb") as f:
        m = parse_model(f.read())
    nodes = m["nodes"]
    names = []
    for i, n in enumerate(nodes):
        if n["op"] == "Conv" and ".project" in n.get("name", ""):
            o = n["outputs"][0]
            for n2 in nodes[i + 1:i + 3]:
                if n2["op"] == "Add" and o in n2["inputs"]:
                    o = n2["outputs"][0]
            names.append((n["name"], o))
    for nm, t in names:
        try:
            out["t:" + nm] = ctx.read_tensor(t, B)
        except Exception as e:       # noqa: BLE001
            out["e:" + nm] = np.array([0])
    np.savez(args.dump, **out)
    tb = args.times_batch
    if tb:
        import torch
        a2 = synth.batch(0, tb, fe.sample_count, fe.sample_rate)
        c2 = clf.create_batch_context(tb, allow_perch=True) if args.family == "perch_v2" else clf.create_batch_context(tb)
        d = torch.from_numpy(a2).cuda()
        for _ in range(3):
            c2.run_device(d.data_ptr(), tb, True)
        c2.set_profiling(True)
        acc = {}
        for _ in range(5):
            c2.run_device(d.data_ptr(), tb, True)
            for n, ms in c2.stage_times():
                acc.setdefault(n, []).append(ms)
        c2.set_profiling(False)
        times = [(n, float(np.mean(v))) for n, v in acc.items()]
        # back-to-back throughput
        import time
        for _ in range(3):
            c2.run_device(d.data_ptr(), tb, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            c2.enqueue_device(d.data_ptr(), tb, True)
            c2.wait()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        with open(args.dump + ".times.json", "w") as f:
            json.dump({"times": times, "total": sum(t for _, t in times), "launches": c2.last_launch_count(), "b2b_ms": dt * 1e3}, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="birdnet_v24")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--times-batch", type=int, default=256)
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    if args.dump:
        return child(args)
    outs = {}
    for tag, env in (("fused", {}), ("layers", {"BN_DISABLE_MBCONV": "1"})):
        dump = f"/tmp/mbconv_{tag}.npz"
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--family", args.family, "--batch", str(args.batch),
                            "--times-batch", str(args.times_batch), "--dump", dump], env=e, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print(f"[{tag}] FAILED rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}")
            return 1
        outs[tag] = (np.load(dump), json.load(open(dump + ".times.json")) if args.times_batch else None)
    a, b = outs["fused"][0], outs["layers"][0]
    worst = 0.0
    for k in a.files:
        if k.startswith("e:"):
            continue
        d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
        scale = max(1e-9, float(np.abs(b[k]).max()))
        bad = int((~np.isfinite(a[k])).sum())
        print(f"{k:28s} max|d| {d.max():.3e}  rel-to-max {d.max() / scale:.3e}  mean|d| {d.mean():.3e}  nonfinite {bad}  |ref|max {scale:.3g}")
        if k == "logits":
            worst = float(d.max())
    if args.times_batch:
        tf, tl = outs["fused"][1], outs["layers"][1]
        print(f"fused : total {tf['total']:.3f} ms, {tf['launches']} launches, back-to-back {tf['b2b_ms']:.3f} ms")
        print(f"layers: total {tl['total']:.3f} ms, {tl['launches']} launches, back-to-back {tl['b2b_ms']:.3f} ms")
        lt = dict(tl["times"])
        for n, ms in tf["times"]:
            if n.endswith(".mbconv"):
                blk = n[:-len(".mbconv")]
                old = sum(v for k, v in lt.items() if k.startswith(blk + "."))
                print(f"  {n:20s} {ms:.4f} ms   (layer by layer: {old:.4f} ms)")
    print("max |dlogit| fused vs layer-by-layer:", worst)
    return 0 if worst < 2e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
