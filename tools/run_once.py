"""One device-resident batch through the engine (profiling target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
B = int(os.environ.get("B", "256")); reps = int(os.environ.get("REPS", "2"))
fam = os.environ.get("FAMILY", "birdnet_v24")
spec = get_spec(fam)
fe = spec.frontend
clf = bb.Classifier.builder().model_path(ensure_model(fam)).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
audio = synth.batch(0, B, fe.sample_count, fe.sample_rate)
d = torch.from_numpy(audio).cuda()
ctx = clf.create_batch_context(B, allow_perch=True) if fam == "perch_v2" else clf.create_batch_context(B)
for _ in range(reps):
    out = ctx.run_device(d.data_ptr(), B, True)
print("ok", out.batch, ctx.last_launch_count())
if os.environ.get("PROFILE"):
    import collections
    ctx.set_profiling(True)
    acc = collections.OrderedDict()
    for _ in range(5):
        ctx.run_device(d.data_ptr(), B, True)
        for n, ms in ctx.stage_times():
            acc.setdefault(n, []).append(ms)
    tot = 0
    for n, v in acc.items():
        m = sum(v) / len(v); tot += m
        print(f"{n:28s} {m:8.4f}")
    print("TOTAL", tot)
    import time
    ctx.set_profiling(False)
    t = time.time()
    for _ in range(20): ctx.enqueue_device(d.data_ptr(), B, True)
    ctx.wait(); dt = (time.time() - t) / 20
    print(f"back-to-back: {dt*1e3:.3f} ms/batch -> {B/dt:.0f} seg/s")
