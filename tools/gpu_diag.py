"""Per-stage parity + timing diagnostic (development aid; run on a GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
from oracle.model_oracle import ModelOracle, load_initializers
from oracle import postprocess_oracle as po

fam = "birdnet_v24"
path = ensure_model(fam)
spec = get_spec(fam)
t0 = time.time()
clf = bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
print("build s", time.time() - t0, "nproc", os.cpu_count())
B = 10
audio = synth.batch(0, B, 144000, 48000)
os.environ["BN_KEEP_NORMALIZED"] = "1"
ctx = clf.create_batch_context(B)
res = clf.predict_batch_with_context(ctx, list(audio))
orc = ModelOracle(spec, load_initializers(path))
ref = orc.forward(audio, keep=["spec"])
xt = torch.from_numpy(audio)
norm_ref = orc.frontend(xt)["normalized"].numpy()
norm = ctx.read_normalized(B)
print("normalized bit-exact:", np.array_equal(norm, norm_ref), "max diff", np.abs(norm - norm_ref).max())
sp = ctx.read_tensor("spec", B).reshape(B, 96, 511, 2).transpose(0, 3, 1, 2)
print("spec max|d|", np.abs(sp - ref["spec"]).max(), "ref max", np.abs(ref["spec"]).max())
got = np.stack([r.raw_scores for r in res])
print("logits max|d|", np.abs(got - ref["output"]).max(), "per seg", np.abs(got - ref["output"]).max(axis=1))
for i, r in enumerate(res):
    o = po.top_k_predictions(ref["output"][i], 5, 0.1)
    ok = [p.index for p in r.predictions] == [j for j, _ in o]
    dc = max([abs(p.confidence - c) for p, (_, c) in zip(r.predictions, o)] + [0])
    print(i, ok, dc, [(p.index, round(p.confidence, 4)) for p in r.predictions][:3])
print("launches", ctx.last_launch_count())
# timing at batch 256
B = 256
audio = synth.batch(0, B, 144000, 48000)
ctx = clf.create_batch_context(B)
segs = list(audio)
for _ in range(2):
    clf.predict_batch_with_context(ctx, segs)
t = time.time(); n = 3
for _ in range(n):
    clf.predict_batch_with_context(ctx, segs)
dt = (time.time() - t) / n
print(f"e2e host batch256: {dt*1e3:.2f} ms -> {B/dt:.0f} seg/s")
d = torch.from_numpy(audio).cuda()
torch.cuda.synchronize()
for _ in range(2):
    ctx.run_device(d.data_ptr(), B, True)
t = time.time()
for _ in range(n):
    ctx.run_device(d.data_ptr(), B, True)
dt = (time.time() - t) / n
print(f"device-resident batch256: {dt*1e3:.2f} ms -> {B/dt:.0f} seg/s")
ctx.set_profiling(True)
ctx.run_device(d.data_ptr(), B, True)
st = ctx.stage_times()
tot = sum(ms for _, ms in st)
for name, ms in sorted(st, key=lambda x: -x[1])[:25]:
    print(f"  {name:28s} {ms:8.3f} ms {100*ms/tot:5.1f}%")
print("total stage ms", tot)
