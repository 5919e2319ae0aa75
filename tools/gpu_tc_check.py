"""Layer-by-layer comparison of the tcgen05 path against the FP32 CUDA-core path (dev aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels

fam = "birdnet_v24"
path = ensure_model(fam); spec = get_spec(fam)
labels = synthetic_labels(spec.num_species)
def mk():
    return bb.Classifier.builder().model_path(path).labels(labels).top_k(5).min_confidence(0.1).build()
os.environ["BN_DISABLE_TC"] = "1"; ref = mk()
os.environ["BN_DISABLE_TC"] = "0"; tc = mk()
B = int(os.environ.get("B", "6"))
audio = list(synth.batch(0, B, 144000, 48000))
cr, ct = ref.create_batch_context(B), tc.create_batch_context(B)
rr = ref.predict_batch_with_context(cr, audio)
rt = tc.predict_batch_with_context(ct, audio)
names = [op["out"] for op in spec.ops if op["op"] in ("conv", "gemm")]
bad = 0
for op in spec.ops:
    if op["op"] not in ("conv", "gemm"): continue
    # tensor names in the plan are the final (activated / residual-added) ONNX value names
    for cand in (op["out"],):
        try:
            a = cr.read_tensor(cand, B); b = ct.read_tensor(cand, B)
        except Exception as ex:
            continue
        d = np.abs(a - b).max(); s = np.abs(a).max()
        flag = "" if d <= 1e-4 * max(s, 1.0) else "   <<<<<<"
        bad += bool(flag)
        print(f"{op['name']:22s} {cand:28s} max|d|={d:.3e} max|ref|={s:.3e}{flag}")
lr = np.stack([r.raw_scores for r in rr]); lt = np.stack([r.raw_scores for r in rt])
print("logits max|d|", np.abs(lr - lt).max(), "bad layers", bad)
print("topk equal", all([p.index for p in a.predictions] == [p.index for p in b.predictions] for a, b in zip(rr, rt)))
import torch
B = 256
audio = synth.batch(0, B, 144000, 48000)
d = torch.from_numpy(audio).cuda()
for name, clf in (("fp32", ref), ("tc", tc)):
    ctx = clf.create_batch_context(B)
    for _ in range(2): ctx.run_device(d.data_ptr(), B, True)
    t = time.time(); n = 5
    for _ in range(n): ctx.enqueue_device(d.data_ptr(), B, True)
    ctx.wait(); dt = (time.time() - t) / n
    print(f"{name}: device-resident batch256 {dt*1e3:.2f} ms -> {B/dt:.0f} seg/s")
    ctx.set_profiling(True); ctx.run_device(d.data_ptr(), B, True)
    st = ctx.stage_times(); tot = sum(ms for _, ms in st)
    for nm, ms in sorted(st, key=lambda x: -x[1])[:14]:
        print(f"    {nm:28s} {ms:8.3f} ms {100*ms/tot:5.1f}%")
    print("    total", tot)
    del ctx
