"""Phase cycle counters of the fused MBConv kernel (CTA 0, thread 0): BN_MB_PROFILE=1 python tools/mb_phase_profile.py"""
import os, sys, ctypes as C
os.environ["BN_MB_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200 import _ffi
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
fam = sys.argv[1] if len(sys.argv) > 1 else "birdnet_v24"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
spec = get_spec(fam)
fe = spec.frontend
clf = bb.Classifier.builder().model_path(ensure_model(fam)).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
d = torch.from_numpy(synth.batch(0, B, fe.sample_count, fe.sample_rate)).cuda()
ctx = clf.create_batch_context(B, allow_perch=True) if fam == "perch_v2" else clf.create_batch_context(B)
for _ in range(2): ctx.run_device(d.data_ptr(), B, True)
ctx.set_profiling(True)
ctx.run_device(d.data_ptr(), B, True)
st = ctx.stage_times()
buf = (C.c_ulonglong * (128 * 16))()
fn = _ffi.lib.bn_debug_tc_profile
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
assert fn(buf, 128) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(128, 16).astype(np.float64)
ms = dict(st)
mb = [n for n, _ in st if n.endswith(".mbconv")]
print("kilo-cycles per SEGMENT (worker thread of CTA 0):  start | waitE | tmem->patch+A | depthwise+B | pool+fence | FC1 | FC2 | proj loop | wait mma | epilogue || total/seg  segs  stage ms")
k = 0
for slot in range(128):
    r = a[slot]
    if r[11] == 0 or r[10] == 0: continue
    n = r[10]
    name = mb[k] if k < len(mb) else "?"
    k += 1
    print(f"{name:14s} " + " ".join(f"{r[i]/n/1e3:9.2f}" for i in (0, 1, 2, 3, 8, 9, 4, 5, 6, 7)) + f" | ctl: tf {r[12]/n/1e3:.1f} w {r[13]/n/1e3:.1f} issue {r[14]/n/1e3:.1f} rest {r[15]/n/1e3:.1f}" + f" || {r[11]/n/1e3:9.2f} {int(n):4d}  {ms.get(name, 0):.4f}")
