"""Per-layer role wait cycles of the tensor-core conv kernel (CTA 0): who waits for whom.
BN_TC_PROFILE=1 python tools/tc_role_profile.py   (run on a GPU box)"""
import os, sys, ctypes as C
os.environ["BN_TC_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200 import _ffi
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
fam = sys.argv[1] if len(sys.argv) > 1 else "birdnet_v24"
spec = get_spec(fam)
S = spec.frontend.sample_count
clf = bb.Classifier.builder().model_path(ensure_model(fam)).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
B = 256
d = torch.from_numpy(synth.batch(0, B, S, spec.frontend.sample_rate)).cuda()
ctx = clf.create_batch_context(B)
for _ in range(2): ctx.run_device(d.data_ptr(), B, True)
ctx.set_profiling(True)
ctx.run_device(d.data_ptr(), B, True)
st = ctx.stage_times()
buf = (C.c_ulonglong * (128 * 16))()
fn = _ffi.lib.bn_debug_tc_profile
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
assert fn(buf, 128) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(128, 16).astype(np.float64)
# op index -> stage name: ops appear in stage order after the front-end stages
names = [n for n, _ in st if n not in ("normalize", "spectrogram0", "spectrogram1", "logmel", "topk_epilogue", "d2h", "h2d", "end")]
ms = dict(st)
print(f"{'slot':>4} {'tiles':>5} | producer wait/total | MMA wait_acc wait_full total | epi wait/total   (kilo-cycles, CTA 0)")
for slot in range(128):
    r = a[slot]
    if r[4] == 0: continue
    print(f"{slot:4d} {int(r[7]):5d} | {r[0]/1e3:8.1f} {r[1]/1e3:8.1f} | {r[2]/1e3:8.1f} {r[3]/1e3:8.1f} {r[4]/1e3:8.1f} | {r[5]/1e3:8.1f} {r[6]/1e3:8.1f}")
for n, m in st: print(f"{n:26s} {m:.4f}")
