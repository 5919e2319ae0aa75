"""Basis-ring timeline of the fused v2.4 front-end kernel (CTA 0): per 12 KB slot the cycle the loader issued its bulk copy,
the cycle the control lane saw it landed, the cycle the control lane committed the slot's MMAs.

    python tools/fe_ring_trace.py [batch]
"""
import os, sys, ctypes as C
os.environ["BN_FE_PROFILE"] = "1"
os.environ["BN_FE_DEBUG"] = "1024"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200 import _ffi
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spec = get_spec("birdnet_v24")
clf = bb.Classifier.builder().model_path(ensure_model("birdnet_v24")).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
d = torch.from_numpy(synth.batch(0, B, 144000, 48000)).cuda()
ctx = clf.create_batch_context(B)
for _ in range(2): ctx.run_device(d.data_ptr(), B, True)
ctx.set_profiling(True)
ctx.run_device(d.data_ptr(), B, True)
buf = (C.c_ulonglong * (128 * 16))()
fn = _ffi.lib.bn_debug_tc_profile
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
assert fn(buf, 128) == 0
a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
n = 670
tr = a[:n * 3].reshape(n, 3)
ok = (tr[:, 0] > 0) | (np.arange(n) == 0)
issue, ready, commit = tr[:, 0], tr[:, 1], tr[:, 2]
NS = 6
def pct(x):
    x = np.asarray(x)
    return "min %d  p10 %d  median %d  p90 %d  max %d" % (x.min(), np.percentile(x, 10), np.median(x), np.percentile(x, 90), x.max())
sl = slice(20, n - NS)            # steady state
print("slots traced:", n, " cycles covered:", commit[n - 1] - issue[0])
print("control: slot period (commit[q+1] - commit[q])     ", pct(np.diff(commit)[sl]))
print("loader : issue period (issue[q+1] - issue[q])       ", pct(np.diff(issue)[sl]))
print("copy   : landed-seen - issued (upper bound latency) ", pct((ready - issue)[sl]))
print("control: commit - ready (issue time of the slot)    ", pct((commit - ready)[sl]))
print("ring   : issue[q+NS] - commit[q] (MMAs done -> loader reissues the slot)", pct((issue[NS:] - commit[:-NS])[20:]))
print("lead   : how many slots ahead of the control the loader is when it issues:", pct([np.searchsorted(ready, issue[q]) - q for q in range(20, n - NS)]))
np.save(os.path.join(ROOT, "gpurun_out", "fe_ring_trace.npy"), tr)
for q in range(100, 112):
    print(q, "issue", issue[q], "ready", ready[q], "commit", commit[q])
