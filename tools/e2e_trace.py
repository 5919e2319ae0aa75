"""BN_TRACE_RUN=1 timeline of `depth` driver threads on the public API (dev aid)."""
import os, sys, threading, time
os.environ["BN_TRACE_RUN"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
spec = get_spec("birdnet_v24"); path = ensure_model("birdnet_v24")
clf = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species))
       .top_k(5).min_confidence(0.1).pack_threads(8).build())
B = 256; depth = int(os.environ.get("DEPTH", "3")); n = int(os.environ.get("N", "30"))
audio = synth.batch(0, B, 144000, 48000)
pinned = bb.pinned_array(audio.shape); pinned[:] = audio
segs = list(pinned)
ctxs = [clf.create_batch_context(B) for _ in range(depth)]
for c in ctxs:
    clf.predict_batch_with_context(c, segs)
sys.stderr.write("[trace] ---- start\n"); sys.stderr.flush()
py = []
def work(t):
    for _ in range(t, n, depth):
        a = time.perf_counter()
        res = clf.predict_batch_with_context(ctxs[t], segs)
        py.append((t, a, time.perf_counter()))
t0 = time.perf_counter()
th = [threading.Thread(target=work, args=(t,)) for t in range(depth)]
[x.start() for x in th]; [x.join() for x in th]
dt = time.perf_counter() - t0
print(f"depth={depth}: {n*B/dt:.0f} seg/s ({dt/n*1e3:.2f} ms/batch)")
for t, a, b in sorted(py, key=lambda r: r[1]):
    print(f"py thread {t}: call {1e3*(a-t0):.2f} -> {1e3*(b-t0):.2f} ({1e3*(b-a):.2f} ms)")
