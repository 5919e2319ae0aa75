"""Host-path sweep: e2e segments/s vs pipeline depth and pack threads (dev aid)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
spec = get_spec("birdnet_v24"); path = ensure_model("birdnet_v24"); labels = synthetic_labels(spec.num_species)
B = 256
audio = synth.batch(0, B, 144000, 48000); segs = list(audio)
print("cpus", os.cpu_count())
# raw host memcpy bandwidth, single thread
dst = np.empty_like(audio); t = time.perf_counter()
for _ in range(3): np.copyto(dst, audio)
print("numpy copy GB/s", 3 * audio.nbytes / (time.perf_counter() - t) / 1e9)
for pt in (4, 8, 16):
    clf = bb.Classifier.builder().model_path(path).labels(labels).top_k(5).min_confidence(0.1).pack_threads(pt).build()
    for depth in (1, 2, 3, 4):
        ctxs = [clf.create_batch_context(B) for _ in range(depth)]
        for c in ctxs: clf.predict_batch_with_context(c, segs)
        n = 24
        def work(t):
            for i in range(t, n, depth): clf.predict_batch_with_context(ctxs[t], segs)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(depth)]
        [x.start() for x in th]; [x.join() for x in th]
        dt = time.perf_counter() - t0
        print(f"pack_threads={pt:2d} depth={depth}: {n*B/dt:8.0f} seg/s  ({dt/n*1e3:.2f} ms/batch)")
        del ctxs
    # split timing of one call
    ctx = clf.create_batch_context(B)
    t0 = time.perf_counter(); 
    from birdnet_b200.classifier import _segment_arrays
    for _ in range(5): _segment_arrays(segs)
    print("  _segment_arrays ms", (time.perf_counter() - t0) / 5 * 1e3)
    del clf
