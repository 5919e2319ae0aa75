"""Where the end-to-end (host slices in, host results out) time goes.  Dev aid; run on a GPU box."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
from birdnet_b200.classifier import _segment_arrays

spec = get_spec("birdnet_v24"); path = ensure_model("birdnet_v24"); labels = synthetic_labels(spec.num_species)
B = 256
audio = synth.batch(0, B, 144000, 48000); segs = list(audio)
print("cpus", os.cpu_count(), "torch threads", torch.get_num_threads())

# 1. raw pinned H2D / D2H bandwidth
h = torch.empty(B * 144000, dtype=torch.float32).pin_memory()
d = torch.empty(B * 144000, dtype=torch.float32, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"pinned H2D 147 MB: {ms:.3f} ms -> {h.numel()*4/ms/1e6:.1f} GB/s")
e0.record()
for _ in range(5): h.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"pinned D2H 147 MB: {ms:.3f} ms -> {h.numel()*4/ms/1e6:.1f} GB/s")

# 2. host gather into pinned memory, 1 thread (numpy) and the engine's own packers via a full call
hp = h.numpy().reshape(B, 144000)
t = time.perf_counter()
for _ in range(3):
    for i in range(B): hp[i] = segs[i]
dt = (time.perf_counter() - t) / 3
print(f"python 1-thread gather: {dt*1e3:.2f} ms -> {audio.nbytes/dt/1e9:.1f} GB/s")

clf = bb.Classifier.builder().model_path(path).labels(labels).top_k(5).min_confidence(0.1).build()
ctx = clf.create_batch_context(B)
for _ in range(3): clf.predict_batch_with_context(ctx, segs)
# 3. one call, split: python marshalling / C call / python results
n = 8
t_m = t_c = t_r = 0.0
import ctypes as C
from birdnet_b200 import _ffi
for _ in range(n):
    t0 = time.perf_counter()
    ptrs, lens, keep = _segment_arrays(segs)
    t1 = time.perf_counter()
    out = _ffi.Outputs()
    st = _ffi.lib.bn_ctx_run(ctx._h, ptrs, lens, len(segs), None, C.byref(out))
    t2 = time.perf_counter()
    res = clf._results(out)
    t3 = time.perf_counter()
    t_m += t1 - t0; t_c += t2 - t1; t_r += t3 - t2
print(f"per call: marshal {t_m/n*1e3:.2f} ms | C call (gather+H2D+kernels+D2H) {t_c/n*1e3:.2f} ms | python results {t_r/n*1e3:.2f} ms")
ctx.set_profiling(True)
clf.predict_batch_with_context(ctx, segs)
st = ctx.stage_times()
tot = sum(ms for _, ms in st)
print("  device timeline of one call: total %.2f ms; h2d stage %.2f ms; d2h %.2f ms" %
      (tot, dict(st).get("h2d", 0.0), dict(st).get("d2h", 0.0)))
ctx.set_profiling(False)

# 4. throughput vs depth
for depth in (1, 2, 3, 4, 6):
    ctxs = [clf.create_batch_context(B) for _ in range(depth)]
    for c in ctxs: clf.predict_batch_with_context(c, segs)
    n = 8 * depth
    def work(t):
        for i in range(t, n, depth): clf.predict_batch_with_context(ctxs[t], segs)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(depth)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.perf_counter() - t0
    print(f"depth={depth}: {n*B/dt:8.0f} seg/s  ({dt/n*1e3:.2f} ms/batch)")
    del ctxs
