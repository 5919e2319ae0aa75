"""Where does the page-locked e2e path lose against the device-resident step?  (dev aid)
1. raw bn_ctx_run from `depth` threads (ctypes releases the GIL; no Python result building)
2. the public API (predict_batch_with_context) from the same threads
3. device-resident steps with and without a concurrent 147 MB H2D stream (does DMA slow the kernels?)"""
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
import birdnet_b200 as bb
from birdnet_b200 import _ffi
from birdnet_b200.classifier import _segment_arrays
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels

spec = get_spec("birdnet_v24"); path = ensure_model("birdnet_v24")
clf = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species))
       .top_k(5).min_confidence(0.1).pack_threads(8).build())
B = 256
audio = synth.batch(0, B, 144000, 48000)
pinned = bb.pinned_array(audio.shape); pinned[:] = audio
segs = list(pinned)
DEPTHS = [int(x) for x in os.environ.get("DEPTHS", "1,2,3,4,6,8").split(",")]
ctxs = [clf.create_batch_context(B) for _ in range(max(DEPTHS))]
for c in ctxs:
    clf.predict_batch_with_context(c, segs)


def timed(depth, fn, n=48):
    def work(t):
        for _ in range(t, n, depth):
            fn(t)
    th = [threading.Thread(target=work, args=(t,)) for t in range(depth)]
    t0 = time.perf_counter()
    [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n * B / dt, dt / n * 1e3


ptrs, lens, keep = _segment_arrays(segs)
outs = [_ffi.Outputs() for _ in ctxs]


def raw(t):
    st = _ffi.lib.bn_ctx_run(ctxs[t]._h, ptrs, lens, B, None, C.byref(outs[t]))
    assert st == 0


def api(t):
    clf.predict_batch_with_context(ctxs[t], segs)


for d in DEPTHS:
    r, ms = timed(d, raw)
    print(f"raw C call depth={d}: {r:8.0f} seg/s ({ms:.2f} ms/batch)", flush=True)
for d in DEPTHS:
    r, ms = timed(d, api)
    print(f"public API depth={d}: {r:8.0f} seg/s ({ms:.2f} ms/batch)", flush=True)

held = [None] * len(ctxs)


def api_hold(t):
    held[t] = clf.predict_batch_with_context(ctxs[t], segs)


for d in DEPTHS:
    r, ms = timed(d, api_hold)
    print(f"public API, result held until the next one, depth={d}: {r:8.0f} seg/s ({ms:.2f} ms/batch)", flush=True)
for d in DEPTHS:
    r, ms = timed(d, api, n=16 * d)
    print(f"public API n=16*depth depth={d}: {r:8.0f} seg/s ({ms:.2f} ms/batch)", flush=True)

# device-resident steps, with / without concurrent DMA
d_audio = torch.from_numpy(audio).cuda()
ctx = ctxs[0]
stream = torch.cuda.ExternalStream(ctx.stream_ptr())


def dev_steps(k=20):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(k):
        ctx.enqueue_device(d_audio.data_ptr(), B, True)
    ev1.record(stream)
    ctx.wait()
    return ev0.elapsed_time(ev1) / k


dev_steps(3)
print(f"device-resident alone: {dev_steps():.3f} ms/batch", flush=True)
hp = torch.from_numpy(pinned)
dst = torch.empty_like(d_audio)
side = torch.cuda.Stream()
stop = False


def dma_loop():
    with torch.cuda.stream(side):
        while not stop:
            for _ in range(4):
                dst.copy_(hp, non_blocking=True)
            side.synchronize()


th = threading.Thread(target=dma_loop); th.start()
time.sleep(0.05)
print(f"device-resident with a saturating H2D stream beside it: {dev_steps(40):.3f} ms/batch", flush=True)
stop = True; th.join()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(side):
    e0.record(side)
    for _ in range(8):
        dst.copy_(hp, non_blocking=True)
    e1.record(side)
side.synchronize()
print(f"H2D alone: {e0.elapsed_time(e1) / 8:.3f} ms per 147 MB", flush=True)

for d in DEPTHS:
    r, ms = timed(d, api)
    print(f"public API again at the end depth={d}: {r:8.0f} seg/s ({ms:.2f} ms/batch)", flush=True)
