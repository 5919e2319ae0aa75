"""Device-resident throughput with one vs two private compute streams (dev aid).  BN_SHARED_COMPUTE=0 python tools/two_stream_test.py"""
import os, sys, time
os.environ.setdefault("BN_SHARED_COMPUTE", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
spec = get_spec("birdnet_v24")
clf = bb.Classifier.builder().model_path(ensure_model("birdnet_v24")).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
for B in (256, 128):
    audio = synth.batch(0, B, 144000, 48000)
    d = torch.from_numpy(audio).cuda()
    ctxs = [clf.create_batch_context(B) for _ in range(2)]
    for c in ctxs:
        for _ in range(3): c.run_device(d.data_ptr(), B, True)
    for n_streams in (1, 2):
        torch.cuda.synchronize(); t = time.perf_counter()
        K = 40
        for i in range(K):
            ctxs[i % n_streams].enqueue_device(d.data_ptr(), B, True)
        for c in ctxs[:n_streams]: c.wait()
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"B={B} streams={n_streams}: {K*B/dt:.0f} seg/s ({dt/K*1e3:.3f} ms/batch)", flush=True)
    del ctxs
