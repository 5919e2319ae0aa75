"""One device-resident batch through the engine (profiling target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
B = int(os.environ.get("B", "256")); reps = int(os.environ.get("REPS", "2"))
spec = get_spec("birdnet_v24")
clf = bb.Classifier.builder().model_path(ensure_model("birdnet_v24")).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
audio = synth.batch(0, B, 144000, 48000)
d = torch.from_numpy(audio).cuda()
ctx = clf.create_batch_context(B)
for _ in range(reps):
    out = ctx.run_device(d.data_ptr(), B, True)
print("ok", out.batch, ctx.last_launch_count())
