"""Role cycle counters of the fused v2.4 front-end kernel (CTA 0): python tools/fe_phase_profile.py [batch]"""
import os, sys, ctypes as C
os.environ["BN_FE_PROFILE"] = "1"
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
st = dict(ctx.stage_times())
buf = (C.c_ulonglong * (128 * 16))()
fn = _ffi.lib.bn_debug_tc_profile
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
assert fn(buf, 128) == 0
r = np.frombuffer(buf, dtype=np.uint64).reshape(128, 16)[127].astype(np.float64) / 1e3
print(f"stage ms: normalize {st.get('normalize', 0):.4f} spectrogram {st.get('spectrogram', 0):.4f}")
print(f"CTA 0, kilo-cycles over {int(r[4]*1e3)} tiles")
print(f"  control : wait accumulator free {r[0]:.1f} | wait patch columns {r[1]:.1f} | wait basis stage {r[2]:.1f} | total {r[3]:.1f}")
print(f"  producer: wait columns free {r[5]:.1f} | total {r[6]:.1f}")
print(f"  epilogue: wait MMAs {r[7]:.1f} | total {r[8]:.1f}")
print(f"  loader  : wait ring slot {r[9]:.1f} | total {r[10]:.1f}")
