"""Generates tests/golden/meta_seed0.npz with the numpy oracle of the range-filter meta model (run in the build
container; the .npz is committed).  PARITY UNPINNED by the reference for the MLP arithmetic (it ships no meta-model file):
these vectors pin *this build's* stand-in model and oracle, so drift in either is caught on any machine."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, ROOT)

from birdnet_b200.modelgen import meta_model as mm  # noqa: E402
from oracle import meta_oracle as mo  # noqa: E402
from oracle.model_oracle import load_initializers  # noqa: E402

PLACES = [(60.17, 24.94, 6, 15), (-33.87, 151.21, 12, 31), (0.0, 0.0, 1, 1), (90.0, -180.0, 2, 8), (-90.0, 180.0, 7, 22)]

if __name__ == "__main__":
    path = mm.ensure_meta_model(6522, 0)
    w = load_initializers(path)
    scores = np.stack([mo.forward(w, np.float32(la), np.float32(lo), np.float32(mo.calculate_week(m, d)))
                       for la, lo, m, d in PLACES])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "meta_seed0.npz"),
                        places=np.array(PLACES, dtype=np.float64), scores_every_8=scores[:, ::8].astype(np.float32),
                        n_above_default_threshold=(scores >= np.float32(0.01)).sum(axis=1),
                        argmax=scores.argmax(axis=1),
                        model_sha256=np.frombuffer(hashlib.sha256(open(path, "rb").read()).digest(), dtype=np.uint8))
    print("wrote golden for the meta model:", scores.shape)
