"""Generates tests/golden/{v24,v30,perch}_seed0.npz with the torch-CPU FP32 oracle (run in the build
container; the .npz is committed).  PARITY UNPINNED by the reference for model numerics: these
vectors pin *this build's* oracle so that drift in graph spec, weights, synthetic audio or the
oracle itself is caught on any machine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, ROOT)

from birdnet_b200.modelgen import get_spec, synth  # noqa: E402
from birdnet_b200.modelgen.make_models import ensure_model  # noqa: E402
from oracle.model_oracle import ModelOracle, load_initializers  # noqa: E402
from oracle import postprocess_oracle as po  # noqa: E402

N_SEG = {"birdnet_v24": 20, "birdnet_v30": 10, "perch_v2": 10}
FILES = {"birdnet_v24": "v24_seed0.npz", "birdnet_v30": "v30_seed0.npz", "perch_v2": "perch_seed0.npz"}

if __name__ == "__main__":
    import hashlib
    for fam in (sys.argv[1:] or list(FILES)):
        spec = get_spec(fam)
        path = ensure_model(fam)
        n = N_SEG[fam]
        audio = synth.batch(0, n, spec.frontend.sample_count, spec.frontend.sample_rate)
        orc = ModelOracle(spec, load_initializers(path))
        out = orc.forward(audio, keep=["spec"])
        logits, emb = orc.logits_and_embeddings(audio)
        idx, conf, counts = po.top_k_batch(logits, 5, 0.1)
        extra = {}
        if emb is not None:
            extra = dict(emb_every_8=emb[:, ::8].astype(np.float32), emb_absmax=np.abs(emb).max(axis=1))
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", FILES[fam]),
            logits_every_32=logits[:, ::32].astype(np.float32),
            logits_absmax=np.abs(logits).max(axis=1),
            top5_idx=idx, top5_conf=conf, top5_count=counts,
            spec_mean=out["spec"].mean(axis=(1, 2, 3)), spec_max=out["spec"].max(axis=(1, 2, 3)),
            audio_sha256=np.frombuffer(hashlib.sha256(audio.tobytes()).digest(), dtype=np.uint8),
            model_sha256=np.frombuffer(hashlib.sha256(open(path, "rb").read()).digest(), dtype=np.uint8), **extra)
        print("wrote golden for", fam, n, "segments")
