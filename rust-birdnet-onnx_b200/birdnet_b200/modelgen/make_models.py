"""Write the deterministic random-init ONNX files + label files under <repo>/models/.

    python -m birdnet_b200.modelgen.make_models [family ...]

Files are reproducible bit for bit from (spec, seed, frozen calibration); they are git-ignored
and regenerated on demand by `ensure_model()`.
"""
from __future__ import annotations

import os
import sys

from .graphspec import get_spec
from .onnx_writer import write_model

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
MODEL_DIR = os.path.join(REPO, "models")


def model_path(family: str, seed: int = 0) -> str:
    return os.path.join(MODEL_DIR, f"{family}_seed{seed}.onnx")


def ensure_model(family: str = "birdnet_v24", seed: int = 0) -> str:
    p = model_path(family, seed)
    if not os.path.exists(p):
        os.makedirs(MODEL_DIR, exist_ok=True)
        write_model(get_spec(family, seed), p)
    return p


def synthetic_labels(n: int):
    """`Genus species_Common name i` in the v2.4 text format (data/labels/birdnet_v2.4/*.txt)."""
    return [f"Avis synthetica{i}_Synthetic Bird {i}" for i in range(n)]


if __name__ == "__main__":
    for fam in (sys.argv[1:] or ["birdnet_v24"]):
        print(ensure_model(fam))
