"""Build-authored BirdNET range-filter "meta model" (SURVEY.md section 8f row 2).

The reference loads it with ONNX Runtime and relies on exactly this contract (src/rangefilter.rs:239-267,
451-479): first input ``[1, 3]`` = ``[latitude, longitude, week]`` f32; exactly ONE output whose last dimension is
the label count; output values are compared directly against the threshold, so they are probabilities.  The real
file is not shipped with the reference; this module writes a deterministic random-init stand-in with that contract:

    x [1,3] -> scale (1/90, 1/180, 1/48) -> Gemm 3->128 + Relu -> Gemm 128->256 + Relu -> Gemm 256->N -> Sigmoid
"""
from __future__ import annotations

import hashlib
import os
from typing import Dict

import numpy as np

from .onnx_writer import _G, _f_bytes, _f_str, _f_varint, _value_info

_HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_DIR = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "..", "models")
HIDDEN = (128, 256)


def meta_weights(num_species: int, seed: int = 0) -> Dict[str, np.ndarray]:
    w = {}
    dims = (3,) + HIDDEN + (num_species,)
    for i in range(len(dims) - 1):
        sub = int.from_bytes(hashlib.sha256(f"meta{i}:{seed}".encode()).digest()[:8], "little")
        rng = np.random.Generator(np.random.PCG64(sub))
        fan_in = dims[i]
        w[f"fc{i}.weight"] = (rng.standard_normal((dims[i + 1], fan_in)) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        w[f"fc{i}.bias"] = (rng.standard_normal(dims[i + 1]) * 0.1).astype(np.float32)
    # spread the output probabilities over (0, 1): some species well below the default 0.01 threshold, some above
    w[f"fc{len(dims) - 2}.weight"] *= np.float32(1.5)
    last = f"fc{len(dims) - 2}.bias"           # a per-species prior, so every location has species on both sides
    w[last] = (w[last] * np.float32(15.0) - np.float32(2.0)).astype(np.float32)
    w["in_scale"] = np.array([1.0 / 90.0, 1.0 / 180.0, 1.0 / 48.0], dtype=np.float32)
    return w


def build_meta_model_bytes(num_species: int, seed: int = 0) -> bytes:
    w = meta_weights(num_species, seed)
    g = _G()
    g.init("in_scale", w["in_scale"])
    x = g.node("Mul", ["input", "in_scale"], ["scaled"])
    n_layers = len(HIDDEN) + 1
    for i in range(n_layers):
        g.init(f"fc{i}.weight", w[f"fc{i}.weight"])
        g.init(f"fc{i}.bias", w[f"fc{i}.bias"])
        out = f"fc{i}.out"
        g.node("Gemm", [x, f"fc{i}.weight", f"fc{i}.bias"], [out], alpha=1.0, beta=1.0, transB=1)
        if i < n_layers - 1:
            x = g.node("Relu", [out], [f"fc{i}.relu"])
        else:
            x = g.node("Sigmoid", [out], ["scores"])
    graph = b"".join(_f_bytes(1, n) for n in g.nodes)
    graph += _f_str(2, f"meta_model_like_seed{seed}")
    graph += b"".join(_f_bytes(5, t) for t in g.inits)
    graph += _f_bytes(11, _value_info("input", [1, 3]))
    graph += _f_bytes(12, _value_info("scores", [1, num_species]))
    model = _f_varint(1, 8) + _f_str(2, "birdnet_b200.modelgen") + _f_str(3, "1")
    model += _f_bytes(7, graph)
    model += _f_bytes(8, _f_str(1, "") + _f_varint(2, 17))
    return model


def meta_model_path(num_species: int = 6522, seed: int = 0) -> str:
    return os.path.normpath(os.path.join(MODEL_DIR, f"meta_model_{num_species}_seed{seed}.onnx"))


def ensure_meta_model(num_species: int = 6522, seed: int = 0) -> str:
    p = meta_model_path(num_species, seed)
    if not os.path.exists(p):
        os.makedirs(os.path.dirname(p), exist_ok=True)
        tmp = f"{p}.{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            f.write(build_meta_model_bytes(num_species, seed))
        os.replace(tmp, p)
    return p
