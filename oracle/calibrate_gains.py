"""Dev tool (run once, result committed): LSUV-style per-layer gains for the random-init graphs.

For each weighted layer in graph order, the gain is chosen so the layer's pre-activation
(bias excluded) has a target standard deviation over a calibration batch of the synthetic
audio distribution; gains are rounded to 4 significant digits so the frozen JSON (not this
script's floating-point noise) defines the weights.  Output:
rust-birdnet-onnx_b200/birdnet_b200/modelgen/calib_<family>.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, _ROOT)

from birdnet_b200.modelgen import get_spec, make_weights, synth  # noqa: E402
from birdnet_b200.modelgen.onnx_writer import build_model_bytes, parse_model  # noqa: E402
from oracle.model_oracle import ModelOracle  # noqa: E402


def calibrate(family: str, n_cal: int = 20, classifier_std: float = 2.0) -> dict:
    spec = get_spec(family)
    spec.gains, spec.biases = {}, {}
    w_np = make_weights(spec, gains={}, biases={})
    inits = parse_model(build_model_bytes(spec, w_np))["initializers"]
    orc = ModelOracle(spec, inits, dtype=torch.float64)
    fe = spec.frontend
    audio = synth.batch(0, n_cal, fe.sample_count, fe.sample_rate)
    with torch.no_grad():
        t = orc.frontend(torch.from_numpy(audio).double())
        gains, biases = {}, {}
        for op in spec.ops:
            k = op["op"]
            if k == "conv":
                n = op["name"]
                w = orc.w[f"{n}.weight"]
                y0 = F.conv2d(t[op["in"]], w, None, stride=op["stride"], padding=op["pad"],
                              groups=op["groups"])
                g = float(f"{1.0 / max(float(y0.std()), 1e-12):.4g}")
                gains[n] = g
                # folded-BN style centring: per-channel mean over the calibration batch
                mu = (y0 * g).mean(dim=(0, 2, 3))
                cb = np.round((-mu).numpy(), 3).astype(np.float32)
                biases[n] = cb
                y = y0 * g + (orc.w[f"{n}.bias"] + torch.from_numpy(cb).double()).view(1, -1, 1, 1)
                if op["act"] == "silu":
                    y = y * torch.sigmoid(y)
                elif op["act"] == "sigmoid":
                    y = torch.sigmoid(y)
                t[op["out"]] = y
            elif k == "add":
                t[op["out"]] = t[op["a"]] + t[op["b"]]
            elif k == "mul":
                t[op["out"]] = t[op["a"]] * t[op["b"]]
            elif k == "gap":
                t[op["out"]] = t[op["in"]].mean(dim=(2, 3), keepdim=True)
            elif k == "flatten":
                t[op["out"]] = t[op["in"]].flatten(1)
            elif k == "to_nhwc":
                t[op["out"]] = t[op["in"]]
            elif k == "gemm":
                n = op["name"]
                y0 = t[op["in"]] @ orc.w[f"{n}.weight"].t()
                g = float(f"{classifier_std / max(float(y0.std()), 1e-12):.4g}")
                gains[n] = g
                t[op["out"]] = y0 * g + orc.w[f"{n}.bias"]
    return gains, biases


if __name__ == "__main__":
    fams = sys.argv[1:] or ["birdnet_v24", "birdnet_v30", "perch_v2"]
    for fam in fams:
        g, b = calibrate(fam)
        p = os.path.join(_ROOT, "rust-birdnet-onnx_b200", "birdnet_b200", "modelgen",
                         f"calib_{fam}.npz")
        arrs = {f"gain:{k}": np.float64(v) for k, v in g.items()}
        arrs.update({f"bias:{k}": v for k, v in b.items()})
        np.savez_compressed(p, **arrs)
        print(fam, "->", p, "min/max gain", min(g.values()), max(g.values()))
