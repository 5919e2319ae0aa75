"""ORACLE (test infrastructure, not product): torch-CPU execution of the build-authored graphs.

PARITY UNPINNED for the model arithmetic: the reference's own implementation of this stage is
ONNX Runtime 1.22 via `ort` 2.0.0-rc.11 (Cargo.lock:654-670) executing model files it does not
ship; neither can run in this image and no reference test pins a model
output.  This module is the stand-in the north star's "ORT CPU path" is replaced by: the same
graph (weights and front-end constants are read back from the .onnx file the CUDA engine loads)
executed op-by-op with torch CPU kernels in FP32 (or FP64 to quantify FP32 noise).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, "rust-birdnet-onnx_b200"))

from birdnet_b200.modelgen.graphspec import GraphSpec  # noqa: E402
from birdnet_b200.modelgen.onnx_writer import parse_model  # noqa: E402


def load_initializers(onnx_path: str) -> Dict[str, np.ndarray]:
    with open(onnx_path, "rb") as f:
        return parse_model(f.read())["initializers"]


class ModelOracle:
    """Executes GraphSpec.ops with weights taken from an ONNX file's initializers."""

    def __init__(self, spec: GraphSpec, inits: Dict[str, np.ndarray], dtype=torch.float32,
                 round_fp16: bool = False):
        self.spec = spec
        self.dtype = dtype
        # round_fp16: emulate an engine that stores activations and weights in fp16 with
        # fp32 accumulation (policy study only)
        self.round_fp16 = round_fp16
        self.w = {k: torch.from_numpy(np.array(v)).to(dtype) if v.dtype == np.float32
                  else torch.from_numpy(np.array(v)) for k, v in inits.items()}
        if round_fp16:
            for k in list(self.w):
                if k.endswith(".weight"):
                    self.w[k] = self.w[k].half().to(dtype)

    def _q(self, t):
        return t.half().to(self.dtype) if self.round_fp16 else t

    # ---- front-ends -----------------------------
    def frontend(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        fe = self.spec.frontend
        w = self.w
        out = {}
        if fe.kind == "birdnet_v24":
            mn = x.min(dim=1, keepdim=True).values
            x1 = x - mn
            mx = x1.max(dim=1, keepdim=True).values
            x = (x1 / (mx + w["fe.eps"]) - w["fe.half"]) * w["fe.two"]
            out["normalized"] = x
            branches = []
            for i, s in enumerate(fe.specs):
                p = f"fe.spec{i}"
                st = torch.stft(x, n_fft=s.n_fft, hop_length=s.hop, win_length=s.n_fft,
                                window=w[f"{p}.window"], center=False, onesided=True,
                                return_complex=True)
                re = st.real.transpose(1, 2)                      # [B, frames, bins]
                m = re @ w[f"{p}.mel"]
                m = m * m
                m = torch.pow(m, w[f"{p}.exponent"])
                m = m.flip(2).transpose(1, 2)                     # [B, mels, frames]
                branches.append(m.unsqueeze(1))
            out["spec"] = torch.cat(branches, dim=1)              # [B, 2, 96, 511]
        else:
            s = fe.specs[0]
            p = "fe.spec0"
            if fe.pad_end:
                x = F.pad(x, (0, fe.pad_end))
            st = torch.stft(x, n_fft=s.n_fft, hop_length=s.hop, win_length=s.n_fft,
                            window=w[f"{p}.window"], center=False, onesided=True,
                            return_complex=True)
            st = st.transpose(1, 2)
            mag = torch.sqrt(st.real * st.real + st.imag * st.imag)
            m = mag @ w[f"{p}.mel"]
            lm = torch.log(m + w["fe.log_floor"]) * w["fe.log_scale"]   # [B, frames, mels]
            out["spectrogram"] = lm
            out["spec"] = lm.unsqueeze(1)
        return out

    @torch.no_grad()
    def forward(self, audio: np.ndarray, keep: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
        x = torch.from_numpy(np.ascontiguousarray(audio)).to(self.dtype)
        t = self.frontend(x)
        t["spec"] = self._q(t["spec"])
        w = self.w
        for op in self.spec.ops:
            k = op["op"]
            if k == "conv":
                n = op["name"]
                y = F.conv2d(t[op["in"]], w[f"{n}.weight"], w[f"{n}.bias"], stride=op["stride"],
                             padding=op["pad"], groups=op["groups"])
                if op["act"] == "silu":
                    y = y * torch.sigmoid(y)
                elif op["act"] == "sigmoid":
                    y = torch.sigmoid(y)
                t[op["out"]] = self._q(y)
            elif k == "add":
                t[op["out"]] = self._q(t[op["a"]] + t[op["b"]])
            elif k == "mul":
                t[op["out"]] = self._q(t[op["a"]] * t[op["b"]])
            elif k == "gap":
                t[op["out"]] = t[op["in"]].mean(dim=(2, 3), keepdim=True)
            elif k == "flatten":
                t[op["out"]] = t[op["in"]].flatten(1)
            elif k == "to_nhwc":
                t[op["out"]] = t[op["in"]].permute(0, 2, 3, 1).contiguous()
            elif k == "gemm":
                n = op["name"]
                t[op["out"]] = t[op["in"]] @ w[f"{n}.weight"].t() + w[f"{n}.bias"]
            else:
                raise ValueError(k)
        names = [o["tensor"] for o in self.spec.outputs] + list(keep or [])
        return {n: t[n].to(torch.float32).numpy() if self.dtype != torch.float64
                else t[n].numpy() for n in names}

    def logits_and_embeddings(self, audio: np.ndarray):
        """(logits [B,N], embeddings [B,E] or None) following classifier.rs:917-934."""
        o = self.forward(audio)
        outs = self.spec.outputs
        if self.spec.family == "birdnet_v24":
            return o[outs[0]["tensor"]], None
        if self.spec.family == "birdnet_v30":
            return o[outs[1]["tensor"]], o[outs[0]["tensor"]]
        return o[outs[3]["tensor"]], o[outs[0]["tensor"]]
