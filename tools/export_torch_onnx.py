#!/usr/bin/env python
"""Independent producer of the model file: the build-authored graph as a torch.nn.Module, exported with torch's own
legacy TorchScript ONNX exporter (opset 17) instead of birdnet_b200.modelgen.onnx_writer.  Used by
tests/test_torch_export_dialect.py to check that the C++ reader / planner are not tied to the in-house writer.

    python tools/export_torch_onnx.py birdnet_v24 /tmp/v24_torch.onnx [--num-species 64]

The image has no `onnx` package; the exporter only imports it to attach onnxscript functions, which this graph does not
use, so that one helper is stubbed.
"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from birdnet_b200.modelgen import get_spec, make_weights


class SpecModule(nn.Module):
    """GraphSpec.ops over torch modules; front-end as in SURVEY.md appendix A.2 / A.3."""

    def __init__(self, spec, weights):
        super().__init__()
        self.spec = spec
        fe = spec.frontend
        from birdnet_b200.modelgen.graphspec import hann_window, mel_matrix, mag_exponent
        for i, s in enumerate(fe.specs):
            self.register_buffer(f"window{i}", torch.from_numpy(hann_window(s.n_fft)))
            self.register_buffer(f"mel{i}", torch.from_numpy(mel_matrix(s.n_mels, s.n_bins, fe.sample_rate, s.fmin, s.fmax)))
            self.exponent = float(mag_exponent(s.mag_scale))
        self.layers = nn.ModuleDict()
        for op in spec.ops:
            n = op["name"].replace(".", "_") if "name" in op else None
            if op["op"] == "conv":
                c = nn.Conv2d(op["cin"], op["cout"], op["k"], op["stride"], op["pad"], groups=op["groups"])
                c.weight.data = torch.from_numpy(weights[op["name"] + ".weight"])
                c.bias.data = torch.from_numpy(weights[op["name"] + ".bias"])
                self.layers[n] = c
            elif op["op"] == "gemm":
                l = nn.Linear(op["cin"], op["cout"])
                l.weight.data = torch.from_numpy(weights[op["name"] + ".weight"])
                l.bias.data = torch.from_numpy(weights[op["name"] + ".bias"])
                self.layers[n] = l

    def frontend(self, x):
        fe = self.spec.frontend
        if fe.kind == "birdnet_v24":
            mn = x.min(dim=1, keepdim=True).values
            x1 = x - mn
            mx = x1.max(dim=1, keepdim=True).values
            x = (x1 / (mx + 1e-6) - 0.5) * 2.0
            br = []
            for i, s in enumerate(fe.specs):
                st = torch.stft(x, n_fft=s.n_fft, hop_length=s.hop, win_length=s.n_fft, window=getattr(self, f"window{i}"),
                                center=False, onesided=True, return_complex=False)       # [B, bins, frames, 2]
                re = st[..., 0].transpose(1, 2)
                m = re @ getattr(self, f"mel{i}")
                m = m * m
                m = torch.pow(m, self.exponent)
                m = m.flip(2).transpose(1, 2)
                br.append(m.unsqueeze(1))
            return torch.cat(br, dim=1)
        s = fe.specs[0]
        if fe.pad_end:
            x = F.pad(x, (0, fe.pad_end))
        st = torch.stft(x, n_fft=s.n_fft, hop_length=s.hop, win_length=s.n_fft, window=self.window0, center=False,
                        onesided=True, return_complex=False)
        st = st.transpose(1, 2)
        mag = torch.sqrt(st[..., 0] * st[..., 0] + st[..., 1] * st[..., 1])
        m = mag @ self.mel0
        return torch.log(m + fe.log_floor) * fe.log_scale           # [B, frames, mels]

    def forward(self, x):
        f = self.frontend(x)
        t = {"spec": f}
        if self.spec.frontend.kind != "birdnet_v24":
            t = {"spectrogram": f, "spec": f.unsqueeze(1)}
        for op in self.spec.ops:
            k = op["op"]
            if k == "conv":
                y = self.layers[op["name"].replace(".", "_")](t[op["in"]])
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
                t[op["out"]] = t[op["in"]].permute(0, 2, 3, 1)
            elif k == "gemm":
                t[op["out"]] = self.layers[op["name"].replace(".", "_")](t[op["in"]])
        return tuple(t[o["tensor"]] for o in self.spec.outputs)


def export(family: str, path: str, **kw) -> str:
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes     # the only user of `onnx`
    spec = get_spec(family, **kw)
    mod = SpecModule(spec, make_weights(spec)).eval()
    x = torch.zeros(2, spec.frontend.sample_count)
    names = [o["name"] for o in spec.outputs]
    dyn = {"input": {0: "batch"}}
    dyn.update({n: {0: "batch"} for n in names})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(mod, (x,), path, input_names=["input"], output_names=names, dynamic_axes=dyn, opset_version=17,
                          dynamo=False, do_constant_folding=True)
    return path


if __name__ == "__main__":
    fam, out = sys.argv[1], sys.argv[2]
    kw = {}
    if "--num-species" in sys.argv:
        kw["num_species"] = int(sys.argv[sys.argv.index("--num-species") + 1])
    print(export(fam, out, **kw))
