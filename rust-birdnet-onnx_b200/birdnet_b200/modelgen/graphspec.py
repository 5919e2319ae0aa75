"""Build-authored graph specs for the three model families the reference drives.

The reference ships no model files; the graphs below are this
build's own "v2.4-like", "v3.0-like" and "Perch-v2-like" networks.  They honour every contract
the reference code relies on:

* input tensor name ``input`` with shape ``[B, sample_count]``      (batch_context.rs:221-223)
* BirdNET v2.4: one output ``output`` ``[B, 6522]`` raw logits        (batch_context.rs:248-250,
  detection.rs:31-41)
* BirdNET v3.0: ``output_0`` ``[B, 1024]`` embeddings, ``output_1`` ``[B, N]`` logits
  (batch_context.rs:252-262, detection.rs:44-56)
* Perch v2: four outputs ``[B,1536]``, ``[B,16,4,1536]``, ``[B,500,128]``, ``[B,14795]``
  accessed by index (detection.rs:58-71, classifier.rs:929-934)

The IR is a flat list of ops over named tensors (NCHW semantics, like ONNX).  Three consumers
read it: ``onnx_writer`` (serialises to a real opset-17 ONNX file), the torch oracle
(``oracle/model_oracle.py``) and the tests.  The CUDA engine never sees this module: it parses
the ONNX file.

Weights are drawn from a seeded numpy PCG64 stream per tensor (name-keyed sub-seed), scaled by
a per-layer gain frozen in ``gains_<family>.json`` (LSUV-style calibration, produced once by
``calibrate.py`` and committed) so that activations stay O(1) through the ~60 layers.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# Front-end specs
# --------------------------------------------------------------------------------------
@dataclass
class MelSpec:
    """One spectrogram branch of the in-graph audio front-end."""
    n_fft: int          # frame length == FFT size (no centre padding)
    hop: int
    fmin: float
    fmax: float
    n_mels: int
    mag_scale: float = 1.23   # BirdNET v2.4: pow(x, 1/(1+exp(mag_scale)))

    def n_frames(self, sample_count: int) -> int:
        return 1 + (sample_count - self.n_fft) // self.hop

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1


@dataclass
class FrontEnd:
    kind: str                  # "birdnet_v24" | "logmel"
    sample_rate: int
    sample_count: int
    specs: List[MelSpec]
    # logmel only: y = log_scale * ln(mel(|STFT|) + log_floor), layout [frames, mels]
    pad_end: int = 0           # zeros appended so the frame count is exact
    # 1e-2 against mel magnitudes that reach ~250 for a full-scale tone (88 dB of range): with 1e-5 the FP32 FFT
    # noise floor sat ABOVE the log floor and ln() amplified it - the FP32 and FP64 oracles then differed by 0.24 in
    # the logits of the chirp segment; with 1e-2 they agree to 7e-5 on all ten synthetic kinds (build-authored
    # constant: the reference fixes none)
    log_floor: float = 1e-2
    log_scale: float = 0.1

    def n_frames(self) -> int:
        s = self.specs[0]
        return 1 + (self.sample_count + self.pad_end - s.n_fft) // s.hop


@dataclass
class GraphSpec:
    family: str                # "birdnet_v24" | "birdnet_v30" | "perch_v2"
    seed: int
    frontend: FrontEnd
    ops: List[dict]
    outputs: List[dict]        # [{"name":..., "tensor":..., "shape":[...]}]
    num_species: int
    embedding_dim: Optional[int]
    gains: Dict[str, float] = field(default_factory=dict)
    biases: Dict[str, np.ndarray] = field(default_factory=dict)

    # -- bookkeeping the roofline tables are derived from ---------------------------
    def layer_table(self) -> List[dict]:
        """Per weighted layer: MACs and activation element counts per segment."""
        rows = []
        shapes = {}
        fe = self.frontend
        if fe.kind == "birdnet_v24":
            shapes["spec"] = (len(fe.specs), fe.specs[0].n_mels, fe.n_frames())
        else:
            shapes["spec"] = (1, fe.n_frames(), fe.specs[0].n_mels)
        for op in self.ops:
            k = op["op"]
            if k == "conv":
                c, h, w = shapes[op["in"]]
                ho = (h + 2 * op["pad"] - op["k"]) // op["stride"] + 1
                wo = (w + 2 * op["pad"] - op["k"]) // op["stride"] + 1
                shapes[op["out"]] = (op["cout"], ho, wo)
                macs = ho * wo * op["cout"] * (op["cin"] // op["groups"]) * op["k"] ** 2
                rows.append(dict(name=op["name"], kind="dw" if op["groups"] > 1 else "conv",
                                 k=op["k"], stride=op["stride"], cin=op["cin"], cout=op["cout"],
                                 hin=h, win=w, hout=ho, wout=wo, macs=macs,
                                 in_elems=c * h * w, out_elems=op["cout"] * ho * wo,
                                 w_elems=op["cout"] * (op["cin"] // op["groups"]) * op["k"] ** 2))
            elif k == "add":
                shapes[op["out"]] = shapes[op["a"]]
            elif k == "mul":
                shapes[op["out"]] = shapes[op["a"]]
            elif k == "gap":
                c, h, w = shapes[op["in"]]
                shapes[op["out"]] = (c, 1, 1)
            elif k == "flatten":
                c, h, w = shapes[op["in"]]
                shapes[op["out"]] = (c * h * w, 1, 1)
            elif k == "gemm":
                shapes[op["out"]] = (op["cout"], 1, 1)
                rows.append(dict(name=op["name"], kind="gemm", k=1, stride=1, cin=op["cin"],
                                 cout=op["cout"], hin=1, win=1, hout=1, wout=1,
                                 macs=op["cin"] * op["cout"], in_elems=op["cin"],
                                 out_elems=op["cout"], w_elems=op["cin"] * op["cout"]))
            elif k == "to_nhwc":
                shapes[op["out"]] = shapes[op["in"]]
            else:
                raise ValueError(k)
        self._shapes = shapes
        return rows

    def frontend_macs(self) -> int:
        fe = self.frontend
        total = 0
        if fe.kind == "birdnet_v24":
            for s in fe.specs:
                # window*cos basis pre-multiplied with the mel matrix: [n_fft x n_mels] per frame
                total += s.n_frames(fe.sample_count) * s.n_fft * s.n_mels
        else:
            s = fe.specs[0]
            t = fe.n_frames()
            total += t * s.n_fft * 2 * s.n_bins + t * s.n_bins * s.n_mels
        return total


# --------------------------------------------------------------------------------------
# IR builders
# --------------------------------------------------------------------------------------
class _B:
    def __init__(self):
        self.ops: List[dict] = []
        self.n = 0

    def t(self, hint="t"):
        self.n += 1
        return f"{hint}_{self.n}"

    def conv(self, name, x, cin, cout, k, stride=1, groups=1, act="silu"):
        out = self.t(name)
        self.ops.append(dict(op="conv", name=name, cin=cin, cout=cout, k=k, stride=stride,
                             pad=k // 2, groups=groups, act=act, **{"in": x}, out=out))
        return out

    def add(self, a, b):
        out = self.t("add")
        self.ops.append(dict(op="add", a=a, b=b, out=out))
        return out

    def mul(self, a, b):
        out = self.t("mul")
        self.ops.append(dict(op="mul", a=a, b=b, out=out))
        return out

    def gap(self, x):
        out = self.t("gap")
        self.ops.append(dict(op="gap", **{"in": x}, out=out))
        return out

    def flatten(self, x):
        out = self.t("flat")
        self.ops.append(dict(op="flatten", **{"in": x}, out=out))
        return out

    def gemm(self, name, x, cin, cout, out=None):
        out = out or self.t(name)
        self.ops.append(dict(op="gemm", name=name, cin=cin, cout=cout, **{"in": x}, out=out))
        return out

    # EfficientNetV2-style fused block: 3x3 expand conv (+SiLU) -> 1x1 project (linear) [+res]
    def fused_mbconv(self, name, x, cin, cout, expand, stride):
        if expand == 1:
            y = self.conv(f"{name}.fused", x, cin, cout, 3, stride, act="silu")
        else:
            mid = cin * expand
            y = self.conv(f"{name}.fused", x, cin, mid, 3, stride, act="silu")
            y = self.conv(f"{name}.project", y, mid, cout, 1, 1, act="none")
        if stride == 1 and cin == cout:
            y = self.add(y, x)
        return y

    # EfficientNet-V1 MBConv + squeeze-excite
    def mbconv(self, name, x, cin, cout, expand, k, stride, se_ratio=0.25):
        mid = cin * expand
        y = self.conv(f"{name}.expand", x, cin, mid, 1, 1, act="silu")
        y = self.conv(f"{name}.dw", y, mid, mid, k, stride, groups=mid, act="silu")
        r = max(8, int(cin * se_ratio) // 4 * 4)
        s = self.gap(y)
        s = self.conv(f"{name}.se_reduce", s, mid, r, 1, 1, act="silu")
        s = self.conv(f"{name}.se_expand", s, r, mid, 1, 1, act="sigmoid")
        y = self.mul(y, s)
        y = self.conv(f"{name}.project", y, mid, cout, 1, 1, act="none")
        if stride == 1 and cin == cout:
            y = self.add(y, x)
        return y


def _backbone(b: _B, x: str, cin: int, stages, head_ch: int) -> str:
    c = cin
    for si, (kind, k, stride, expand, cout, reps) in enumerate(stages, start=1):
        for r in range(reps):
            st = stride if r == 0 else 1
            nm = f"s{si}b{r}"
            if kind == "fused":
                x = b.fused_mbconv(nm, x, c, cout, expand, st)
            else:
                x = b.mbconv(nm, x, c, cout, expand, k, st)
            c = cout
    x = b.conv("head_conv", x, c, head_ch, 1, 1, act="silu")
    return x


# Channel widths are multiples of 16 so every dense layer maps onto UMMA N granularity and
# 16-byte NHWC fp16 vectors without padding (B0 uses 24/40; we use 32/48).
_V24_STAGES = [
    # kind     k  s  e  cout reps
    ("fused", 3, 1, 1, 16, 1),
    ("fused", 3, 2, 4, 32, 2),
    ("fused", 3, 2, 4, 48, 2),
    ("mb",    3, 2, 6, 80, 3),
    ("mb",    5, 1, 6, 112, 3),
    ("mb",    5, 2, 6, 192, 4),
    ("mb",    3, 1, 6, 320, 1),
]


def birdnet_v24_spec(seed: int = 0, num_species: int = 6522) -> GraphSpec:
    fe = FrontEnd(kind="birdnet_v24", sample_rate=48000, sample_count=144000,
                  specs=[MelSpec(2048, 278, 0.0, 3000.0, 96),
                         MelSpec(1024, 280, 500.0, 15000.0, 96)])
    b = _B()
    x = b.conv("stem", "spec", 2, 32, 3, 2, act="silu")
    x = _backbone(b, x, 32, _V24_STAGES, 1024)
    x = b.gap(x)
    x = b.flatten(x)
    b.gemm("classifier", x, 1024, num_species, out="output")
    g = GraphSpec("birdnet_v24", seed, fe, b.ops,
                  [dict(name="output", tensor="output", shape=["batch", num_species])],
                  num_species, None)
    g.gains, g.biases = _load_calibration("birdnet_v24")
    return g


def birdnet_v30_spec(seed: int = 0, num_species: int = 11560) -> GraphSpec:
    """32 kHz, 5 s, log-mel front-end (north star), 1024-d embedding + logits."""
    fe = FrontEnd(kind="logmel", sample_rate=32000, sample_count=160000,
                  specs=[MelSpec(1024, 320, 40.0, 15000.0, 128)],
                  pad_end=0, log_floor=1e-2, log_scale=0.1)
    b = _B()
    x = b.conv("stem", "spec", 1, 32, 3, 2, act="silu")
    x = _backbone(b, x, 32, _V24_STAGES, 1024)
    x = b.gap(x)
    emb = b.flatten(x)
    b.ops[-1]["out"] = "output_0"
    b.gemm("classifier", "output_0", 1024, num_species, out="output_1")
    g = GraphSpec("birdnet_v30", seed, fe, b.ops,
                  [dict(name="output_0", tensor="output_0", shape=["batch", 1024]),
                   dict(name="output_1", tensor="output_1", shape=["batch", num_species])],
                  num_species, 1024)
    g.gains, g.biases = _load_calibration("birdnet_v30")
    return g


_PERCH_STAGES = [
    ("fused", 3, 1, 1, 32, 1),
    ("fused", 3, 2, 4, 48, 2),
    ("fused", 3, 2, 4, 64, 2),
    ("mb",    3, 2, 6, 128, 3),
    ("mb",    5, 1, 6, 160, 3),
    ("mb",    5, 2, 6, 256, 4),
    ("mb",    3, 1, 6, 384, 1),
]


def perch_v2_spec(seed: int = 0, num_species: int = 14795) -> GraphSpec:
    """32 kHz, 5 s, log-mel 500x128 (20 ms window / 10 ms hop), 1536-d embedding.

    Outputs follow detection.rs:214-232: embedding, spatial embedding [B,16,4,1536],
    spectrogram [B,500,128], logits.  500/32 -> 16 and 128/32 -> 4 need 5 stride-2 steps.
    """
    fe = FrontEnd(kind="logmel", sample_rate=32000, sample_count=160000,
                  specs=[MelSpec(640, 320, 60.0, 16000.0, 128)],
                  pad_end=320, log_floor=1e-2, log_scale=0.1)
    b = _B()
    x = b.conv("stem", "spec", 1, 32, 3, 2, act="silu")
    x = _backbone(b, x, 32, _PERCH_STAGES, 1536)
    feat = x
    x = b.gap(x)
    b.flatten(x)
    b.ops[-1]["out"] = "embedding"
    b.ops.append(dict(op="to_nhwc", **{"in": feat}, out="spatial_embedding"))
    b.gemm("classifier", "embedding", 1536, num_species, out="logits")
    g = GraphSpec("perch_v2", seed, fe, b.ops,
                  [dict(name="embedding", tensor="embedding", shape=["batch", 1536]),
                   dict(name="spatial_embedding", tensor="spatial_embedding",
                        shape=["batch", 16, 4, 1536]),
                   dict(name="spectrogram", tensor="spectrogram", shape=["batch", 500, 128]),
                   dict(name="logits", tensor="logits", shape=["batch", num_species])],
                  num_species, 1536)
    g.gains, g.biases = _load_calibration("perch_v2")
    return g


SPECS = {"birdnet_v24": birdnet_v24_spec, "birdnet_v30": birdnet_v30_spec,
         "perch_v2": perch_v2_spec}


def get_spec(family: str, seed: int = 0, **kw) -> GraphSpec:
    return SPECS[family](seed=seed, **kw)


# --------------------------------------------------------------------------------------
# Front-end constants (float64 maths, cast to f32 at the end)
# --------------------------------------------------------------------------------------
def hann_window(n: int) -> np.ndarray:
    """Periodic Hann, as tf.signal.hann_window(periodic=True)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)


def _hz_to_mel(f):
    return 1127.0 * np.log1p(np.asarray(f, dtype=np.float64) / 700.0)


def mel_matrix(n_mels: int, n_bins: int, sample_rate: int, fmin: float, fmax: float) -> np.ndarray:
    """HTK mel weights [n_bins, n_mels], as tf.signal.linear_to_mel_weight_matrix."""
    nyq = sample_rate / 2.0
    lin = np.linspace(0.0, nyq, n_bins)[1:]          # DC bin excluded (zero row)
    mel_f = _hz_to_mel(lin)[:, None]
    edges = np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2)
    lo, ce, hi = edges[:-2][None, :], edges[1:-1][None, :], edges[2:][None, :]
    lower = (mel_f - lo) / (ce - lo)
    upper = (hi - mel_f) / (hi - ce)
    w = np.maximum(0.0, np.minimum(lower, upper))
    w = np.concatenate([np.zeros((1, n_mels)), w], axis=0)
    return w.astype(np.float32)


def mag_exponent(mag_scale: float) -> np.float32:
    return np.float32(1.0 / (1.0 + math.exp(mag_scale)))


# --------------------------------------------------------------------------------------
# Weights
# --------------------------------------------------------------------------------------
def _load_calibration(family: str):
    """Frozen calibration (oracle/calibrate_gains.py): per-layer gain + per-channel bias."""
    p = os.path.join(_HERE, f"calib_{family}.npz")
    gains, biases = {}, {}
    if os.path.exists(p):
        with np.load(p) as z:
            for k in z.files:
                kind, name = k.split(":", 1)
                if kind == "gain":
                    gains[name] = float(z[k])
                else:
                    biases[name] = np.array(z[k], dtype=np.float32)
    return gains, biases


def _rng(seed: int, name: str) -> np.random.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def make_weights(spec: GraphSpec, gains: Optional[Dict[str, float]] = None,
                 biases: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
    """Deterministic seeded weights: N(0,1)/sqrt(fan_in) * gain, small biases.

    Conv weights are ONNX layout [cout, cin/groups, k, k]; gemm weights [cout, cin].
    BatchNorm is assumed folded (eval-mode export does that).
    """
    gains = spec.gains if gains is None else gains
    biases = spec.biases if biases is None else biases
    w: Dict[str, np.ndarray] = {}
    for op in spec.ops:
        if op["op"] == "conv":
            n = op["name"]
            cpg = op["cin"] // op["groups"]
            fan_in = cpg * op["k"] ** 2
            g = gains.get(n, 1.0)
            w[f"{n}.weight"] = (_rng(spec.seed, n + ".weight").standard_normal(
                (op["cout"], cpg, op["k"], op["k"])) * (g / math.sqrt(fan_in))).astype(np.float32)
            bias = _rng(spec.seed, n + ".bias").standard_normal(op["cout"]) * 0.25
            if op["act"] == "sigmoid":           # SE gate: centre the gate around ~0.62
                bias = bias + 0.5
            if n in biases:                      # folded-BN style centring (frozen calibration)
                bias = bias + biases[n]
            w[f"{n}.bias"] = bias.astype(np.float32)
        elif op["op"] == "gemm":
            n = op["name"]
            g = gains.get(n, 1.0)
            w[f"{n}.weight"] = (_rng(spec.seed, n + ".weight").standard_normal(
                (op["cout"], op["cin"])) * (g / math.sqrt(op["cin"]))).astype(np.float32)
            # logits centred at -4 with a spread so a few classes clear min_confidence
            w[f"{n}.bias"] = (_rng(spec.seed, n + ".bias").standard_normal(op["cout"]) * 0.5
                              - 4.0).astype(np.float32)
    return w
