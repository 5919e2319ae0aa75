"""Minimal ONNX (opset 17) serialiser + parser — no `onnx`/`protobuf` package needed.

Only the protobuf wire features ONNX files use are implemented: varint (wire type 0),
64-bit (1), length-delimited (2) and 32-bit (5).  Field numbers follow onnx.proto3:

  ModelProto   : ir_version=1 producer_name=2 producer_version=3 graph=7 opset_import=8
  GraphProto   : node=1 name=2 initializer=5 input=11 output=12
  NodeProto    : input=1 output=2 name=3 op_type=4 attribute=5
  AttributeProto: name=1 f=2 i=3 s=4 t=5 floats=7 ints=8 type=20
  TensorProto  : dims=1 data_type=2 name=8 raw_data=9
  ValueInfoProto: name=1 type=2 ; TypeProto.tensor_type=1 ; Tensor: elem_type=1 shape=2
  TensorShapeProto.dim=1 ; Dimension: dim_value=1 dim_param=2

The emitted graph uses the op vocabulary a torch eval-mode export of the same module produces
(SURVEY.md appendix A.4): ReduceMin, ReduceMax, Sub, Add, Div, Mul, Unsqueeze, STFT, Gather,
MatMul, Pow, Slice, Transpose, Concat, Conv, Sigmoid, GlobalAveragePool, Flatten, Gemm (+ Pad,
Sqrt, Log for the log-mel families).  BatchNorm is folded into Conv, SiLU is Sigmoid+Mul.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

from .graphspec import GraphSpec, hann_window, mag_exponent, make_weights, mel_matrix

FLOAT, INT64 = 1, 7
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS = 1, 2, 3, 4, 6, 7


# ---------------------------------------------------------------- wire encoding
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def _f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def _f_str(field: int, s: str) -> bytes:
    return _f_bytes(field, s.encode("utf-8"))


def _f_f32(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", v)


def tensor_proto(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    if arr.dtype == np.float32:
        dt = FLOAT
    elif arr.dtype == np.int64:
        dt = INT64
    else:
        raise TypeError(arr.dtype)
    out = b"".join(_f_varint(1, d) for d in arr.shape)
    out += _f_varint(2, dt) + _f_str(8, name) + _f_bytes(9, arr.tobytes())
    return out


def _attr(name: str, v) -> bytes:
    out = _f_str(1, name)
    if isinstance(v, float):
        out += _f_f32(2, v) + _f_varint(20, A_FLOAT)
    elif isinstance(v, int):
        out += _f_varint(3, v) + _f_varint(20, A_INT)
    elif isinstance(v, str):
        out += _f_bytes(4, v.encode()) + _f_varint(20, A_STRING)
    elif isinstance(v, (list, tuple)) and all(isinstance(x, int) for x in v):
        out += b"".join(_f_varint(8, x) for x in v) + _f_varint(20, A_INTS)
    else:
        raise TypeError(f"attr {name}: {type(v)}")
    return out


def _node(op: str, ins: List[str], outs: List[str], name: str = "", **attrs) -> bytes:
    out = b"".join(_f_str(1, i) for i in ins) + b"".join(_f_str(2, o) for o in outs)
    out += _f_str(3, name or outs[0]) + _f_str(4, op)
    for k, v in attrs.items():
        out += _f_bytes(5, _attr(k, v))
    return out


def _value_info(name: str, shape) -> bytes:
    dims = b""
    for d in shape:
        dim = _f_str(2, d) if isinstance(d, str) else _f_varint(1, d)
        dims += _f_bytes(1, dim)
    ttype = _f_varint(1, FLOAT) + _f_bytes(2, dims)
    return _f_str(1, name) + _f_bytes(2, _f_bytes(1, ttype))


# ---------------------------------------------------------------- graph emission
class _G:
    def __init__(self):
        self.nodes: List[bytes] = []
        self.inits: List[bytes] = []
        self._names = set()

    def init(self, name: str, arr) -> str:
        assert name not in self._names, name
        self._names.add(name)
        self.inits.append(tensor_proto(name, np.asarray(arr)))
        return name

    def node(self, op, ins, outs, **attrs):
        self.nodes.append(_node(op, ins, outs, **attrs))
        return outs[0]


def _emit_frontend(g: _G, spec: GraphSpec) -> str:
    fe = spec.frontend
    g.init("fe.axes2", np.array([2], dtype=np.int64))
    g.init("fe.axes1", np.array([1], dtype=np.int64))
    g.init("fe.idx0", np.array(0, dtype=np.int64))
    x = "input"
    if fe.kind == "birdnet_v24":
        # per-segment min/max normalisation to [-1, 1]  
        g.init("fe.eps", np.array(1e-6, dtype=np.float32))
        g.init("fe.half", np.array(0.5, dtype=np.float32))
        g.init("fe.two", np.array(2.0, dtype=np.float32))
        mn = g.node("ReduceMin", [x], ["fe.min"], axes=[1], keepdims=1)
        x1 = g.node("Sub", [x, mn], ["fe.x1"])
        mx = g.node("ReduceMax", [x1], ["fe.max"], axes=[1], keepdims=1)
        den = g.node("Add", [mx, "fe.eps"], ["fe.den"])
        x2 = g.node("Div", [x1, den], ["fe.x2"])
        x3 = g.node("Sub", [x2, "fe.half"], ["fe.x3"])
        x = g.node("Mul", [x3, "fe.two"], ["fe.x4"])
        sig = g.node("Unsqueeze", [x, "fe.axes2"], ["fe.signal"])
        g.init("fe.flip_starts", np.array([-1], dtype=np.int64))
        g.init("fe.flip_ends", np.array([-(2 ** 63) + 1], dtype=np.int64))
        g.init("fe.flip_steps", np.array([-1], dtype=np.int64))
        branches = []
        for i, s in enumerate(fe.specs):
            p = f"fe.spec{i}"
            g.init(f"{p}.frame_step", np.array(s.hop, dtype=np.int64))
            g.init(f"{p}.frame_length", np.array(s.n_fft, dtype=np.int64))
            g.init(f"{p}.window", hann_window(s.n_fft))
            g.init(f"{p}.mel", mel_matrix(s.n_mels, s.n_bins, fe.sample_rate, s.fmin, s.fmax))
            g.init(f"{p}.exponent", np.array(mag_exponent(s.mag_scale), dtype=np.float32))
            st = g.node("STFT", [sig, f"{p}.frame_step", f"{p}.window", f"{p}.frame_length"],
                        [f"{p}.stft"], onesided=1)
            re = g.node("Gather", [st, "fe.idx0"], [f"{p}.real"], axis=3)
            mel = g.node("MatMul", [re, f"{p}.mel"], [f"{p}.melspec"])
            sq = g.node("Mul", [mel, mel], [f"{p}.power"])
            pw = g.node("Pow", [sq, f"{p}.exponent"], [f"{p}.compressed"])
            fl = g.node("Slice", [pw, "fe.flip_starts", "fe.flip_ends", "fe.axes2",
                                  "fe.flip_steps"], [f"{p}.flipped"])
            tr = g.node("Transpose", [fl], [f"{p}.t"], perm=[0, 2, 1])
            branches.append(g.node("Unsqueeze", [tr, "fe.axes1"], [f"{p}.nchw"]))
        return g.node("Concat", branches, ["spec"], axis=1)
    if fe.kind == "logmel":
        s = fe.specs[0]
        p = "fe.spec0"
        if fe.pad_end:
            g.init("fe.pads", np.array([0, 0, 0, fe.pad_end], dtype=np.int64))
            x = g.node("Pad", [x, "fe.pads"], ["fe.padded"], mode="constant")
        sig = g.node("Unsqueeze", [x, "fe.axes2"], ["fe.signal"])
        g.init(f"{p}.frame_step", np.array(s.hop, dtype=np.int64))
        g.init(f"{p}.frame_length", np.array(s.n_fft, dtype=np.int64))
        g.init(f"{p}.window", hann_window(s.n_fft))
        g.init(f"{p}.mel", mel_matrix(s.n_mels, s.n_bins, fe.sample_rate, s.fmin, s.fmax))
        g.init("fe.idx1", np.array(1, dtype=np.int64))
        g.init("fe.log_floor", np.array(fe.log_floor, dtype=np.float32))
        g.init("fe.log_scale", np.array(fe.log_scale, dtype=np.float32))
        st = g.node("STFT", [sig, f"{p}.frame_step", f"{p}.window", f"{p}.frame_length"],
                    [f"{p}.stft"], onesided=1)
        re = g.node("Gather", [st, "fe.idx0"], [f"{p}.real"], axis=3)
        im = g.node("Gather", [st, "fe.idx1"], [f"{p}.imag"], axis=3)
        re2 = g.node("Mul", [re, re], [f"{p}.re2"])
        im2 = g.node("Mul", [im, im], [f"{p}.im2"])
        pw = g.node("Add", [re2, im2], [f"{p}.power"])
        mag = g.node("Sqrt", [pw], [f"{p}.magnitude"])
        mel = g.node("MatMul", [mag, f"{p}.mel"], [f"{p}.melspec"])
        fl = g.node("Add", [mel, "fe.log_floor"], [f"{p}.floored"])
        lg = g.node("Log", [fl], [f"{p}.log"])
        spectro = "spectrogram" if spec.family == "perch_v2" else f"{p}.logmel"
        lm = g.node("Mul", [lg, "fe.log_scale"], [spectro])
        return g.node("Unsqueeze", [lm, "fe.axes1"], ["spec"])
    raise ValueError(fe.kind)


def build_model_bytes(spec: GraphSpec, weights: Dict[str, np.ndarray] = None) -> bytes:
    weights = make_weights(spec) if weights is None else weights
    g = _G()
    _emit_frontend(g, spec)
    for op in spec.ops:
        k = op["op"]
        if k == "conv":
            n = op["name"]
            g.init(f"{n}.weight", weights[f"{n}.weight"])
            g.init(f"{n}.bias", weights[f"{n}.bias"])
            act = op["act"]
            conv_out = op["out"] if act == "none" else f"{n}.conv"
            g.node("Conv", [op["in"], f"{n}.weight", f"{n}.bias"], [conv_out],
                   dilations=[1, 1], group=op["groups"], kernel_shape=[op["k"], op["k"]],
                   pads=[op["pad"]] * 4, strides=[op["stride"], op["stride"]])
            if act == "silu":
                g.node("Sigmoid", [conv_out], [f"{n}.sig"])
                g.node("Mul", [conv_out, f"{n}.sig"], [op["out"]])
            elif act == "sigmoid":
                g.node("Sigmoid", [conv_out], [op["out"]])
        elif k == "add":
            g.node("Add", [op["a"], op["b"]], [op["out"]])
        elif k == "mul":
            g.node("Mul", [op["a"], op["b"]], [op["out"]])
        elif k == "gap":
            g.node("GlobalAveragePool", [op["in"]], [op["out"]])
        elif k == "flatten":
            g.node("Flatten", [op["in"]], [op["out"]], axis=1)
        elif k == "to_nhwc":
            g.node("Transpose", [op["in"]], [op["out"]], perm=[0, 2, 3, 1])
        elif k == "gemm":
            n = op["name"]
            g.init(f"{n}.weight", weights[f"{n}.weight"])
            g.init(f"{n}.bias", weights[f"{n}.bias"])
            g.node("Gemm", [op["in"], f"{n}.weight", f"{n}.bias"], [op["out"]],
                   alpha=1.0, beta=1.0, transB=1)
        else:
            raise ValueError(k)

    graph = b"".join(_f_bytes(1, n) for n in g.nodes)
    graph += _f_str(2, f"{spec.family}_like_seed{spec.seed}")
    graph += b"".join(_f_bytes(5, t) for t in g.inits)
    graph += _f_bytes(11, _value_info("input", ["batch", spec.frontend.sample_count]))
    for o in spec.outputs:
        graph += _f_bytes(12, _value_info(o["name"], o["shape"]))
    model = _f_varint(1, 8)                                  # ir_version 8
    model += _f_str(2, "birdnet_b200.modelgen") + _f_str(3, "1")
    model += _f_bytes(7, graph)
    model += _f_bytes(8, _f_str(1, "") + _f_varint(2, 17))   # default domain, opset 17
    return model


def write_model(spec: GraphSpec, path: str, weights: Dict[str, np.ndarray] = None) -> None:
    data = build_model_bytes(spec, weights)
    import os
    tmp = f"{path}.{os.getpid()}.tmp"          # several ranks may generate the same (deterministic) file at once
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, path)                      # atomic: readers see either no file or the complete one


# ---------------------------------------------------------------- parsing (tests / tools)
def _read_varint(b: bytes, i: int) -> Tuple[int, int]:
    v = 0
    shift = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << shift
        if not c & 0x80:
            return v, i
        shift += 7


def _fields(b: bytes):
    i = 0
    n = len(b)
    while i < n:
        key, i = _read_varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(b, i)
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        elif wt == 2:
            ln, i = _read_varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        else:
            raise ValueError(f"wire type {wt}")
        yield f, wt, v


def _s64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_tensor(b: bytes):
    dims, dt, name, raw = [], None, "", b""
    for f, wt, v in _fields(b):
        if f == 1:
            dims.append(_s64(v))
        elif f == 2:
            dt = v
        elif f == 8:
            name = v.decode()
        elif f == 9:
            raw = v
    dtype = {FLOAT: np.float32, INT64: np.int64}[dt]
    return name, np.frombuffer(raw, dtype=dtype).reshape(dims)


def parse_model(data: bytes) -> dict:
    """Returns {"nodes":[{op,name,inputs,outputs,attrs}], "initializers":{name:array},
    "inputs":[(name,shape)], "outputs":[(name,shape)], "opset":int}."""
    graph = None
    opset = None
    for f, wt, v in _fields(data):
        if f == 7:
            graph = v
        elif f == 8:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    opset = v2
    out = dict(nodes=[], initializers={}, inputs=[], outputs=[], opset=opset)

    def vinfo(b):
        name, shape = "", []
        for f, _, v in _fields(b):
            if f == 1:
                name = v.decode()
            elif f == 2:
                for f2, _, v2 in _fields(v):          # TypeProto
                    if f2 == 1:
                        for f3, _, v3 in _fields(v2):  # Tensor
                            if f3 == 2:
                                for f4, _, v4 in _fields(v3):  # shape.dim
                                    d = -1
                                    for f5, _, v5 in _fields(v4):
                                        if f5 == 1:
                                            d = _s64(v5)
                                    shape.append(d)
        return name, shape

    for f, wt, v in _fields(graph):
        if f == 1:
            nd = dict(op="", name="", inputs=[], outputs=[], attrs={})
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    nd["inputs"].append(v2.decode())
                elif f2 == 2:
                    nd["outputs"].append(v2.decode())
                elif f2 == 3:
                    nd["name"] = v2.decode()
                elif f2 == 4:
                    nd["op"] = v2.decode()
                elif f2 == 5:
                    an, av, ints = "", None, []
                    for f3, wt3, v3 in _fields(v2):
                        if f3 == 1:
                            an = v3.decode()
                        elif f3 == 2:
                            av = struct.unpack("<f", v3)[0]
                        elif f3 == 3:
                            av = _s64(v3)
                        elif f3 == 4:
                            av = v3.decode()
                        elif f3 == 8:
                            if wt3 == 2:      # packed
                                j = 0
                                while j < len(v3):
                                    x, j = _read_varint(v3, j)
                                    ints.append(_s64(x))
                            else:
                                ints.append(_s64(v3))
                    nd["attrs"][an] = ints if ints else av
            out["nodes"].append(nd)
        elif f == 5:
            name, arr = parse_tensor(v)
            out["initializers"][name] = arr
        elif f == 11:
            out["inputs"].append(vinfo(v))
        elif f == 12:
            out["outputs"].append(vinfo(v))
    return out
