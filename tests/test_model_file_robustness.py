"""A malformed / hostile model file must come back as BN_ERR_MODEL_LOAD (the reference reports such files
through ort as Error::ModelLoad, src/classifier.rs:340-357), never as a crash or an out-of-bounds read.
CPU only: bn_model_inspect parses and plans the file without touching a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from birdnet_b200 import _ffi
from birdnet_b200.modelgen import build_model_bytes, get_spec, make_weights
from birdnet_b200.modelgen import onnx_writer as ow


@pytest.fixture(scope="module")
def small_model_bytes():
    # the v2.4 graph with a small head keeps every case below fast
    spec = get_spec("birdnet_v24", num_species=64)
    return spec, build_model_bytes(spec)


def _inspect(tmp_path, data: bytes, name="m.onnx"):
    p = os.path.join(str(tmp_path), name)
    with open(p, "wb") as f:
        f.write(data)
    info = _ffi.IoInfo()
    st = _ffi.lib.bn_model_inspect(p.encode(), -1, C.byref(info))
    return st, _ffi.last_error()


def test_valid_file_loads(small_model_bytes, tmp_path):
    st, msg = _inspect(tmp_path, small_model_bytes[1])
    assert st == _ffi.BN_OK, msg


def test_truncations_are_model_load_errors(small_model_bytes, tmp_path):
    data = small_model_bytes[1]
    n = len(data)
    cuts = sorted({0, 1, 2, 7, 64, 1000, n // 7, n // 3, n // 2, n - 4097, n - 100, n - 5, n - 1})
    for cut in cuts:
        st, msg = _inspect(tmp_path, data[:cut])
        assert st in (_ffi.BN_ERR_MODEL_LOAD, _ffi.BN_ERR_MODEL_DETECTION), (cut, st, msg)
        assert msg


def test_random_corruption_never_crashes(small_model_bytes, tmp_path):
    data = bytearray(small_model_bytes[1])
    rng = np.random.default_rng(0)
    # the graph structure lives in the last ~100 KB (nodes, small initializers); the head of the file is weights
    for trial in range(150):
        d = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            lo = len(d) - 120_000 if rng.random() < 0.8 else 0
            pos = int(rng.integers(max(lo, 0), len(d)))
            d[pos] = int(rng.integers(0, 256))
        st, msg = _inspect(tmp_path, bytes(d))
        # a flipped weight byte can leave a perfectly valid model: OK is a legal answer, a crash is not
        assert st in (_ffi.BN_OK, _ffi.BN_ERR_MODEL_LOAD, _ffi.BN_ERR_MODEL_DETECTION), (trial, st, msg)


def test_wrong_wire_types_are_rejected(tmp_path):
    # a float attribute encoded as a varint (null dereference in f32_of), a string as a varint
    bad_attr = ow._f_str(1, "alpha") + ow._f_varint(2, 7) + ow._f_varint(20, 1)
    node = ow._f_str(1, "x") + ow._f_str(2, "y") + ow._f_str(4, "Gemm") + ow._f_bytes(5, bad_attr)
    graph = ow._f_bytes(1, node)
    model = ow._f_varint(1, 8) + ow._f_bytes(7, graph)
    st, msg = _inspect(tmp_path, model)
    assert st == _ffi.BN_ERR_MODEL_LOAD and "wire type" in msg, msg
    graph = ow._f_varint(2, 5)                    # graph name as a varint
    st, msg = _inspect(tmp_path, ow._f_bytes(7, graph))
    assert st == _ffi.BN_ERR_MODEL_LOAD and "wire type" in msg, msg


def test_initializer_payload_must_match_its_shape(tmp_path):
    def tensor(dims, raw=None, floats=None, dt=1):
        b = b"".join(ow._f_varint(1, d & ((1 << 64) - 1)) for d in dims) + ow._f_varint(2, dt) + ow._f_str(8, "w")
        if raw is not None:
            b += ow._f_bytes(9, raw)
        if floats is not None:
            b += ow._f_bytes(4, np.asarray(floats, np.float32).tobytes())
        return ow._f_bytes(7, ow._f_bytes(5, b))
    for blob, needle in [
        (tensor([4, 4], raw=b"\0" * 60), "size mismatch"),
        (tensor([4, 4], floats=[1.0] * 3), "holds 3 values"),
        (tensor([4, 4]), "holds 0 values"),
        (tensor([-1, 4], raw=b""), "negative dimension"),
        (tensor([1 << 40, 1 << 40], raw=b""), "too large"),
        (tensor([2], raw=b"\0" * 8, dt=11), "unsupported data_type"),
    ]:
        st, msg = _inspect(tmp_path, blob)
        assert st == _ffi.BN_ERR_MODEL_LOAD and needle in msg, (needle, msg)


def test_weights_outside_the_fp16_range_are_refused(small_model_bytes, tmp_path):
    """hi/lo fp16 operand format: |w| > 65504 cannot be represented -> ModelLoad."""
    spec = small_model_bytes[0]
    w = make_weights(spec)
    w["s3b0.project.weight"] = w["s3b0.project.weight"].copy()
    w["s3b0.project.weight"][3, 5, 0, 0] = 7.0e4
    st, msg = _inspect(tmp_path, build_model_bytes(spec, w))
    assert st == _ffi.BN_ERR_MODEL_LOAD and "fp16 range" in msg and "s3b0.project" in msg, msg
    w["s3b0.project.weight"][3, 5, 0, 0] = np.float32("nan")
    st, msg = _inspect(tmp_path, build_model_bytes(spec, w))
    assert st == _ffi.BN_ERR_MODEL_LOAD and "fp16 range" in msg, msg


def test_directory_and_missing_path(tmp_path):
    info = _ffi.IoInfo()
    assert _ffi.lib.bn_model_inspect(str(tmp_path).encode(), -1, C.byref(info)) == _ffi.BN_ERR_MODEL_LOAD
    assert _ffi.lib.bn_model_inspect(os.path.join(str(tmp_path), "nope.onnx").encode(), -1, C.byref(info)) == _ffi.BN_ERR_MODEL_LOAD
    assert _ffi.lib.bn_model_inspect(None, -1, C.byref(info)) == _ffi.BN_ERR_MODEL_PATH_REQUIRED
