"""The C-ABI library loads and exports every symbol include/birdnet_b200.h declares."""
import ctypes as C
import os
import re

from birdnet_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "birdnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bn_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(_ffi.lib, n)]
    assert not missing, missing


def test_binding_covers_header():
    assert sorted(_ffi.SIGNATURES) == _declared()


def test_rust_binding_lists_every_symbol():
    """bindings/rust/birdnet_b200_sys.rs (source-only: no Rust toolchain in the image) declares exactly the header's symbols."""
    src = open(os.path.join(ROOT, "bindings", "rust", "birdnet_b200_sys.rs")).read()
    src = re.sub(r"//.*", "", src)
    assert sorted(set(re.findall(r"\bpub fn (bn_[a-z0-9_]+)\s*\(", src))) == _declared()


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.Pred) == 8
    assert C.sizeof(_ffi.DeviceCfg) == 16
    assert C.sizeof(_ffi.TensorInfo) == 64 + 8 + 64
    assert C.sizeof(_ffi.Outputs) == 64
    assert C.sizeof(_ffi.RunOpts) == 24


def test_version_and_error_channel():
    assert b"sm_100a" in _ffi.lib.bn_version()
    info = _ffi.IoInfo()
    st = _ffi.lib.bn_model_inspect(b"/nonexistent.onnx", -1, C.byref(info))
    assert st == _ffi.BN_ERR_MODEL_LOAD and "cannot open" in _ffi.last_error()
    assert _ffi.lib.bn_model_inspect(None, -1, C.byref(info)) == _ffi.BN_ERR_MODEL_PATH_REQUIRED


def test_no_cpu_fallback_without_gpu(has_gpu, v24_model_path):
    """Compute entry points must fail loudly (RuntimeInit), never compute on the host."""
    if has_gpu:
        return
    import birdnet_b200 as bb
    import pytest
    with pytest.raises(bb.RuntimeInit):
        bb.Classifier.builder().model_path(v24_model_path).labels(["x"] * 6522).build()
