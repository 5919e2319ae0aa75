"""The C++ ONNX reader and planner against a model file written by an INDEPENDENT producer: the build-authored graph as
a torch.nn.Module exported with torch's own legacy ONNX exporter (tools/export_torch_onnx.py), instead of the in-house
protobuf writer.  Both files must be matched into the same layer plan (same fused ops, same shapes, same weight bits).
The reference loads arbitrary files through ONNX Runtime (src/classifier.rs:340-357); this pins the engine's dialect on
something it did not write itself.  CPU only (bn_model_plan_summary never touches a GPU)."""
import ctypes as C
import importlib.util
import os

import pytest

from birdnet_b200 import _ffi
from birdnet_b200.modelgen import get_spec, write_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _summary(path: str) -> str:
    need = C.c_uint64()
    st = _ffi.lib.bn_model_plan_summary(path.encode(), -1, None, 0, C.byref(need))
    assert st == 0, _ffi.last_error()
    buf = C.create_string_buffer(int(need.value))
    assert _ffi.lib.bn_model_plan_summary(path.encode(), -1, buf, need.value, None) == 0
    return buf.value.decode()


@pytest.fixture(scope="module")
def exporter():
    spec = importlib.util.spec_from_file_location("export_torch_onnx", os.path.join(ROOT, "tools", "export_torch_onnx.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("family,kw", [("birdnet_v24", dict(num_species=64)), ("birdnet_v30", dict(num_species=48)),
                                       ("perch_v2", dict(num_species=40))])
def test_torch_exported_file_gives_the_same_plan(exporter, tmp_path, family, kw):
    ours = os.path.join(str(tmp_path), "inhouse.onnx")
    theirs = os.path.join(str(tmp_path), "torch.onnx")
    write_model(get_spec(family, **kw), ours)
    exporter.export(family, theirs, **kw)
    a, b = _summary(ours), _summary(theirs)
    assert a.count("\n") > 20
    la, lb = a.splitlines(), b.splitlines()
    assert len(la) == len(lb), (len(la), len(lb))
    for x, y in zip(la, lb):
        assert x == y, (x, y)
