import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rust-birdnet-onnx_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from birdnet_b200 import _ffi
        return _ffi.lib.bn_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()


@pytest.fixture(scope="session")
def v24_model_path():
    from birdnet_b200.modelgen.make_models import ensure_model
    return ensure_model("birdnet_v24")


@pytest.fixture(scope="session")
def v24_spec():
    from birdnet_b200.modelgen import get_spec
    return get_spec("birdnet_v24")
