"""Error variants of the reference with their exact Display strings (src/error.rs:6-128)."""
from __future__ import annotations

from . import _ffi


def _rust_f32(v: float) -> str:
    """Rust `{}` for f32: shortest round-trip, integral values print without '.0'... with '.0'?
    Rust prints 91.0f32 as "91" for `{}`: Display for floats omits the fraction only if it has
    none -> "91".  (std::fmt float Display: 1.0 -> "1", 1.5 -> "1.5".)"""
    import numpy as np
    f = float(np.float32(v))
    if f != f:
        return "NaN"
    if f in (float("inf"), float("-inf")):
        return "inf" if f > 0 else "-inf"
    if f == int(f) and abs(f) < 1e16:
        return str(int(f))
    return np.format_float_positional(np.float32(v), unique=True, trim="-")


def _rust_duration(seconds: float) -> str:
    """Rust `{:?}` of std::time::Duration (e.g. 30s, 1.5s, 100ms, 10µs, 5ns)."""
    ns = int(round(seconds * 1e9))
    if ns >= 1_000_000_000:
        whole, frac, unit = ns // 1_000_000_000, ns % 1_000_000_000, "s"
        digits = 9
    elif ns >= 1_000_000:
        whole, frac, unit = ns // 1_000_000, ns % 1_000_000, "ms"
        digits = 6
    elif ns >= 1_000:
        whole, frac, unit = ns // 1_000, ns % 1_000, "µs"
        digits = 3
    else:
        return f"{ns}ns"
    if frac == 0:
        return f"{whole}{unit}"
    return f"{whole}.{str(frac).zfill(digits).rstrip('0')}{unit}"


class Error(Exception):
    """Base of all classifier errors (reference: `birdnet_onnx::Error`)."""


class InputSize(Error):
    def __init__(self, expected: int, got: int):
        self.expected, self.got = expected, got
        super().__init__(f"input size mismatch: expected {expected} samples, got {got}")


class BatchInputSize(Error):
    def __init__(self, index: int, expected: int, got: int):
        self.index, self.expected, self.got = index, expected, got
        super().__init__(
            f"batch input size mismatch: segment {index} has {got} samples, expected {expected}")


class ModelDetection(Error):
    def __init__(self, reason: str):
        self.reason = reason
        super().__init__(f"model detection failed: {reason}")


class LabelCount(Error):
    def __init__(self, expected: int, got: int):
        self.expected, self.got = expected, got
        super().__init__(f"label count mismatch: model expects {expected}, got {got}")


class ModelPathRequired(Error):
    def __init__(self):
        super().__init__("model path required")


class LabelsRequired(Error):
    def __init__(self):
        super().__init__("labels required (provide path or vec)")


class ModelLoad(Error):
    def __init__(self, msg: str):
        super().__init__(f"failed to load model: {msg}")


class LabelLoad(Error):
    def __init__(self, path: str, reason: str):
        self.path, self.reason = path, reason
        super().__init__(f"failed to load labels from {path}: {reason}")


class LabelParse(Error):
    def __init__(self, msg: str):
        super().__init__(f"failed to parse labels: {msg}")


class Inference(Error):
    def __init__(self, msg: str):
        self.message = msg
        super().__init__(f"inference failed: {msg}")


class InvalidCoordinates(Error):
    def __init__(self, latitude: float, longitude: float, reason: str):
        self.latitude, self.longitude, self.reason = latitude, longitude, reason
        super().__init__(f"invalid coordinates: latitude: {_rust_f32(latitude)}, "
                         f"longitude: {_rust_f32(longitude)}, reason: {reason}")


class InvalidDate(Error):
    def __init__(self, month: int, day: int, reason: str):
        self.month, self.day, self.reason = month, day, reason
        super().__init__(f"invalid date: month: {month}, day: {day}, reason: {reason}")


class RangeFilterInference(Error):
    def __init__(self, msg: str):
        super().__init__(f"range filter inference failed: {msg}")


class Timeout(Error):
    def __init__(self, duration: float):
        self.duration = duration           # seconds
        super().__init__(f"inference timed out after {_rust_duration(duration)}")


class Cancelled(Error):
    def __init__(self):
        super().__init__("inference was cancelled")


class RuntimeInit(Error):
    """Reference text says "ONNX Runtime" (error.rs:110); here the runtime is the CUDA device."""
    def __init__(self, msg: str):
        super().__init__(f"failed to initialize ONNX Runtime: {msg}")


class AudioFormat(Error):
    def __init__(self, reason: str):
        super().__init__(f"unsupported audio format: {reason}")


class AudioRead(Error):
    def __init__(self, path: str, reason: str):
        super().__init__(f"failed to read audio file {path}: {reason}")


def raise_for_status(status: int, timeout_s: float = None):
    """Translate a bn_status + thread-local message into the reference's Error variant."""
    if status == _ffi.BN_OK:
        return
    msg = _ffi.last_error()
    a, b, c = _ffi.last_error_detail()
    if status == _ffi.BN_ERR_INPUT_SIZE:
        raise InputSize(b, c)
    if status == _ffi.BN_ERR_BATCH_INPUT_SIZE:
        raise BatchInputSize(a, b, c)
    if status == _ffi.BN_ERR_MODEL_DETECTION:
        raise ModelDetection(msg)
    if status == _ffi.BN_ERR_MODEL_PATH_REQUIRED:
        raise ModelPathRequired()
    if status == _ffi.BN_ERR_MODEL_LOAD:
        raise ModelLoad(msg)
    if status == _ffi.BN_ERR_TIMEOUT:
        raise Timeout(timeout_s if timeout_s is not None else a / 1e9)
    if status == _ffi.BN_ERR_CANCELLED:
        raise Cancelled()
    if status == _ffi.BN_ERR_RUNTIME_INIT:
        raise RuntimeInit(msg)
    if status == _ffi.BN_ERR_RANGE_FILTER_INFERENCE:
        raise RangeFilterInference(msg)
    raise Inference(msg)
