"""ORACLE (test infrastructure, not product): ctypes loader for postprocess_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build() -> str:
    src = os.path.join(_HERE, "postprocess_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Pred(C.Structure):
    _fields_ = [("index", C.c_uint32), ("confidence", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_sigmoid.restype = C.c_float
        _lib.oracle_sigmoid.argtypes = [C.c_float]
        _lib.oracle_top_k.restype = C.c_uint64
        _lib.oracle_top_k.argtypes = [C.POINTER(C.c_float), C.c_uint64, C.c_uint64, C.c_int, C.c_float,
                                      C.POINTER(_Pred)]
        _lib.oracle_top_k_batch.restype = None
        _lib.oracle_top_k_batch.argtypes = [C.POINTER(C.c_float), C.c_uint64, C.c_uint64, C.c_uint64,
                                            C.c_int, C.c_float, C.POINTER(_Pred), C.POINTER(C.c_uint32)]
        _lib.oracle_filter.restype = C.c_uint64
        _lib.oracle_filter.argtypes = [C.POINTER(_Pred), C.c_uint64, C.POINTER(C.c_uint8),
                                       C.POINTER(C.c_float), C.c_float, C.c_int, C.POINTER(_Pred)]
        _lib.oracle_calculate_week.restype = C.c_float
        _lib.oracle_calculate_week.argtypes = [C.c_uint32, C.c_uint32]
        _lib.oracle_random_logits.restype = None
        _lib.oracle_random_logits.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_float)]
        _lib.oracle_mock_embeddings.restype = None
        _lib.oracle_mock_embeddings.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_float)]
    return _lib


def sigmoid(x: float) -> float:
    return float(lib().oracle_sigmoid(C.c_float(x)))


def top_k_predictions(logits, top_k: int, min_confidence: Optional[float]) -> List[Tuple[int, float]]:
    """[(index, confidence)] in the reference's output order (postprocess.rs:40-87)."""
    a = np.ascontiguousarray(logits, dtype=np.float32)
    n = a.shape[0]
    k = min(max(int(top_k), 0), n)
    out = (_Pred * max(k, 1))()
    m = lib().oracle_top_k(a.ctypes.data_as(C.POINTER(C.c_float)), n, min(int(top_k), 2 ** 64 - 1),
                           0 if min_confidence is None else 1,
                           0.0 if min_confidence is None else float(min_confidence), out)
    return [(int(out[i].index), float(out[i].confidence)) for i in range(m)]


def top_k_batch(logits: np.ndarray, top_k: int, min_confidence: Optional[float]):
    """(idx [B,k] uint32, conf [B,k] f32, counts [B])."""
    a = np.ascontiguousarray(logits, dtype=np.float32)
    rows, n = a.shape
    k = min(max(int(top_k), 0), n)
    out = np.zeros((rows, max(k, 1), 2), dtype=np.uint32)
    counts = np.zeros(rows, dtype=np.uint32)
    if k:
        lib().oracle_top_k_batch(a.ctypes.data_as(C.POINTER(C.c_float)), rows, n, int(top_k),
                                 0 if min_confidence is None else 1,
                                 0.0 if min_confidence is None else float(min_confidence),
                                 out.ctypes.data_as(C.POINTER(_Pred)), counts.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out[:, :k, 0].copy(), out[:, :k, 1].copy().view(np.float32), counts


def filter_predictions(preds: Sequence[Tuple[str, float, int]], location_scores: Sequence[Tuple[str, float]],
                       threshold: float, rerank: bool) -> List[Tuple[str, float, int]]:
    """filter_predictions_impl (rangefilter.rs:333-386) on (species, confidence, index) tuples
    and (species, score) location scores; map built with last-duplicate-wins like HashMap."""
    m = {}
    for sp, sc in location_scores:
        m[sp] = sc
    n = len(preds)
    if n == 0:
        return []
    pin = (_Pred * n)()
    in_map = (C.c_uint8 * n)()
    score = (C.c_float * n)()
    for j, (sp, conf, idx) in enumerate(preds):
        pin[j].index = j
        pin[j].confidence = conf
        if sp in m:
            in_map[j] = 1
            score[j] = m[sp]
    out = (_Pred * n)()
    cnt = lib().oracle_filter(pin, n, in_map, score, C.c_float(threshold), 1 if rerank else 0, out)
    return [(preds[out[i].index][0], float(out[i].confidence), preds[out[i].index][2]) for i in range(cnt)]


def calculate_week(month: int, day: int) -> float:
    return float(lib().oracle_calculate_week(month, day))


def random_logits(count: int, seed: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.float32)
    lib().oracle_random_logits(count, seed & (2 ** 64 - 1), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def mock_embeddings(dim: int, seed: int) -> np.ndarray:
    out = np.zeros(dim, dtype=np.float32)
    lib().oracle_mock_embeddings(dim, seed & (2 ** 64 - 1), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out
