"""Multi-GPU sharding: segments are independent, so whole batches are handed to
the devices (in-process pool: a shared queue feeding `depth` contexts per device; one process per GPU:
contiguous blocks, `shard_range`) and the host gathers results in caller order.  No collective is
involved; the reference's only hook is CUDAConfig::with_device_id (src/cuda_config.rs:179-182),
i.e. "one Classifier per GPU, the user shards".
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .errors import raise_for_status


def shard_range(n_segments: int, rank: int, world: int, batch: int) -> Tuple[int, int]:
    """[lo, hi) of the segment index range owned by `rank`: whole batches, block-partitioned (the
    one-process-per-GPU partition of bench.py; the in-process pool balances through a shared queue)."""
    nb = (n_segments + batch - 1) // batch
    b0, b1 = rank * nb // world, (rank + 1) * nb // world
    return min(n_segments, b0 * batch), min(n_segments, b1 * batch)


class DevicePool:
    """In-process dispatcher over several GPUs of one box (bn_pool_*)."""

    def __init__(self, model_path: str, device_ids: Sequence[int], ctx_batch: int = 256,
                 model_type_override: int = -1, depth: int = 3):
        self._h = C.c_void_p()
        ids = (C.c_int32 * len(device_ids))(*device_ids)
        raise_for_status(_ffi.lib.bn_pool_create_ex(model_path.encode(), ids, len(device_ids),
                                                    model_type_override, ctx_batch, depth, C.byref(self._h)))
        info = _ffi.IoInfo()
        raise_for_status(_ffi.lib.bn_model_inspect(model_path.encode(), model_type_override, C.byref(info)))
        self.num_species = int(info.num_species)
        self.embedding_dim = int(info.embedding_dim)
        self.sample_count = int(info.sample_count)
        self._top_k = 10

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _ffi.lib.bn_pool_destroy(h)

    def set_postprocess(self, top_k: int, min_confidence: Optional[float]) -> None:
        self._top_k = top_k
        raise_for_status(_ffi.lib.bn_pool_set_postprocess(
            self._h, top_k, 0 if min_confidence is None else 1, min_confidence or 0.0))

    def set_range_filter(self, state: np.ndarray, score: np.ndarray, rerank: bool) -> None:
        raise_for_status(_ffi.lib.bn_pool_set_range_filter(
            self._h, state.ctypes.data_as(C.POINTER(C.c_uint8)),
            score.ctypes.data_as(C.POINTER(C.c_float)), len(state), 1 if rerank else 0))

    def run(self, segments: Sequence[np.ndarray], options=None):
        """-> (logits [n,N], embeddings [n,E] | None, topk idx [n,k], conf [n,k], counts [n]).
        `options`: InferenceOptions (timeout / cancellation), applied to every batch of the call."""
        from .classifier import _segment_arrays, _run_opts
        n = len(segments)
        ptrs, lens, keep = _segment_arrays(segments)
        k = min(self._top_k, self.num_species)
        logits = np.empty((n, self.num_species), dtype=np.float32)
        emb = np.empty((n, self.embedding_dim), dtype=np.float32) if self.embedding_dim else None
        topk = np.zeros((n, max(k, 1), 2), dtype=np.uint32)
        counts = np.zeros(n, dtype=np.uint32)
        ro, timeout = _run_opts(options)
        st = _ffi.lib.bn_pool_run(
            self._h, ptrs, lens, n, C.byref(ro) if ro is not None else None, logits.ctypes.data_as(C.POINTER(C.c_float)),
            emb.ctypes.data_as(C.POINTER(C.c_float)) if emb is not None else None,
            topk.ctypes.data_as(C.POINTER(_ffi.Pred)), counts.ctypes.data_as(C.POINTER(C.c_uint32)), max(k, 1))
        raise_for_status(st, timeout)
        return logits, emb, topk[:, :k, 0].copy(), topk[:, :k, 1].copy().view(np.float32), counts
