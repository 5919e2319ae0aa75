"""Classifier / ClassifierBuilder / BatchInferenceContext — the reference's public API
(src/classifier.rs, src/batch_context.rs) over the C ABI of include/birdnet_b200.h.

Everything numeric happens on the GPU (front-end, CNN, top-k/sigmoid/threshold/range mask);
this layer validates, marshals pointers and turns (index, confidence) pairs into `Prediction`s
with label strings — what classifier.rs:872-1058 does after `session.run`.
"""
from __future__ import annotations

import ctypes as C
import sys
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .errors import (Inference, InputSize, LabelCount, LabelsRequired, ModelPathRequired,
                     raise_for_status)
from .inference_options import InferenceOptions
from .labels import load_labels_from_file
from .types import (ExecutionProviderInfo, LocationScore, ModelConfig, ModelType, Prediction,
                    PredictionResult)

_lib = _ffi.lib


def _run_opts(options: Optional[InferenceOptions]):
    """bn_run_opts for an InferenceOptions (classifier.rs:504-526)."""
    if options is None or not options.needs_monitor():
        return None, None
    ro = _ffi.RunOpts()
    if options.cancellation_token is not None:
        ro.cancel_flag = C.pointer(options.cancellation_token._flag)
    if options.timeout is not None:
        ro.has_timeout = 1
        ro.timeout_ns = max(0, int(round(options.timeout * 1e9)))
    return ro, options.timeout


_addressof, _c_char = C.addressof, C.c_char


def _segment_arrays(segments: Sequence) -> tuple:
    """Pointer + length arrays for B caller slices; slices stay where they are (no host pack
    here: the engine gathers them straight into pinned staging)."""
    n = len(segments)
    f32 = np.dtype(np.float32)
    keep = [seg if (type(seg) is np.ndarray and seg.dtype == f32 and seg.ndim == 1 and seg.flags.c_contiguous)
            else np.ascontiguousarray(seg, dtype=np.float32).reshape(-1) for seg in segments]
    # address / length tables as two numpy vectors (an order of magnitude cheaper than filling ctypes arrays item by item)
    try:
        fb = _c_char.from_buffer
        addr = np.array([_addressof(fb(a)) for a in keep], dtype=np.uint64)
    except (TypeError, ValueError):                          # read-only or empty slices: the slower attribute
        addr = np.array([a.ctypes.data for a in keep], dtype=np.uint64)
    size = np.array([a.shape[0] for a in keep], dtype=np.uint64)
    keep.append(addr)
    keep.append(size)
    return addr.ctypes.data_as(C.POINTER(C.c_void_p)), size.ctypes.data_as(C.POINTER(C.c_uint64)), keep


class ClassifierBuilder:
    """src/classifier.rs:46-383."""

    def __init__(self):
        self._model_path: Optional[str] = None
        self._labels_path: Optional[str] = None
        self._labels: Optional[List[str]] = None
        self._model_type: Optional[ModelType] = None
        self._requested = ExecutionProviderInfo.Cpu
        self._top_k = 10                       # classifier.rs:72
        self._min_confidence: Optional[float] = None   # classifier.rs:73
        self._device_id = 0
        self._pack_threads = 0

    def model_path(self, path: str) -> "ClassifierBuilder":
        self._model_path = str(path)
        return self

    def labels_path(self, path: str) -> "ClassifierBuilder":
        self._labels_path, self._labels = str(path), None
        return self

    def labels(self, labels: List[str]) -> "ClassifierBuilder":
        self._labels, self._labels_path = list(labels), None
        return self

    def model_type(self, model_type: ModelType) -> "ClassifierBuilder":
        self._model_type = model_type
        return self

    def execution_provider(self, provider=None) -> "ClassifierBuilder":
        """Accepted for source compatibility; there is exactly one provider here (the B200
        engine).  Like the reference (classifier.rs:107-116) it does not change
        requested_provider()."""
        return self

    def with_cuda(self) -> "ClassifierBuilder":
        if self._requested is ExecutionProviderInfo.Cpu:       # classifier.rs:28-31
            self._requested = ExecutionProviderInfo.Cuda
        return self

    def with_b200(self, device_id: int = 0) -> "ClassifierBuilder":
        if self._requested is ExecutionProviderInfo.Cpu:
            self._requested = ExecutionProviderInfo.B200
        self._device_id = device_id
        return self

    def device_id(self, device_id: int) -> "ClassifierBuilder":   # CUDAConfig::with_device_id
        self._device_id = device_id
        return self

    def pack_threads(self, n: int) -> "ClassifierBuilder":
        self._pack_threads = n
        return self

    def top_k(self, k: int) -> "ClassifierBuilder":
        self._top_k = int(k)
        return self

    def min_confidence(self, threshold: float) -> "ClassifierBuilder":
        self._min_confidence = float(threshold)
        return self

    def build(self) -> "Classifier":
        if self._model_path is None:                      # classifier.rs:336
            raise ModelPathRequired()
        if self._labels is None and self._labels_path is None:   # classifier.rs:337
            raise LabelsRequired()
        cfg = _ffi.DeviceCfg(self._device_id,
                             -1 if self._model_type is None else self._model_type.value,
                             self._pack_threads, 0)
        handle = C.c_void_p()
        # model load + detection first (classifier.rs:340-357), labels after (360-371)
        raise_for_status(_lib.bn_engine_create(self._model_path.encode(), C.byref(cfg), C.byref(handle)))
        try:
            info = _ffi.IoInfo()
            raise_for_status(_lib.bn_engine_io_info(handle, C.byref(info)))
            mt = ModelType(info.model_type)
            config = ModelConfig(mt, int(info.sample_rate), float(info.segment_duration),
                                 int(info.sample_count), int(info.num_species),
                                 int(info.embedding_dim) if info.embedding_dim else None)
            labels = (load_labels_from_file(self._labels_path, mt)
                      if self._labels is None else list(self._labels))
            if len(labels) != config.num_species:
                raise LabelCount(config.num_species, len(labels))
            top_k = min(max(self._top_k, 0), 2 ** 64 - 1)
            raise_for_status(_lib.bn_engine_set_postprocess(
                handle, top_k, 0 if self._min_confidence is None else 1,
                0.0 if self._min_confidence is None else self._min_confidence))
        except Exception:
            _lib.bn_engine_destroy(handle)
            raise
        return Classifier(handle, config, labels, self._requested, self._top_k, self._min_confidence,
                          info)


class _PinnedBlock:
    """Owner of one bn_host_alloc block (freed when the last array view dies)."""

    def __init__(self, nbytes: int):
        self.ptr = _lib.bn_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(f"bn_host_alloc({nbytes}) failed")
        self.nbytes = nbytes

    def __del__(self, _free=_lib.bn_host_free):          # bound now: module globals may be gone at interpreter shutdown
        p, self.ptr = getattr(self, "ptr", None), None
        if p:
            _free(p)


def pinned_array(shape, dtype=np.float32) -> np.ndarray:
    """numpy array in page-locked host memory (bn_host_alloc): segments that live here are copied to the GPU in
    place by predict_batch / predict_batch_with_context, without the gather into the engine's staging slab."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    blk = _PinnedBlock(max(n, 1))
    buf = (C.c_char * blk.nbytes).from_address(blk.ptr)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    _PINNED_OWNERS[id(buf)] = blk
    import weakref
    weakref.finalize(buf, _PINNED_OWNERS.pop, id(buf), None)
    return arr


_PINNED_OWNERS = {}


class registered(object):
    """Context manager: page-lock an existing numpy array in place (bn_host_register) for as long as the block runs, so
    that segments sliced from it are copied to the GPU without the gather into the engine's staging slab.

        with bb.registered(recording):                      # any C-contiguous array the caller already owns
            results = clf.predict_batch_with_context(ctx, [recording[i] for i in range(n)])
    """

    def __init__(self, array: np.ndarray):
        if not isinstance(array, np.ndarray) or not array.flags["C_CONTIGUOUS"] or array.nbytes == 0:
            raise ValueError("registered(): a non-empty C-contiguous numpy array is required")
        self._array = array
        self._ptr = array.ctypes.data

    def __enter__(self):
        raise_for_status(_lib.bn_host_register(C.c_void_p(self._ptr), self._array.nbytes))
        return self._array

    def __exit__(self, *exc):
        _lib.bn_host_unregister(C.c_void_p(self._ptr))
        return False


class BatchInferenceContext:
    """src/batch_context.rs:70-339: pinned input slab + device slabs + streams, reused."""

    def __init__(self, classifier: "Classifier", handle, max_batch_size: int):
        self._classifier = classifier       # keeps the engine alive
        self._h = handle
        self._max = max_batch_size

    def __del__(self, _destroy=_lib.bn_ctx_destroy):     # bound now: module globals may be gone at interpreter shutdown
        h, self._h = getattr(self, "_h", None), None
        if h:
            _destroy(h)

    def max_batch_size(self) -> int:
        return int(_lib.bn_ctx_max_batch_size(self._h))

    def sample_count(self) -> int:
        return self._classifier.config().sample_count

    def input_buffer_capacity(self) -> int:
        return self.max_batch_size() * self.sample_count()

    def input_buffer_bytes(self) -> int:
        return int(_lib.bn_ctx_input_buffer_bytes(self._h))

    def model_type(self) -> ModelType:
        return self._classifier.config().model_type

    def last_launch_count(self) -> int:
        return int(_lib.bn_ctx_last_launch_count(self._h))

    def last_run_in_place(self) -> bool:
        """True when the last run copied host -> device straight from the caller's page-locked slices (no gather)."""
        return bool(_lib.bn_ctx_last_run_in_place(self._h))

    def nonfinite_segments(self) -> int:
        """Segments of the last run whose logits were not all finite (see bn_ctx_nonfinite_segments)."""
        return int(_lib.bn_ctx_nonfinite_segments(self._h))

    # ---- introspection used by the parity tests and bench.py ------------------------------
    def read_tensor(self, name: str, batch: int) -> np.ndarray:
        """Intermediate tensor of the last run, NHWC, flattened per segment: [batch, elems]."""
        per = C.c_uint64()
        raise_for_status(_lib.bn_ctx_read_tensor(self._h, name.encode(), None, 0, C.byref(per)))
        out = np.empty((batch, per.value), dtype=np.float32)
        raise_for_status(_lib.bn_ctx_read_tensor(self._h, name.encode(),
                                                 out.ctypes.data_as(C.POINTER(C.c_float)), out.size, None))
        return out

    def read_normalized(self, batch: int) -> np.ndarray:
        out = np.empty((batch, self.sample_count()), dtype=np.float32)
        raise_for_status(_lib.bn_ctx_read_normalized(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
        return out

    def set_profiling(self, enabled: bool) -> None:
        raise_for_status(_lib.bn_ctx_set_profiling(self._h, 1 if enabled else 0))

    def stage_times(self):
        """[(stage name, device ms)] of the last run (profiling must be on)."""
        n = C.c_uint64()
        cap = 512
        ms = (C.c_float * cap)()
        names = ((C.c_char * 48) * cap)()
        raise_for_status(_lib.bn_ctx_stage_times(self._h, ms, C.cast(names, C.c_void_p), cap, C.byref(n)))
        return [(names[i].value.decode(), float(ms[i])) for i in range(min(n.value, cap))]

    def enqueue_device(self, device_ptr: int, batch: int, fetch_outputs: bool = True) -> None:
        raise_for_status(_lib.bn_ctx_enqueue_device(self._h, C.c_void_p(device_ptr), batch, 1 if fetch_outputs else 0))

    def wait(self, options: Optional[InferenceOptions] = None):
        ro, timeout = _run_opts(options)
        out = _ffi.Outputs()
        raise_for_status(_lib.bn_ctx_wait(self._h, C.byref(ro) if ro is not None else None, C.byref(out)), timeout)
        return out

    def stream_ptr(self) -> int:
        return int(_lib.bn_ctx_stream(self._h) or 0)

    def run_device(self, device_ptr: int, batch: int, fetch_outputs: bool = True,
                   options: Optional[InferenceOptions] = None):
        """Run on a [batch, sample_count] FP32 buffer already resident on the engine's GPU."""
        ro, timeout = _run_opts(options)
        out = _ffi.Outputs()
        st = _lib.bn_ctx_run_device(self._h, C.c_void_p(device_ptr), batch, 1 if fetch_outputs else 0,
                                    C.byref(ro) if ro is not None else None, C.byref(out))
        raise_for_status(st, timeout)
        return out


class Classifier:
    """src/classifier.rs:446-1058.  Clone + Send + Sync in the reference; here one object may
    be shared by threads: predict*/predict_batch use one internal context per calling thread."""

    def __init__(self, handle, config, labels, requested, top_k, min_confidence, info):
        self._h = handle
        self._pool: List[np.ndarray] = []
        self._pool_lock = threading.Lock()
        self._config = config
        self._labels = labels
        self._requested = requested
        self._top_k = top_k
        self._min_confidence = min_confidence
        self._info = info

    def __del__(self, _destroy=_lib.bn_engine_destroy):  # bound now: module globals may be gone at interpreter shutdown
        h, self._h = getattr(self, "_h", None), None
        if h:
            _destroy(h)

    @staticmethod
    def builder() -> ClassifierBuilder:
        return ClassifierBuilder()

    def config(self) -> ModelConfig:
        return self._config

    def labels(self) -> List[str]:
        return self._labels

    def requested_provider(self) -> ExecutionProviderInfo:
        return self._requested

    def io_info(self):
        return self._info

    # ---- fused range mask (rangefilter.rs:333-386 as a dense per-class tri-state) ----------
    def set_range_filter(self, location_scores: Sequence[LocationScore], threshold: float,
                         rerank: bool) -> None:
        """Fuse filter_predictions into the device epilogue.  Keyed by label *string* exactly
        like the reference's HashMap<&str, f32> (last duplicate wins)."""
        from .rangefilter import dense_range_state
        state, score = dense_range_state(self._labels, location_scores, threshold)
        raise_for_status(_lib.bn_engine_set_range_filter(
            self._h, state.ctypes.data_as(C.POINTER(C.c_uint8)),
            score.ctypes.data_as(C.POINTER(C.c_float)), len(state), 1 if rerank else 0))

    def set_range_filter_dense(self, state: np.ndarray, score: np.ndarray, rerank: bool) -> None:
        """Same, from the dense form directly: state[i] in {0 absent -> keep, 1 keep (x score when rerank), 2 drop}."""
        state = np.ascontiguousarray(state, dtype=np.uint8)
        score = np.ascontiguousarray(score, dtype=np.float32)
        raise_for_status(_lib.bn_engine_set_range_filter(
            self._h, state.ctypes.data_as(C.POINTER(C.c_uint8)),
            score.ctypes.data_as(C.POINTER(C.c_float)), len(state), 1 if rerank else 0))

    def compute_lanes(self) -> int:
        """Number of compute lanes of the engine (contexts are assigned to them in turn)."""
        return int(_lib.bn_engine_compute_lanes(self._h))

    def clear_range_filter(self) -> None:
        raise_for_status(_lib.bn_engine_clear_range_filter(self._h))

    # ---- inference ---------------------------------------------------------------------
    def predict(self, segment, options: Optional[InferenceOptions] = None) -> PredictionResult:
        seg = np.ascontiguousarray(segment, dtype=np.float32).reshape(-1)
        expected = self._config.sample_count
        if seg.shape[0] != expected:                           # classifier.rs:612-618
            raise InputSize(expected, int(seg.shape[0]))
        return self._run_engine([seg], options)[0]

    def predict_batch(self, segments: Sequence, options: Optional[InferenceOptions] = None) -> List[PredictionResult]:
        if len(segments) == 0:                                 # classifier.rs:681-683
            return []
        return self._run_engine(segments, options)

    def create_batch_context(self, max_batch_size: int, allow_perch: bool = False) -> BatchInferenceContext:
        """classifier.rs:777-792.  `allow_perch=True` (not in the reference) lifts the PerchV2
        refusal of batch_context.rs:107-114 and serves Perch through the same staged path."""
        h = C.c_void_p()
        if allow_perch:
            raise_for_status(_lib.bn_ctx_create_ex(self._h, int(max_batch_size), _ffi.BN_CTX_ALLOW_PERCH, C.byref(h)))
        else:
            raise_for_status(_lib.bn_ctx_create(self._h, int(max_batch_size), C.byref(h)))
        return BatchInferenceContext(self, h, int(max_batch_size))

    def predict_batch_with_context(self, context: BatchInferenceContext, segments: Sequence,
                                   options: Optional[InferenceOptions] = None) -> List[PredictionResult]:
        if len(segments) == 0:                                 # classifier.rs:832-834
            return []
        ptrs, lens, keep = _segment_arrays(segments)
        ro, timeout = _run_opts(options)
        out = _ffi.Outputs()
        st = _lib.bn_ctx_run(context._h, ptrs, lens, len(segments),
                             C.byref(ro) if ro is not None else None, C.byref(out))
        raise_for_status(st, timeout)
        return self._results(out)

    def predict_pcm16_stream(self, context: BatchInferenceContext, pcm, overlap_secs: float = 0.0,
                             options: Optional[InferenceOptions] = None):
        """The analysis loop of the reference CLI (src/bin/birdnet-analyze.rs:520-640) for a recording held as
        16-bit mono PCM at the model's sample rate: read_wav's i16 -> f32 / 32768 conversion (653-704) and
        chunk_audio (707-743) run on the device, so the recording crosses PCIe once at 2 bytes per sample.
        Returns [(start_time_secs, PredictionResult)] in recording order, like the CLI's `(f32, Vec<f32>)`
        segment list joined with its results."""
        pcm = np.ascontiguousarray(pcm)
        if pcm.dtype != np.int16 or pcm.ndim != 1:
            from .errors import AudioFormat
            raise AudioFormat("PCM must be a 1-D int16 array (mono, 16-bit)")
        if pcm.size == 0:                                        # birdnet-analyze.rs:694-698
            from .errors import AudioFormat
            raise AudioFormat("WAV file has no samples")
        cfg = self._config
        S, sr = cfg.sample_count, cfg.sample_rate
        # f32 arithmetic and truncation exactly as `(overlap_secs * sample_rate as f32) as usize` (712-718)
        overlap = int(np.float32(overlap_secs) * np.float32(sr))
        step = max(S - max(overlap, 0), 0)                       # saturating_sub
        if step == 0:
            return []
        n = int(pcm.size)
        n_seg = (n + step - 1) // step                           # while pos < samples.len()
        B = context.max_batch_size()
        ro, timeout = _run_opts(options)
        results = []
        ptr = pcm.ctypes.data_as(C.c_void_p)
        for first in range(0, n_seg, B):
            nb = min(B, n_seg - first)
            out = _ffi.Outputs()
            st = _lib.bn_ctx_run_pcm16(context._h, ptr, n, first * step, step, nb,
                                       C.byref(ro) if ro is not None else None, C.byref(out))
            raise_for_status(st, timeout)
            res = self._results(out)
            for j, r in enumerate(res):
                results.append((float(np.float32((first + j) * step) / np.float32(sr)), r))
        return results

    def _run_engine(self, segments, options) -> List[PredictionResult]:
        ptrs, lens, keep = _segment_arrays(segments)
        ro, timeout = _run_opts(options)
        out = _ffi.Outputs()
        st = _lib.bn_engine_run(self._h, ptrs, lens, len(segments),
                                C.byref(ro) if ro is not None else None, C.byref(out))
        raise_for_status(st, timeout)
        return self._results(out)

    _BLOCK_POOL_MAX = 16

    def _owned_block(self, rows: int, cols: int) -> np.ndarray:
        """[rows, cols] f32 block the results of one call own (their raw_scores / embeddings rows are views of it).
        Blocks whose results the caller has dropped are handed out again, so a steady stream of calls touches no fresh
        pages: a block is free exactly when nothing but the pool references it (every view of it, however derived,
        holds a reference to it through `.base`)."""
        shape = (rows, cols)
        with self._pool_lock:
            for blk in self._pool:
                if blk.shape == shape and sys.getrefcount(blk) == 3:     # the pool's list, `blk`, getrefcount's argument
                    return blk.view()
            blk = np.empty(shape, dtype=np.float32)
            if len(self._pool) >= self._BLOCK_POOL_MAX:
                self._pool.pop(0)
            self._pool.append(blk)
            return blk.view()

    def _results(self, out) -> List[PredictionResult]:
        """process_batch_outputs_from_flat (classifier.rs:872-911) on the borrowed slabs."""
        B, N, E, K = int(out.batch), int(out.num_species), int(out.embedding_dim), int(out.topk_stride)
        mt = self._config.model_type
        # the copies out of the borrowed slabs (raw_scores / embeddings are owned Vec<f32> in the reference) go through
        # memmove: a foreign call, so other driver threads run while the 6.7 MB move
        logits = self._owned_block(B, N)
        C.memmove(logits.ctypes.data, out.logits, logits.nbytes)
        emb = None
        if E and out.embeddings:
            emb = self._owned_block(B, E)
            C.memmove(emb.ctypes.data, out.embeddings, emb.nbytes)
        counts = np.ctypeslib.as_array(out.topk_count, shape=(B,)).tolist()
        labels = self._labels
        nl = len(labels)
        if K:
            raw = np.ctypeslib.as_array(C.cast(out.topk, C.POINTER(C.c_uint32)), shape=(B, K, 2))
            idx = raw[:, :, 0].tolist()
            conf = raw[:, :, 1].copy().view(np.float32).tolist()
        results = []
        for i in range(B):
            preds = []
            if K:
                ii, cc = idx[i], conf[i]
                for j in range(counts[i]):
                    k = ii[j]
                    preds.append(Prediction(labels[k] if k < nl else f"unknown_{k}", cc[j], k))
            results.append(PredictionResult(mt, preds, None if emb is None else emb[i], logits[i]))
        return results
