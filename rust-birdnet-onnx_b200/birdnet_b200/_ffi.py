"""ctypes binding of include/birdnet_b200.h (the C-ABI drop-in boundary).

There is no fallback: if lib/libbirdnet_b200.so is missing this module raises at import, and
every compute entry point fails loudly when no sm_100 GPU is present (BN_ERR_RUNTIME_INIT).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libbirdnet_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C rust-birdnet-onnx_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

BN_MAX_DIMS, BN_MAX_OUTPUTS = 8, 8

# status codes (bn_status)
(BN_OK, BN_ERR_INPUT_SIZE, BN_ERR_BATCH_INPUT_SIZE, BN_ERR_MODEL_DETECTION, BN_ERR_LABEL_COUNT,
 BN_ERR_MODEL_PATH_REQUIRED, BN_ERR_LABELS_REQUIRED, BN_ERR_MODEL_LOAD, BN_ERR_LABEL_LOAD,
 BN_ERR_LABEL_PARSE, BN_ERR_INFERENCE, BN_ERR_INVALID_COORDINATES, BN_ERR_INVALID_DATE,
 BN_ERR_RANGE_FILTER_INFERENCE, BN_ERR_TIMEOUT, BN_ERR_CANCELLED, BN_ERR_RUNTIME_INIT) = range(17)
BN_ERR_INVALID_ARGUMENT = 19
BN_CTX_ALLOW_PERCH = 1          # bn_ctx_create_ex flag


class DeviceCfg(C.Structure):
    _fields_ = [("device_id", C.c_int32), ("model_type_override", C.c_int32),
                ("pack_threads", C.c_int32), ("reserved", C.c_int32)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("rank", C.c_int32), ("dims", C.c_int64 * BN_MAX_DIMS)]

    def shape(self):
        return [int(self.dims[i]) for i in range(self.rank)]


class IoInfo(C.Structure):
    _fields_ = [("input", TensorInfo), ("n_outputs", C.c_int32),
                ("outputs", TensorInfo * BN_MAX_OUTPUTS), ("model_type", C.c_int32),
                ("sample_rate", C.c_uint32), ("segment_duration", C.c_float),
                ("sample_count", C.c_uint64), ("num_species", C.c_uint64),
                ("embedding_dim", C.c_uint64)]


class RunOpts(C.Structure):
    _fields_ = [("cancel_flag", C.POINTER(C.c_int32)), ("has_timeout", C.c_int32),
                ("timeout_ns", C.c_uint64)]


class Pred(C.Structure):
    _fields_ = [("index", C.c_uint32), ("confidence", C.c_float)]


class Outputs(C.Structure):
    _fields_ = [("batch", C.c_uint64), ("num_species", C.c_uint64),
                ("logits", C.POINTER(C.c_float)), ("embedding_dim", C.c_uint64),
                ("embeddings", C.POINTER(C.c_float)), ("topk_stride", C.c_uint64),
                ("topk_count", C.POINTER(C.c_uint32)), ("topk", C.POINTER(Pred))]


_P = C.POINTER
_vp = C.c_void_p

SIGNATURES = {
    "bn_last_error": (C.c_char_p, []),
    "bn_last_error_detail": (None, [_P(C.c_uint64)]),
    "bn_version": (C.c_char_p, []),
    "bn_engine_create": (C.c_int, [C.c_char_p, _P(DeviceCfg), _P(_vp)]),
    "bn_engine_destroy": (None, [_vp]),
    "bn_engine_io_info": (C.c_int, [_vp, _P(IoInfo)]),
    "bn_model_inspect": (C.c_int, [C.c_char_p, C.c_int32, _P(IoInfo)]),
    "bn_model_plan_summary": (C.c_int, [C.c_char_p, C.c_int32, C.c_char_p, C.c_uint64, _P(C.c_uint64)]),
    "bn_detect_model_type": (C.c_int, [_P(C.c_int64), C.c_int32, _P(C.c_int64), _P(C.c_int32),
                                       C.c_int32, C.c_int32, _P(IoInfo)]),
    "bn_engine_set_postprocess": (C.c_int, [_vp, C.c_uint64, C.c_int32, C.c_float]),
    "bn_engine_set_range_filter": (C.c_int, [_vp, _P(C.c_uint8), _P(C.c_float), C.c_uint64, C.c_int32]),
    "bn_engine_clear_range_filter": (C.c_int, [_vp]),
    "bn_engine_run": (C.c_int, [_vp, _P(_vp), _P(C.c_uint64), C.c_uint64, _P(RunOpts), _P(Outputs)]),
    "bn_ctx_create": (C.c_int, [_vp, C.c_uint64, _P(_vp)]),
    "bn_ctx_create_ex": (C.c_int, [_vp, C.c_uint64, C.c_uint32, _P(_vp)]),
    "bn_ctx_destroy": (None, [_vp]),
    "bn_ctx_run": (C.c_int, [_vp, _P(_vp), _P(C.c_uint64), C.c_uint64, _P(RunOpts), _P(Outputs)]),
    "bn_ctx_max_batch_size": (C.c_uint64, [_vp]),
    "bn_ctx_input_buffer_bytes": (C.c_uint64, [_vp]),
    "bn_ctx_run_device": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int32, _P(RunOpts), _P(Outputs)]),
    "bn_ctx_run_pcm16": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P(RunOpts), _P(Outputs)]),
    "bn_ctx_enqueue_device": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int32]),
    "bn_ctx_wait": (C.c_int, [_vp, _P(RunOpts), _P(Outputs)]),
    "bn_ctx_read_tensor": (C.c_int, [_vp, C.c_char_p, _P(C.c_float), C.c_uint64, _P(C.c_uint64)]),
    "bn_ctx_read_normalized": (C.c_int, [_vp, _P(C.c_float), C.c_uint64]),
    "bn_ctx_last_launch_count": (C.c_uint64, [_vp]),
    "bn_ctx_last_run_in_place": (C.c_int, [_vp]),
    "bn_ctx_nonfinite_segments": (C.c_uint64, [_vp]),
    "bn_ctx_set_profiling": (C.c_int, [_vp, C.c_int32]),
    "bn_ctx_stage_times": (C.c_int, [_vp, _P(C.c_float), _vp, C.c_uint64, _P(C.c_uint64)]),
    "bn_ctx_stream": (_vp, [_vp]),
    "bn_engine_compute_lanes": (C.c_int32, [_vp]),
    "bn_range_filter_apply": (C.c_int, [_vp, _P(Pred), _P(C.c_uint32), C.c_uint64, C.c_uint64,
                                        _P(C.c_uint8), _P(C.c_float), C.c_uint64, C.c_int32,
                                        _P(Pred), _P(C.c_uint32)]),
    "bn_topk_apply": (C.c_int, [_vp, _P(C.c_float), C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32,
                                C.c_float, _P(C.c_uint8), _P(C.c_float), C.c_int32, _P(Pred),
                                _P(C.c_uint32)]),
    "bn_pool_create": (C.c_int, [C.c_char_p, _P(C.c_int32), C.c_int32, C.c_int32, C.c_uint64, _P(_vp)]),
    "bn_pool_create_ex": (C.c_int, [C.c_char_p, _P(C.c_int32), C.c_int32, C.c_int32, C.c_uint64, C.c_int32, _P(_vp)]),
    "bn_pool_destroy": (None, [_vp]),
    "bn_pool_set_postprocess": (C.c_int, [_vp, C.c_uint64, C.c_int32, C.c_float]),
    "bn_pool_set_range_filter": (C.c_int, [_vp, _P(C.c_uint8), _P(C.c_float), C.c_uint64, C.c_int32]),
    "bn_pool_run": (C.c_int, [_vp, _P(_vp), _P(C.c_uint64), C.c_uint64, _P(RunOpts), _P(C.c_float),
                              _P(C.c_float), _P(Pred), _P(C.c_uint32), C.c_uint64]),
    "bn_meta_create": (C.c_int, [C.c_char_p, C.c_int32, _P(_vp)]),
    "bn_meta_destroy": (None, [_vp]),
    "bn_meta_num_outputs": (C.c_uint64, [_vp]),
    "bn_meta_predict": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float, _P(C.c_float), C.c_uint64]),
    "bn_meta_install_range_filter": (C.c_int, [_vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32]),
    "bn_device_count": (C.c_int, []),
    "bn_host_alloc": (C.c_void_p, [C.c_uint64]),
    "bn_host_free": (None, [C.c_void_p]),
    "bn_host_register": (C.c_int, [C.c_void_p, C.c_uint64]),
    "bn_host_unregister": (C.c_int, [C.c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.bn_last_error() or b"").decode("utf-8", "replace")


def last_error_detail():
    d = (C.c_uint64 * 3)()
    lib.bn_last_error_detail(d)
    return int(d[0]), int(d[1]), int(d[2])
