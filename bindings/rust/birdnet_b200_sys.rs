//! Raw FFI declarations of `include/birdnet_b200.h` for the reference crate (tphakala/rust-birdnet-onnx).
//!
//! Source-only: the build image has no Rust toolchain, so this file is not compiled here.  It is kept in lock step with
//! the header by `tests/test_abi.py::test_rust_binding_lists_every_symbol` (every exported symbol must be declared).
//! Drop it into the reference as `src/b200_sys.rs` with `#![allow(unsafe_code)]` for this module only
//! (the crate denies `unsafe_code`, Cargo.toml:41) and link with
//! `cargo:rustc-link-lib=dylib=birdnet_b200` from `build.rs`.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const BN_OK: c_int = 0;
pub const BN_ERR_INPUT_SIZE: c_int = 1;
pub const BN_ERR_BATCH_INPUT_SIZE: c_int = 2;
pub const BN_ERR_MODEL_DETECTION: c_int = 3;
pub const BN_ERR_LABEL_COUNT: c_int = 4;
pub const BN_ERR_MODEL_PATH_REQUIRED: c_int = 5;
pub const BN_ERR_LABELS_REQUIRED: c_int = 6;
pub const BN_ERR_MODEL_LOAD: c_int = 7;
pub const BN_ERR_LABEL_LOAD: c_int = 8;
pub const BN_ERR_LABEL_PARSE: c_int = 9;
pub const BN_ERR_INFERENCE: c_int = 10;
pub const BN_ERR_INVALID_COORDINATES: c_int = 11;
pub const BN_ERR_INVALID_DATE: c_int = 12;
pub const BN_ERR_RANGE_FILTER_INFERENCE: c_int = 13;
pub const BN_ERR_TIMEOUT: c_int = 14;
pub const BN_ERR_CANCELLED: c_int = 15;
pub const BN_ERR_RUNTIME_INIT: c_int = 16;
pub const BN_ERR_INVALID_ARGUMENT: c_int = 19;
pub const BN_CTX_ALLOW_PERCH: u32 = 1;

#[repr(C)] pub struct bn_device_cfg { pub device_id: i32, pub model_type_override: i32, pub pack_threads: i32, pub reserved: i32 }
#[repr(C)] pub struct bn_tensor_info { pub name: [c_char; 64], pub rank: i32, pub dims: [i64; 8] }
#[repr(C)] pub struct bn_io_info {
    pub input: bn_tensor_info, pub n_outputs: i32, pub outputs: [bn_tensor_info; 8],
    pub model_type: i32, pub sample_rate: u32, pub segment_duration: f32,
    pub sample_count: u64, pub num_species: u64, pub embedding_dim: u64,
}
#[repr(C)] pub struct bn_run_opts { pub cancel_flag: *const i32, pub has_timeout: i32, pub timeout_ns: u64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct bn_pred { pub index: u32, pub confidence: f32 }
#[repr(C)] pub struct bn_outputs {
    pub batch: u64, pub num_species: u64, pub logits: *const f32,
    pub embedding_dim: u64, pub embeddings: *const f32,
    pub topk_stride: u64, pub topk_count: *const u32, pub topk: *const bn_pred,
}

#[link(name = "birdnet_b200")]
extern "C" {
    pub fn bn_last_error() -> *const c_char;
    pub fn bn_last_error_detail(out: *mut u64);
    pub fn bn_version() -> *const c_char;
    pub fn bn_device_count() -> c_int;
    // Session::builder()...commit_from_file + session.inputs()/outputs() + detect_model_type (classifier.rs:340-420)
    pub fn bn_engine_create(path: *const c_char, cfg: *const bn_device_cfg, out: *mut *mut c_void) -> c_int;
    pub fn bn_engine_destroy(e: *mut c_void);
    pub fn bn_engine_io_info(e: *const c_void, out: *mut bn_io_info) -> c_int;
    pub fn bn_model_inspect(path: *const c_char, model_type_override: i32, out: *mut bn_io_info) -> c_int;
    pub fn bn_model_plan_summary(path: *const c_char, model_type_override: i32, buf: *mut c_char, cap: u64, needed: *mut u64) -> c_int;
    pub fn bn_detect_model_type(input_dims: *const i64, input_rank: i32, output_dims: *const i64, output_ranks: *const i32,
                                n_outputs: i32, model_type_override: i32, out: *mut bn_io_info) -> c_int;
    // fused epilogue: top_k / min_confidence (postprocess.rs:40-87) and the range mask (rangefilter.rs:333-386)
    pub fn bn_engine_set_postprocess(e: *mut c_void, top_k: u64, has_min: i32, min_conf: f32) -> c_int;
    pub fn bn_engine_set_range_filter(e: *mut c_void, state: *const u8, score: *const f32, n: u64, rerank: i32) -> c_int;
    pub fn bn_engine_clear_range_filter(e: *mut c_void) -> c_int;
    pub fn bn_engine_compute_lanes(e: *const c_void) -> i32;
    // predict / predict_batch (classifier.rs:610-727): any batch size, results valid until the thread's next call
    pub fn bn_engine_run(e: *mut c_void, segs: *const *const f32, lens: *const u64, batch: u64,
                         opts: *const bn_run_opts, out: *mut bn_outputs) -> c_int;
    // BatchInferenceContext (batch_context.rs:70-339) + predict_batch_with_context (classifier.rs:826-867)
    pub fn bn_ctx_create(e: *mut c_void, max_batch: u64, out: *mut *mut c_void) -> c_int;
    pub fn bn_ctx_create_ex(e: *mut c_void, max_batch: u64, flags: u32, out: *mut *mut c_void) -> c_int;
    pub fn bn_ctx_destroy(c: *mut c_void);
    pub fn bn_ctx_run(c: *mut c_void, segs: *const *const f32, lens: *const u64, batch: u64,
                      opts: *const bn_run_opts, out: *mut bn_outputs) -> c_int;
    pub fn bn_ctx_max_batch_size(c: *const c_void) -> u64;
    pub fn bn_ctx_input_buffer_bytes(c: *const c_void) -> u64;
    pub fn bn_ctx_run_device(c: *mut c_void, d_audio: *const f32, batch: u64, fetch_outputs: i32,
                             opts: *const bn_run_opts, out: *mut bn_outputs) -> c_int;
    pub fn bn_ctx_run_pcm16(c: *mut c_void, pcm: *const i16, n_samples: u64, first_pos: u64, step: u64, batch: u64,
                            opts: *const bn_run_opts, out: *mut bn_outputs) -> c_int;
    pub fn bn_ctx_enqueue_device(c: *mut c_void, d_audio: *const f32, batch: u64, fetch_outputs: i32) -> c_int;
    pub fn bn_ctx_wait(c: *mut c_void, opts: *const bn_run_opts, out: *mut bn_outputs) -> c_int;
    pub fn bn_ctx_read_tensor(c: *mut c_void, name: *const c_char, dst: *mut f32, dst_elems: u64, elems_out: *mut u64) -> c_int;
    pub fn bn_ctx_read_normalized(c: *mut c_void, dst: *mut f32, dst_elems: u64) -> c_int;
    pub fn bn_ctx_last_launch_count(c: *const c_void) -> u64;
    pub fn bn_ctx_last_run_in_place(c: *const c_void) -> c_int;
    pub fn bn_ctx_nonfinite_segments(c: *const c_void) -> u64;
    pub fn bn_ctx_set_profiling(c: *mut c_void, enabled: i32) -> c_int;
    pub fn bn_ctx_stage_times(c: *const c_void, ms_out: *mut f32, names_out: *mut [c_char; 48], cap: u64, n_out: *mut u64) -> c_int;
    pub fn bn_ctx_stream(c: *mut c_void) -> *mut c_void;
    // the same epilogue kernels on caller-supplied data
    pub fn bn_range_filter_apply(e: *mut c_void, inp: *const bn_pred, in_count: *const u32, rows: u64, stride: u64,
                                 state: *const u8, score: *const f32, n: u64, rerank: i32,
                                 out: *mut bn_pred, out_count: *mut u32) -> c_int;
    pub fn bn_topk_apply(e: *mut c_void, logits: *const f32, rows: u64, n: u64, top_k: u64, has_min: i32, min_conf: f32,
                         state: *const u8, score: *const f32, rerank: i32, out: *mut bn_pred, out_count: *mut u32) -> c_int;
    // multi-GPU dispatcher (CUDAConfig::with_device_id, cuda_config.rs:179-182)
    pub fn bn_pool_create(path: *const c_char, device_ids: *const i32, n_devices: i32, model_type_override: i32,
                          ctx_batch: u64, out: *mut *mut c_void) -> c_int;
    pub fn bn_pool_create_ex(path: *const c_char, device_ids: *const i32, n_devices: i32, model_type_override: i32,
                             ctx_batch: u64, depth: i32, out: *mut *mut c_void) -> c_int;
    pub fn bn_pool_destroy(p: *mut c_void);
    pub fn bn_pool_set_postprocess(p: *mut c_void, top_k: u64, has_min: i32, min_conf: f32) -> c_int;
    pub fn bn_pool_set_range_filter(p: *mut c_void, state: *const u8, score: *const f32, n: u64, rerank: i32) -> c_int;
    pub fn bn_pool_run(p: *mut c_void, segs: *const *const f32, lens: *const u64, n_segments: u64, opts: *const bn_run_opts,
                       logits: *mut f32, embeddings: *mut f32, topk: *mut bn_pred, topk_count: *mut u32, topk_stride: u64) -> c_int;
    // range-filter meta model (rangefilter.rs:239-267, 435-502)
    pub fn bn_meta_create(path: *const c_char, device_id: i32, out: *mut *mut c_void) -> c_int;
    pub fn bn_meta_destroy(m: *mut c_void);
    pub fn bn_meta_num_outputs(m: *const c_void) -> u64;
    pub fn bn_meta_predict(m: *mut c_void, lat: f32, lon: f32, week: f32, scores: *mut f32, n: u64) -> c_int;
    pub fn bn_meta_install_range_filter(m: *mut c_void, e: *mut c_void, lat: f32, lon: f32, week: f32,
                                        predict_thr: f32, filter_thr: f32, rerank: i32) -> c_int;
    // page-locked host memory: slices that live here are DMA'd in place
    pub fn bn_host_alloc(bytes: u64) -> *mut c_void;
    pub fn bn_host_free(p: *mut c_void);
    pub fn bn_host_register(p: *mut c_void, bytes: u64) -> c_int;
    pub fn bn_host_unregister(p: *mut c_void) -> c_int;
}
