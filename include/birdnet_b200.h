/* birdnet_b200 — C ABI of the B200-native batched inference path.
 *
 * Drop-in boundary.  The reference (tphakala/rust-birdnet-onnx) has no
 * FFI of its own (`unsafe_code = "deny"`, Cargo.toml:41); its seam is the set of calls it makes
 * into the `ort` crate.  Each entry point below names the reference call site it replaces.
 * Plain C: opaque handles, pointers and sizes only.  Every function returns a bn_status
 * (0 = ok); the message of the last failure on the calling thread is bn_last_error().
 *
 * Threading: a bn_engine may be shared by threads (concurrent bn_engine_run calls each check out one of
 * up to four internal contexts and queue when all are busy - the reference serialises on its
 * Mutex<Session>, src/classifier.rs:435); a bn_ctx is
 * single-threaded, one per thread (src/batch_context.rs:56-60).
 */
#ifndef BIRDNET_B200_H
#define BIRDNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: one per variant of the reference's Error enum (src/error.rs:6-128). */
typedef enum bn_status {
    BN_OK = 0,
    BN_ERR_INPUT_SIZE = 1,             /* Error::InputSize            error.rs:8-14   */
    BN_ERR_BATCH_INPUT_SIZE = 2,       /* Error::BatchInputSize       error.rs:17-25  */
    BN_ERR_MODEL_DETECTION = 3,        /* Error::ModelDetection       error.rs:28-32  */
    BN_ERR_LABEL_COUNT = 4,            /* Error::LabelCount           error.rs:35-41  */
    BN_ERR_MODEL_PATH_REQUIRED = 5,    /* Error::ModelPathRequired    error.rs:44-45  */
    BN_ERR_LABELS_REQUIRED = 6,        /* Error::LabelsRequired       error.rs:48-49  */
    BN_ERR_MODEL_LOAD = 7,             /* Error::ModelLoad            error.rs:52-53  */
    BN_ERR_LABEL_LOAD = 8,             /* Error::LabelLoad            error.rs:56-62  */
    BN_ERR_LABEL_PARSE = 9,            /* Error::LabelParse           error.rs:65-66  */
    BN_ERR_INFERENCE = 10,             /* Error::Inference            error.rs:69-70  */
    BN_ERR_INVALID_COORDINATES = 11,   /* Error::InvalidCoordinates   error.rs:73-81  */
    BN_ERR_INVALID_DATE = 12,          /* Error::InvalidDate          error.rs:84-92  */
    BN_ERR_RANGE_FILTER_INFERENCE = 13,/* Error::RangeFilterInference error.rs:95-96  */
    BN_ERR_TIMEOUT = 14,               /* Error::Timeout              error.rs:99-103 */
    BN_ERR_CANCELLED = 15,             /* Error::Cancelled            error.rs:106-107*/
    BN_ERR_RUNTIME_INIT = 16,          /* Error::RuntimeInit          error.rs:110-111 (no usable CUDA device) */
    BN_ERR_INVALID_ARGUMENT = 19
} bn_status;

/* ModelType (src/types.rs:3-11) */
enum { BN_MODEL_AUTO = -1, BN_MODEL_BIRDNET_V24 = 0, BN_MODEL_BIRDNET_V30 = 1, BN_MODEL_PERCH_V2 = 2 };

typedef struct bn_engine bn_engine;     /* replaces ort::session::Session        */
typedef struct bn_ctx bn_ctx;           /* replaces ort::io_binding::IoBinding + the host slab */
typedef struct bn_pool bn_pool;         /* one engine replica + ctx per GPU, host-side gather */

typedef struct bn_device_cfg {
    int32_t device_id;            /* CUDAConfig::with_device_id, src/cuda_config.rs:179-182 */
    int32_t model_type_override;  /* BN_MODEL_AUTO or a ModelType (ClassifierBuilder::model_type,
                                     src/classifier.rs:98-102) */
    int32_t pack_threads;         /* host threads gathering segments into pinned staging; 0 = auto */
    int32_t reserved;
} bn_device_cfg;

#define BN_MAX_DIMS 8
#define BN_MAX_OUTPUTS 8
typedef struct bn_tensor_info {
    char name[64];
    int32_t rank;
    int64_t dims[BN_MAX_DIMS];    /* -1 = dynamic, as ort reports it */
} bn_tensor_info;

/* What session.inputs()/outputs() + detect_model_type give the reference
 * (src/classifier.rs:353-357, 387-420; src/detection.rs:15-80; src/types.rs:72-85). */
typedef struct bn_io_info {
    bn_tensor_info input;
    int32_t n_outputs;
    bn_tensor_info outputs[BN_MAX_OUTPUTS];
    int32_t model_type;
    uint32_t sample_rate;
    float segment_duration;
    uint64_t sample_count;
    uint64_t num_species;
    uint64_t embedding_dim;       /* 0 = model has no embeddings (v2.4) */
} bn_io_info;

/* RunOptions + the monitor thread of Classifier::run_inference (src/classifier.rs:504-574):
 * the engine polls while it waits for the device.  cancel_flag may be NULL. */
typedef struct bn_run_opts {
    const volatile int32_t* cancel_flag;  /* CancellationToken: non-zero = cancelled */
    int32_t has_timeout;
    uint64_t timeout_ns;
} bn_run_opts;

typedef struct bn_pred {
    uint32_t index;               /* Prediction::index      src/types.rs:95 */
    float confidence;             /* Prediction::confidence src/types.rs:93 */
} bn_pred;

/* Borrowed, engine-owned host (pinned) buffers, valid until the next run on the same ctx /
 * engine — the reference holds the session lock while outputs exist (classifier.rs:628-630). */
typedef struct bn_outputs {
    uint64_t batch;
    uint64_t num_species;
    const float* logits;          /* [batch][num_species]  PredictionResult::raw_scores */
    uint64_t embedding_dim;
    const float* embeddings;      /* [batch][embedding_dim] or NULL */
    uint64_t topk_stride;         /* slots per segment = min(top_k, num_species) */
    const uint32_t* topk_count;   /* [batch] predictions that survived min_confidence / range mask */
    const bn_pred* topk;          /* [batch][topk_stride], first topk_count[i] valid, sorted */
} bn_outputs;

/* ---- errors ---------------------------------------------------------------------- */
const char* bn_last_error(void);                 /* payload text ({0} / {reason}) of the last failure */
void bn_last_error_detail(uint64_t out[3]);      /* BatchInputSize: {index, expected, got}; InputSize: {0, expected, got} */
const char* bn_version(void);

/* ---- load time --------------------------------------------------------------------
 * bn_engine_create  <- Session::builder().with_execution_providers(..).commit_from_file(path)
 *                      src/classifier.rs:340-350 (+ detect_model_type, classifier.rs:357)
 * bn_engine_io_info <- session.inputs()/outputs() shapes, src/classifier.rs:387-420
 * bn_model_inspect  :  same parse + detection with no GPU touched (load-time checks, tests) */
int bn_engine_create(const char* onnx_path, const bn_device_cfg* cfg, bn_engine** out);
void bn_engine_destroy(bn_engine* engine);
int bn_engine_io_info(const bn_engine* engine, bn_io_info* out);
int bn_model_inspect(const char* onnx_path, int32_t model_type_override, bn_io_info* out);
/* detect_model_type on explicit shapes (src/detection.rs:15-80); shapes are rank-prefixed rows */
/* Canonical text of the layer plan the file is matched into (front-end constants, one line per fused op with shapes,
 * activation, gate / residual flags and a hash of the weight bits; value names left out): two files describing the same
 * network give the same text whichever exporter wrote them.  buf may be NULL to query `needed` (bytes incl. NUL). */
int bn_model_plan_summary(const char* onnx_path, int32_t model_type_override, char* buf, uint64_t cap, uint64_t* needed);
int bn_detect_model_type(const int64_t* input_dims, int32_t input_rank, const int64_t* output_dims,
                         const int32_t* output_ranks, int32_t n_outputs, int32_t model_type_override,
                         bn_io_info* out);

/* ---- fused epilogue configuration ---------------------------------------------------
 * top_k / min_confidence <- ClassifierBuilder::top_k / min_confidence (classifier.rs:118-129)
 *                           consumed by top_k_predictions (src/postprocess.rs:40-87)
 * range filter           <- filter_predictions_impl (src/rangefilter.rs:333-386) as a dense
 *                           per-class tri-state: 0 absent (keep), 1 keep (x score if rerank), 2 drop */
int bn_engine_set_postprocess(bn_engine* engine, uint64_t top_k, int32_t has_min_confidence, float min_confidence);
int bn_engine_set_range_filter(bn_engine* engine, const uint8_t* state, const float* score, uint64_t n, int32_t rerank);
int bn_engine_clear_range_filter(bn_engine* engine);

/* ---- hot path ----------------------------------------------------------------------
 * bn_engine_run <- Value::from_array + session.run_with_options (classifier.rs:698-723)
 *                  validation order of predict_batch (classifier.rs:681-696)
 * bn_ctx_create <- BatchInferenceContext::new / session.create_binding (batch_context.rs:102-133)
 * bn_ctx_run    <- prepare_input + bind_outputs_to_device + run_binding_with_options +
 *                  synchronize + extract_outputs (batch_context.rs:188-338, classifier.rs:839-865)
 * seg_lens[i] is the sample count of segment i (slices carry their length in Rust).
 * bn_engine_run takes a batch of ANY size (like predict_batch): it runs chunks of <= 256 segments on one of a small
 * bounded pool of internal contexts and gathers the results in a per-thread host buffer; the pointers in `out`
 * stay valid until the calling thread's next bn_engine_run. */
int bn_engine_run(bn_engine* engine, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
                  const bn_run_opts* opts, bn_outputs* out);
int bn_ctx_create(bn_engine* engine, uint64_t max_batch_size, bn_ctx** out);
/* bn_ctx_create with flags.  BN_CTX_ALLOW_PERCH lifts the reference's refusal of
 * PerchV2 models (batch_context.rs:107-114): the staged context path then serves all four Perch outputs by index
 * (embedding [0] and logits [3] in bn_outputs; spatial_embedding / spectrogram through bn_ctx_read_tensor). */
enum { BN_CTX_ALLOW_PERCH = 1 };
int bn_ctx_create_ex(bn_engine* engine, uint64_t max_batch_size, uint32_t flags, bn_ctx** out);
void bn_ctx_destroy(bn_ctx* ctx);
int bn_ctx_run(bn_ctx* ctx, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
               const bn_run_opts* opts, bn_outputs* out);
uint64_t bn_ctx_max_batch_size(const bn_ctx* ctx);     /* batch_context.rs:137-140 */
uint64_t bn_ctx_input_buffer_bytes(const bn_ctx* ctx); /* batch_context.rs:155-158 */
/* Device-resident variant: `d_audio` is a [batch][sample_count] FP32 device buffer on the engine's
 * GPU (bench: inputs already in HBM).  fetch_outputs != 0 copies results to the pinned host slab. */
int bn_ctx_run_device(bn_ctx* ctx, const float* d_audio, uint64_t batch, int32_t fetch_outputs,
                      const bn_run_opts* opts, bn_outputs* out);
/* CLI ingest with the conversion and chunking on the device (replaces, for callers
 * that hold the recording as 16-bit PCM, read_wav's i16 -> f32 / 32768 (src/bin/birdnet-analyze.rs:21, 684-687)
 * and chunk_audio (src/bin/birdnet-analyze.rs:707-743)): segment b of this call covers samples
 * [first_pos + b*step, first_pos + b*step + sample_count) of `pcm` (n_samples long), zero-padded past the end.
 * step = sample_count - overlap_samples, 1 <= step <= sample_count; batch <= max_batch_size.  The recording
 * crosses PCIe once, 2 bytes per sample, overlap not duplicated.  Outputs as bn_ctx_run. */
int bn_ctx_run_pcm16(bn_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t first_pos, uint64_t step,
                     uint64_t batch, const bn_run_opts* opts, bn_outputs* out);
/* Asynchronous halves of bn_ctx_run_device: enqueue returns once the work is on the stream;
 * several enqueues may be queued back to back (each overwrites the previous outputs); wait blocks
 * (polling opts) until everything enqueued so far has finished and exposes the last outputs. */
int bn_ctx_enqueue_device(bn_ctx* ctx, const float* d_audio, uint64_t batch, int32_t fetch_outputs);
int bn_ctx_wait(bn_ctx* ctx, const bn_run_opts* opts, bn_outputs* out);
/* Intermediate tensors of the last run, by ONNX value name ("spec", ...) — parity tests. */
int bn_ctx_read_tensor(bn_ctx* ctx, const char* name, float* dst, uint64_t dst_elems, uint64_t* elems_out);
int bn_ctx_read_normalized(bn_ctx* ctx, float* dst, uint64_t dst_elems);
/* Kernel launches enqueued by the last run on this ctx, and per-stage device times (ms) of the
 * last run when profiling was enabled with bn_ctx_set_profiling(ctx, 1). */
uint64_t bn_ctx_last_launch_count(const bn_ctx* ctx);
/* 1 when the last bn_ctx_run found every segment in page-locked host memory and copied host -> device straight from the
 * caller's slices; 0 when it gathered them into the staging slab first (batch_context.rs:199-211). */
int bn_ctx_last_run_in_place(const bn_ctx* ctx);
/* Segments of the last completed run on this ctx whose logits were not all finite.  The tensor-core path keeps
 * operands as fp16 hi + fp16 lo pairs: ~22 mantissa bits but the fp16 RANGE, so an activation beyond +-65504 turns
 * into NaN downstream (weights beyond it are refused at load time with BN_ERR_MODEL_LOAD).  The run still returns
 * BN_OK - the reference also returns NaN scores for NaN input - this counter lets a caller tell. */
uint64_t bn_ctx_nonfinite_segments(const bn_ctx* ctx);
int bn_ctx_set_profiling(bn_ctx* ctx, int32_t enabled);
int bn_ctx_stage_times(const bn_ctx* ctx, float* ms_out, char (*names_out)[48], uint64_t cap, uint64_t* n_out);
void* bn_ctx_stream(bn_ctx* ctx);                      /* cudaStream_t the kernels run on */
/* Compute lanes of the engine: contexts are assigned to them in turn at creation; batches of contexts on the same
 * lane run back to back, batches on different lanes overlap (BN_COMPUTE_LANES=1|2, default 2). */
int32_t bn_engine_compute_lanes(const bn_engine* engine);

/* RangeFilter::filter_predictions / filter_batch_predictions on lists already selected
 * (src/rangefilter.rs:527-579), executed on the engine's device. */
int bn_range_filter_apply(bn_engine* engine, const bn_pred* in, const uint32_t* in_count, uint64_t rows,
                          uint64_t stride, const uint8_t* state, const float* score, uint64_t n, int32_t rerank,
                          bn_pred* out, uint32_t* out_count);
/* top_k_predictions on caller-supplied logits (src/postprocess.rs:40-87) through the same
 * epilogue kernel: known-answer tests of the reference run against this. */
int bn_topk_apply(bn_engine* engine, const float* logits, uint64_t rows, uint64_t n, uint64_t top_k,
                  int32_t has_min_confidence, float min_confidence, const uint8_t* state, const float* score,
                  int32_t rerank, bn_pred* out, uint32_t* out_count);

/* ---- multi-GPU: shared batch queue, no collective --------------
 * Per device one engine replica and `depth` contexts (each with its own host thread inside bn_pool_run); all of them
 * pull whole batches of ctx_batch segments from one queue, so H2D, kernels and D2H of consecutive batches overlap on
 * every device.  bn_pool_create = depth 3 (or BN_POOL_DEPTH).  A device id may be listed more than once. */
int bn_pool_create(const char* onnx_path, const int32_t* device_ids, int32_t n_devices, int32_t model_type_override,
                   uint64_t ctx_batch, bn_pool** out);
int bn_pool_create_ex(const char* onnx_path, const int32_t* device_ids, int32_t n_devices, int32_t model_type_override,
                      uint64_t ctx_batch, int32_t depth, bn_pool** out);
void bn_pool_destroy(bn_pool* pool);
int bn_pool_set_postprocess(bn_pool* pool, uint64_t top_k, int32_t has_min_confidence, float min_confidence);
int bn_pool_set_range_filter(bn_pool* pool, const uint8_t* state, const float* score, uint64_t n, int32_t rerank);
/* Runs all segments; results are gathered in caller order into caller-provided host arrays:
 * logits [n_segments][num_species] (or NULL), embeddings (or NULL), topk [n_segments][topk_stride].
 * opts (timeout / cancellation, src/inference_options.rs) applies to every batch of the call. */
int bn_pool_run(bn_pool* pool, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t n_segments,
                const bn_run_opts* opts, float* logits, float* embeddings, bn_pred* topk, uint32_t* topk_count,
                uint64_t topk_stride);
/* Range-filter meta model.  Replaces the ONNX Runtime session RangeFilter owns:
 * bn_meta_create   <- Session::builder()...commit_from_file + the "exactly one output" check (src/rangefilter.rs:239-258);
 *                     the caller compares bn_meta_num_outputs with its label count (LabelCount, 260-266)
 * bn_meta_predict  <- session.run on [1,3] = [latitude, longitude, week] + try_extract_tensor (src/rangefilter.rs:451-479);
 *                     coordinate / date validation, calculate_week, thresholding and the sort stay in the host facade
 * bn_meta_install_range_filter: the same forward pass, then the dense per-class tri-state of filter_predictions_impl
 *                     (src/rangefilter.rs:333-386) built on the device and installed as the engine's fused range filter:
 *                     species with score >= predict_threshold are "in the map" (predict returns only those, 482-496);
 *                     in the map, score >= filter_threshold keeps (x score when rerank) and lower drops; the rest keep
 *                     unchanged.  Requires the meta model's label list to be the classifier's (same order). */
typedef struct bn_meta bn_meta;
int bn_meta_create(const char* onnx_path, int32_t device_id, bn_meta** out);
void bn_meta_destroy(bn_meta* meta);
uint64_t bn_meta_num_outputs(const bn_meta* meta);
int bn_meta_predict(bn_meta* meta, float latitude, float longitude, float week, float* scores, uint64_t n);
int bn_meta_install_range_filter(bn_meta* meta, bn_engine* engine, float latitude, float longitude, float week,
                                 float predict_threshold, float filter_threshold, int32_t rerank);
int bn_device_count(void);
/* Page-locked host memory for callers that want their segments DMA-able in place: when every segment pointer
 * handed to bn_ctx_run / bn_engine_run lies in page-locked host memory (from here, cudaHostAlloc or
 * cudaHostRegister), the engine skips the gather into its own staging slab (the memcpy of
 * batch_context.rs:199-211) and copies host -> device straight from the caller's slices. */
void* bn_host_alloc(uint64_t bytes);
void bn_host_free(void* p);
/* Page-lock memory the caller already owns (its ring buffer, a decoded recording) in place - cudaHostRegister without the
 * caller linking the CUDA runtime.  Segments inside a registered range then take the same gather-free path; the literal
 * pageable case is bounded by the host's pageable -> pinned memcpy rate.  Returns BN_OK or
 * BN_ERR_INVALID_ARGUMENT (null / empty range, range already registered, registration refused by the driver).  Unregister
 * before freeing the memory. */
int bn_host_register(void* p, uint64_t bytes);
int bn_host_unregister(void* p);

#ifdef __cplusplus
}
#endif
#endif /* BIRDNET_B200_H */
