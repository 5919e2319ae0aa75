// Engine internals shared by engine.cu / capi.cu / pool.cu.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/birdnet_b200.h"
#include "kernels.h"
#include "plan.h"
#include "mbconv.h"
#include "tc_conv.h"

namespace bn {

// ---- thread-local error channel ------------------------------------------------------
int set_error(int status, const std::string& msg);
int set_error_detail(int status, const std::string& msg, uint64_t a, uint64_t b, uint64_t c);
int cuda_fail(cudaError_t e, const char* what);   // -> BN_ERR_INFERENCE / BN_ERR_RUNTIME_INIT
const std::string& last_error();
const uint64_t* last_detail();
int load_plan(const char* path, int override_type, Plan& plan);

#define BN_CUDA(expr)                                                     \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return ::bn::cuda_fail(_e, #expr);         \
    } while (0)

struct DevOp {
    float* weight = nullptr;     // FP32 engine layout (CUDA-core kernels)
    float* bias = nullptr;
    // tensor-core path (tc_conv.cu)
    bool use_tc = false;
    void* wpack = nullptr;
    int nt = 0, n_tiles = 0, k_chunks = 0, stages = 2, tmem_cols = 32;
    int halo_slots = 0;          // > 0: TC_IN_HALO (patch staged once, taps as shifted views)
    int epi_warps = 0;           // epilogue warps when the layer runs one CTA per SM (measured per layer class)
    // fused MBConv block starting at this (expand) op: expand -> depthwise -> pool -> FC -> FC -> projection = ops i .. i+5
    int mb_group = 0;            // channel group of the fused kernel (0: block runs layer by layer)
    void* mb_we_pack = nullptr;  // expand weights packed with N tile = mb_group
    void* mb_wp_pack = nullptr;  // projection weights packed with one N tile = cout
    float* mb_wpT = nullptr;     // projection weights [cout][cexp] FP32
};

struct RangeDev {   // dense per-class tri-state on the device
    uint8_t* state = nullptr;
    float* score = nullptr;
    uint64_t n = 0;
    int rerank = 0;
    int device = 0;
    ~RangeDev();
};

struct PostCfg {
    uint64_t top_k = 10;          // ClassifierBuilder default, src/classifier.rs:72
    int has_min_conf = 0;         // min_confidence: None,      src/classifier.rs:73
    float min_conf = 0.f;
    std::shared_ptr<RangeDev> range;
};

}  // namespace bn

struct bn_engine {
    int device = 0;
    int pack_threads = 0;
    bool tc_mode = true;          // tensor-core path with hi/lo-plane activations (BN_DISABLE_TC=1 -> FP32 CUDA-core path)
    int num_sms = 148;
    bool use_tma = true;          // TMA tile loads for 1x1 layers (BN_DISABLE_TMA=1 -> cp.async gather)
    bn::Plan plan;
    bn_io_info info{};
    std::vector<bn::DevOp> dev_ops;
    std::vector<float*> d_basis;   // per front-end branch [n_fft][ldb]
    std::vector<int> ldb;
    // tensor-core front-end: per branch the block-Toeplitz frame matrix geometry + packed basis
    struct FeTc { void* wpack = nullptr; int hop = 0, row_stride = 0, rows = 0, K = 0, nt = 0, n_tiles = 0, k_chunks = 0, stages = 2, tmem_cols = 32; };
    std::vector<FeTc> fe_tc;
    // fused v2.4 front-end (frontend_v24.cu): replaces the frame-matrix planes + per-branch GEMMs when the shapes fit
    struct FeV24 {
        bool on = false;
        void* wpack[2] = {nullptr, nullptr};
        int slot_branch[2] = {0, 0};      // kernel slot -> front-end branch index (slot 0 = longer K loop)
        int hop[2] = {0, 0}, kcells[2] = {0, 0}, rows[2] = {0, 0}, n_ksteps[2] = {0, 0}, blocks[2] = {0, 0}, split[2] = {0, 0}, pad[2] = {0, 0};
        int n_br = 0, n_pad = 0, row_pitch = 0, n_stages = 0; uint32_t patch_plane = 0, smem_bytes = 0;
    } fe_v24;
    // log-mel front-end (32 kHz graphs): window, float64-built twiddles, banded mel filters
    struct FeLogmel {
        float* window = nullptr; float* twiddle = nullptr; int* mel_lo = nullptr; int* mel_cnt = nullptr; int* mel_off = nullptr;
        float* mel_w = nullptr; int radix[8] = {0}; int n_stages = 0;
    } fe_lm;
    std::mutex post_mu;
    bn::PostCfg post;
    // bn_engine_run: bounded pool of internal contexts, checked out per call and returned (capi.cu)
    static constexpr int MAX_RUN_CTX = 4;
    std::mutex ctx_mu;
    std::condition_variable ctx_cv;
    std::vector<bn_ctx*> run_free;
    int run_created = 0;
    // Compute lanes of the engine: every context's kernels go to one of these streams (contexts alternate), one whole
    // batch at a time (launch_mu), so batches of concurrent callers run back to back instead of time-slicing the SMs
    // six ways; two lanes let one batch's tail / launch gaps / under-filled waves be covered by the other's kernels
    // (measured device-resident: 65.3 k -> 68.6 k segments/s).  H2D / D2H copies stay on their own streams and overlap
    // the lanes.  BN_COMPUTE_LANES=1|2; BN_SHARED_COMPUTE=0 gives every context a private compute stream.
    static constexpr int MAX_LANES = 4;
    int n_lanes = 2;
    cudaStream_t compute[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    std::mutex launch_mu[MAX_LANES];
    std::atomic<int> next_lane{0};
    cudaStream_t h2d = nullptr;     // page-locked caller memory is copied on this one stream, in kernel order
    std::mutex h2d_mu;
    ~bn_engine();
};

struct bn_ctx {
    bn_engine* eng = nullptr;
    int device = 0;                     // copy of eng->device: the destructor does not touch the engine
    std::vector<uint8_t> owns_tensor;   // per plan tensor: this context allocated it (not an alias / virtual tensor)
    uint64_t max_batch = 0;
    cudaStream_t stream = nullptr;      // compute: one of the engine's lanes (shared) or a private stream
    int lane = 0;
    bool owns_stream = true;
    cudaStream_t in_stream = nullptr;   // H2D staging of this context (== stream when the compute stream is private)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_in = nullptr;        // in_stream: this run's input is on the device
    cudaEvent_t done = nullptr;
    cudaEvent_t ev_results = nullptr;   // compute stream: logits / top-k of this run are final
    cudaEvent_t ev_fetched = nullptr;   // copy stream: the previous run's results have left the device buffers
    bool fetch_pending = false;
    std::vector<int8_t> f32_out_cache;   // per plan op: -1 unknown, 0 planes output, 1 FP32 rows (hand-off to the fused depthwise kernel)
    float* h_in = nullptr;       // pinned [max_batch][S]
    float* d_in = nullptr;       // [max_batch][S]
    int16_t* h_pcm = nullptr;    // pinned [max_batch * S] 16-bit PCM staging (bn_ctx_run_pcm16), allocated on first use
    int16_t* d_pcm = nullptr;
    float* d_norm = nullptr;     // [max_batch][S] normalised audio (v2.4 front-end)
    uint32_t* d_minmax = nullptr;   // [max_batch][2] order-preserving keys of the per-segment min / max
    bool keep_normalized = true;   // write the FP32 normalised audio (tests: BN_KEEP_NORMALIZED=1; always in FP32 mode)
    std::vector<__half*> d_xp;   // per branch: hi/lo planes of the frame matrix [max_batch][rows][row_stride]
    std::vector<float*> d_tensor;   // per plan tensor (aliases resolved to their root)
    std::vector<CUtensorMap> tmaps;   // per plan op: tensor map of its input planes (TC_IN_TMA layers), encoded on first use
    std::vector<uint8_t> tmap_state;  // 0 = not tried, 1 = ready, 2 = unavailable
    std::vector<CUtensorMap> omaps;   // per plan op: tensor map of its output (TMA-store epilogue)
    std::vector<uint8_t> omap_state;
    std::vector<CUtensorMap> mb_xmaps, mb_dmaps;   // per plan op (expand op of a fused MBConv block): block input / depthwise output
    std::vector<uint8_t> mb_state;                 // 0 = not tried, 1 = ready, 2 = unavailable
    float* h_logits = nullptr;   // pinned
    float* h_emb = nullptr;      // pinned
    bn::Pred* d_topk = nullptr;
    uint32_t* d_count = nullptr;
    bn::Pred* h_topk = nullptr;
    uint32_t* h_count = nullptr;
    uint64_t topk_cap = 0;       // slots per segment currently allocated
    std::shared_ptr<bn::RangeDev> range_in_flight;
    bool draining = false;
    uint64_t pending_batch = 0, pending_k = 0;
    uint64_t last_launches = 0;
    bool last_in_place = false;   // last bn_ctx_run copied straight from the caller's page-locked slices (no gather)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<std::string> prof_names;
    std::vector<float> prof_ms;
    std::vector<std::thread> packers;
    ~bn_ctx();
};

namespace bn {
int engine_create(const char* path, const bn_device_cfg* cfg, bn_engine** out);
int ctx_create(bn_engine* e, uint64_t max_batch, bn_ctx** out);
int ctx_run_host(bn_ctx* c, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
                 bool check_max_first, const bn_run_opts* opts, bn_outputs* out);
int ctx_run_device(bn_ctx* c, const float* d_audio, uint64_t batch, bool fetch, const bn_run_opts* opts,
                   bn_outputs* out);
int ctx_enqueue_device(bn_ctx* c, const float* d_audio, uint64_t batch, bool fetch, const bn_run_opts* opts);
int ctx_run_pcm16(bn_ctx* c, const int16_t* pcm, uint64_t n_samples, uint64_t first_pos, uint64_t step, uint64_t batch,
                  const bn_run_opts* opts, bn_outputs* out);
int ctx_wait(bn_ctx* c, const bn_run_opts* opts, bn_outputs* out);
int fill_io_info(const Plan& plan, bn_io_info* out);
}  // namespace bn
