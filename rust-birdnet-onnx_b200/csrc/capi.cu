// extern "C" surface declared in include/birdnet_b200.h.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "engine.h"

using namespace bn;

namespace {
struct DevBuf {                 // freed on every exit path of the *_apply helpers
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() { return static_cast<T*>(p); }
};
}  // namespace

extern "C" {

const char* bn_last_error(void) { return last_error().c_str(); }
void bn_last_error_detail(uint64_t out[3]) {
    const uint64_t* d = last_detail();
    out[0] = d[0]; out[1] = d[1]; out[2] = d[2];
}
const char* bn_version(void) { return "birdnet_b200 0.1.0 (sm_100a)"; }

void* bn_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void bn_host_free(void* p) { if (p) cudaFreeHost(p); }
int bn_host_register(void* p, uint64_t bytes) {
    if (!p || bytes == 0) return set_error(BN_ERR_INVALID_ARGUMENT, "bn_host_register: null pointer or empty range");
    const cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(BN_ERR_INVALID_ARGUMENT, std::string("bn_host_register: ") + cudaGetErrorString(e));
    }
    return BN_OK;
}
int bn_host_unregister(void* p) {
    if (!p) return set_error(BN_ERR_INVALID_ARGUMENT, "bn_host_unregister: null pointer");
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(BN_ERR_INVALID_ARGUMENT, std::string("bn_host_unregister: ") + cudaGetErrorString(e));
    }
    return BN_OK;
}

int bn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int bn_engine_create(const char* onnx_path, const bn_device_cfg* cfg, bn_engine** out) {
    return engine_create(onnx_path, cfg, out);
}
void bn_engine_destroy(bn_engine* engine) { delete engine; }

int bn_engine_io_info(const bn_engine* engine, bn_io_info* out) {
    if (!engine || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    *out = engine->info;
    return BN_OK;
}

int bn_model_inspect(const char* onnx_path, int32_t model_type_override, bn_io_info* out) {
    if (!out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    Plan plan;
    int st = load_plan(onnx_path, model_type_override, plan);
    if (st != BN_OK) return st;
    return fill_io_info(plan, out);
}

// Canonical text form of the layer plan a file is matched into: value names are left out, weights enter through a
// hash of their bits, so two files that describe the same network (whoever wrote them) give the same text.
static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h = 1469598103934665603ull) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
int bn_model_plan_summary(const char* onnx_path, int32_t model_type_override, char* buf, uint64_t cap, uint64_t* needed) {
    Plan plan;
    int st = load_plan(onnx_path, model_type_override, plan);
    if (st != BN_OK) return st;
    std::string out;
    char line[512];
    const FrontEndPlan& fe = plan.fe;
    snprintf(line, sizeof line, "frontend kind=%d samples=%d pad_end=%d normalize=%d eps=%.9g half=%.9g two=%.9g log_floor=%.9g log_scale=%.9g\n",
             fe.kind, fe.sample_count, fe.pad_end, (int)fe.normalize, fe.eps, fe.half, fe.two, fe.log_floor, fe.log_scale);
    out += line;
    for (auto& b : fe.branches) {
        snprintf(line, sizeof line, "  branch n_fft=%d hop=%d bins=%d mels=%d frames=%d exponent=%.9g flip=%d window=%016llx mel=%016llx\n",
                 b.n_fft, b.hop, b.n_bins, b.n_mels, b.n_frames, b.exponent, (int)b.flip,
                 (unsigned long long)fnv1a(b.window.data(), b.window.size() * 4), (unsigned long long)fnv1a(b.mel.data(), b.mel.size() * 4));
        out += line;
    }
    for (auto& op : plan.ops) {
        // weights hashed in a layout-independent order: the engine layout is [k*k*cin][ldw] with pad columns
        uint64_t hw = 1469598103934665603ull;
        if (op.kind == OP_DWCONV) hw = fnv1a(op.weight.data(), op.weight.size() * 4);
        else if (op.kind != OP_GAP) {
            const size_t rows = op.weight.size() / (size_t)std::max(op.ldw, 1);
            for (size_t r = 0; r < rows; ++r) hw = fnv1a(op.weight.data() + r * op.ldw, (size_t)op.cout * 4, hw);
        }
        snprintf(line, sizeof line, "op kind=%d k=%d s=%d p=%d cin=%d cout=%d in=%dx%d out=%dx%d act=%d gated=%d residual=%d w=%016llx b=%016llx\n",
                 op.kind, op.k, op.stride, op.pad, op.cin, op.cout, op.hin, op.win, op.hout, op.wout, op.act, (int)(op.in_scale >= 0),
                 (int)(op.residual >= 0), (unsigned long long)hw, (unsigned long long)fnv1a(op.bias.data(), op.bias.size() * 4));
        out += line;
    }
    snprintf(line, sizeof line, "outputs=%zu model_type=%d num_species=%d embedding_dim=%d\n", plan.outputs.size(), plan.model_type, plan.num_species, plan.embedding_dim);
    out += line;
    if (needed) *needed = out.size() + 1;
    if (buf && cap) {
        const size_t n = std::min<size_t>(out.size(), (size_t)cap - 1);
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return BN_OK;
}

int bn_detect_model_type(const int64_t* input_dims, int32_t input_rank, const int64_t* output_dims,
                         const int32_t* output_ranks, int32_t n_outputs, int32_t model_type_override,
                         bn_io_info* out) {
    if (!out || (input_rank > 0 && !input_dims)) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (n_outputs < 0 || (n_outputs > 0 && (!output_ranks || !output_dims))) return set_error(BN_ERR_INVALID_ARGUMENT, "null output shape arrays");
    for (int i = 0; i < n_outputs; ++i)
        if (output_ranks[i] < 0 || output_ranks[i] > BN_MAX_DIMS) return set_error(BN_ERR_INVALID_ARGUMENT, "output rank out of range");
    std::vector<int64_t> in(input_dims, input_dims + std::max(0, input_rank));
    std::vector<std::vector<int64_t>> outs;
    const int64_t* p = output_dims;
    for (int i = 0; i < n_outputs; ++i) {
        outs.emplace_back(p, p + output_ranks[i]);
        p += output_ranks[i];
    }
    int mt = -1, sc = 0, ns = 0, ed = 0;
    std::string reason;
    if (!detect_model_type(in, outs, model_type_override, &mt, &sc, &ns, &ed, &reason))
        return set_error(BN_ERR_MODEL_DETECTION, reason);
    memset(out, 0, sizeof(*out));
    out->model_type = mt;
    out->sample_rate = mt == MT_BIRDNET_V24 ? 48000u : 32000u;
    out->segment_duration = mt == MT_BIRDNET_V24 ? 3.0f : 5.0f;
    out->sample_count = (uint64_t)sc;
    out->num_species = (uint64_t)ns;
    out->embedding_dim = (uint64_t)ed;
    out->n_outputs = n_outputs;
    return BN_OK;
}

int bn_engine_set_postprocess(bn_engine* engine, uint64_t top_k, int32_t has_min_confidence, float min_confidence) {
    if (!engine) return set_error(BN_ERR_INVALID_ARGUMENT, "null engine");
    std::lock_guard<std::mutex> lk(engine->post_mu);
    engine->post.top_k = top_k;
    engine->post.has_min_conf = has_min_confidence ? 1 : 0;
    engine->post.min_conf = min_confidence;
    return BN_OK;
}

static int make_range(int device, const uint8_t* state, const float* score, uint64_t n, int rerank,
                      std::shared_ptr<RangeDev>& out) {
    if (!state || !score || n == 0) return set_error(BN_ERR_INVALID_ARGUMENT, "range filter needs state and score arrays");
    for (uint64_t i = 0; i < n; ++i)
        if (state[i] > 2) return set_error(BN_ERR_INVALID_ARGUMENT, "range state must be 0, 1 or 2");
    BN_CUDA(cudaSetDevice(device));
    auto r = std::make_shared<RangeDev>();
    r->device = device;
    r->n = n;
    r->rerank = rerank ? 1 : 0;
    BN_CUDA(cudaMalloc(&r->state, n));
    BN_CUDA(cudaMalloc(&r->score, n * sizeof(float)));
    BN_CUDA(cudaMemcpy(r->state, state, n, cudaMemcpyHostToDevice));
    BN_CUDA(cudaMemcpy(r->score, score, n * sizeof(float), cudaMemcpyHostToDevice));
    out = r;
    return BN_OK;
}

int bn_engine_set_range_filter(bn_engine* engine, const uint8_t* state, const float* score, uint64_t n, int32_t rerank) {
    if (!engine) return set_error(BN_ERR_INVALID_ARGUMENT, "null engine");
    if (n != engine->info.num_species)
        return set_error(BN_ERR_INVALID_ARGUMENT, "range filter length " + std::to_string(n) + " != num_species " + std::to_string(engine->info.num_species));
    std::shared_ptr<RangeDev> r;
    int st = make_range(engine->device, state, score, n, rerank, r);
    if (st != BN_OK) return st;
    std::lock_guard<std::mutex> lk(engine->post_mu);
    engine->post.range = r;
    return BN_OK;
}

int bn_engine_clear_range_filter(bn_engine* engine) {
    if (!engine) return set_error(BN_ERR_INVALID_ARGUMENT, "null engine");
    std::lock_guard<std::mutex> lk(engine->post_mu);
    engine->post.range.reset();
    return BN_OK;
}

// bn_engine_run (predict / predict_batch, classifier.rs:676-727): a small bounded pool of internal contexts that
// are checked out per call and returned, and a chunk loop, so that
//   * a call of any size works in bounded device memory (the reference's predict_batch takes any batch),
//   * services that churn threads do not accumulate one context per thread id ever seen,
//   * a context is only replaced by a larger one AFTER the larger one exists.
// Results of all chunks are gathered in a per-thread host buffer that stays valid until the same thread calls
// bn_engine_run again (the pooled context itself goes straight back to the pool).
namespace {
constexpr uint64_t RUN_CHUNK = 256;       // segments per internal context (about 3 GB of activations for v2.4)

struct HostResults {
    std::vector<float> logits, emb;
    std::vector<bn_pred> topk;
    std::vector<uint32_t> count;
};
thread_local HostResults t_results;

struct CtxLease {
    bn_engine* e;
    bn_ctx* c = nullptr;
    explicit CtxLease(bn_engine* eng) : e(eng) {}
    ~CtxLease() {
        if (!c) return;
        { std::lock_guard<std::mutex> lk(e->ctx_mu); e->run_free.push_back(c); }
        e->ctx_cv.notify_one();
    }
};

int checkout_ctx(bn_engine* e, uint64_t need, CtxLease& lease) {
    uint64_t cap = 1;
    while (cap < need) cap <<= 1;
    std::unique_lock<std::mutex> lk(e->ctx_mu);
    while (true) {
        int best = -1;                       // smallest free context that is large enough
        for (size_t i = 0; i < e->run_free.size(); ++i)
            if (e->run_free[i]->max_batch >= need && (best < 0 || e->run_free[i]->max_batch < e->run_free[best]->max_batch)) best = (int)i;
        if (best >= 0) {
            lease.c = e->run_free[best];
            e->run_free.erase(e->run_free.begin() + best);
            return BN_OK;
        }
        if (e->run_created < bn_engine::MAX_RUN_CTX || !e->run_free.empty()) {
            // room for one more, or a too-small free one to replace: build the new context first
            bn_ctx* victim = nullptr;
            if (e->run_created >= bn_engine::MAX_RUN_CTX) { victim = e->run_free.back(); e->run_free.pop_back(); }
            else ++e->run_created;
            lk.unlock();
            bn_ctx* c = nullptr;
            int st = ctx_create(e, cap, &c);
            if (st != BN_OK && victim && victim->max_batch < cap) {
                // out of memory with the small one still alive: free it and try once more
                delete victim;
                victim = nullptr;
                st = ctx_create(e, cap, &c);
                if (st != BN_OK) { lk.lock(); --e->run_created; return st; }
            } else if (st != BN_OK) {
                lk.lock();
                if (victim) e->run_free.push_back(victim); else --e->run_created;
                return st;
            }
            delete victim;
            lease.c = c;
            return BN_OK;
        }
        e->ctx_cv.wait(lk);                 // every internal context is busy: wait for one to come back
    }
}
}  // namespace

int bn_engine_run(bn_engine* engine, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
                  const bn_run_opts* opts, bn_outputs* out) {
    if (!engine || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) { memset(out, 0, sizeof(*out)); return BN_OK; }
    if (!seg_ptrs || !seg_lens) return set_error(BN_ERR_INVALID_ARGUMENT, "null segment array");
    // validate before any allocation (classifier.rs:688-696)
    const uint64_t S = engine->info.sample_count;
    for (uint64_t i = 0; i < batch; ++i)
        if (seg_lens[i] != S) return set_error_detail(BN_ERR_BATCH_INPUT_SIZE, "batch input size mismatch", i, S, seg_lens[i]);
    CtxLease lease(engine);
    int st = checkout_ctx(engine, std::min(batch, RUN_CHUNK), lease);
    if (st != BN_OK) return st;
    bn_ctx* c = lease.c;
    HostResults& hr = t_results;
    const uint64_t N = engine->info.num_species, E = engine->info.embedding_dim;
    uint64_t k_stride = 0;
    bool sized = false;
    for (uint64_t lo = 0; lo < batch; lo += c->max_batch) {
        const uint64_t nb = std::min<uint64_t>(c->max_batch, batch - lo);
        bn_outputs part;
        st = ctx_run_host(c, seg_ptrs + lo, seg_lens + lo, nb, false, opts, &part);
        if (st != BN_OK) return st;
        if (!sized) {                        // the first chunk fixes the top-k stride of the whole call
            k_stride = part.topk_stride;
            hr.logits.resize(batch * N);
            hr.emb.resize(part.embeddings ? batch * E : 0);
            hr.topk.resize(batch * k_stride);
            hr.count.resize(batch);
            sized = true;
        }
        memcpy(hr.logits.data() + lo * N, part.logits, nb * N * sizeof(float));
        if (part.embeddings) memcpy(hr.emb.data() + lo * E, part.embeddings, nb * E * sizeof(float));
        if (k_stride) memcpy(hr.topk.data() + lo * k_stride, part.topk, nb * k_stride * sizeof(bn_pred));
        memcpy(hr.count.data() + lo, part.topk_count, nb * sizeof(uint32_t));
    }
    memset(out, 0, sizeof(*out));
    out->batch = batch;
    out->num_species = N;
    out->logits = hr.logits.data();
    out->embedding_dim = E;
    out->embeddings = hr.emb.empty() ? nullptr : hr.emb.data();
    out->topk_stride = k_stride;
    out->topk_count = hr.count.data();
    out->topk = hr.topk.data();
    return BN_OK;
}

int bn_ctx_create(bn_engine* engine, uint64_t max_batch_size, bn_ctx** out) {
    if (!engine || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (engine->info.model_type == BN_MODEL_PERCH_V2)      // batch_context.rs:107-114
        return set_error(BN_ERR_INFERENCE, "BatchInferenceContext does not yet support PerchV2 models. Use predict_batch() instead.");
    return ctx_create(engine, max_batch_size, out);
}
int bn_ctx_create_ex(bn_engine* engine, uint64_t max_batch_size, uint32_t flags, bn_ctx** out) {
    if (!engine || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (flags & ~(uint32_t)BN_CTX_ALLOW_PERCH) return set_error(BN_ERR_INVALID_ARGUMENT, "unknown context flag");
    if (!(flags & BN_CTX_ALLOW_PERCH)) return bn_ctx_create(engine, max_batch_size, out);
    return ctx_create(engine, max_batch_size, out);
}
void bn_ctx_destroy(bn_ctx* ctx) { delete ctx; }

int bn_ctx_run(bn_ctx* ctx, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
               const bn_run_opts* opts, bn_outputs* out) {
    return ctx_run_host(ctx, seg_ptrs, seg_lens, batch, true, opts, out);
}
uint64_t bn_ctx_max_batch_size(const bn_ctx* ctx) { return ctx ? ctx->max_batch : 0; }
uint64_t bn_ctx_input_buffer_bytes(const bn_ctx* ctx) {
    return ctx ? ctx->max_batch * (uint64_t)ctx->eng->plan.sample_count * sizeof(float) : 0;
}

int bn_ctx_run_device(bn_ctx* ctx, const float* d_audio, uint64_t batch, int32_t fetch_outputs,
                      const bn_run_opts* opts, bn_outputs* out) {
    if (!d_audio && batch) return set_error(BN_ERR_INVALID_ARGUMENT, "null device buffer");
    return ctx_run_device(ctx, d_audio, batch, fetch_outputs != 0, opts, out);
}

int bn_ctx_run_pcm16(bn_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t first_pos, uint64_t step, uint64_t batch,
                     const bn_run_opts* opts, bn_outputs* out) {
    return ctx_run_pcm16(ctx, pcm, n_samples, first_pos, step, batch, opts, out);
}

int bn_ctx_enqueue_device(bn_ctx* ctx, const float* d_audio, uint64_t batch, int32_t fetch_outputs) {
    if (!d_audio && batch) return set_error(BN_ERR_INVALID_ARGUMENT, "null device buffer");
    return ctx_enqueue_device(ctx, d_audio, batch, fetch_outputs != 0, nullptr);
}
int bn_ctx_wait(bn_ctx* ctx, const bn_run_opts* opts, bn_outputs* out) { return ctx_wait(ctx, opts, out); }

// development aid (not in the public header): role cycle counters of the tensor-core conv launches
extern "C" int bn_debug_tc_profile(unsigned long long* out, int slots) {
    cudaError_t e = bn::tc_conv_prof_read(out, slots);
    return e == cudaSuccess ? BN_OK : bn::cuda_fail(e, "tc_conv_prof_read");
}

int bn_ctx_read_tensor(bn_ctx* ctx, const char* name, float* dst, uint64_t dst_elems, uint64_t* elems_out) {
    if (!ctx || !name) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    const Plan& p = ctx->eng->plan;
    for (size_t i = 0; i < p.tensors.size(); ++i) {
        if (p.tensors[i].name != name) continue;
        if (p.tensors[i].scale_base >= 0) return set_error(BN_ERR_INVALID_ARGUMENT, "tensor is virtual (gated)");
        uint64_t per = p.tensors[i].elems();
        if (elems_out) *elems_out = per;
        if (!dst) return BN_OK;
        uint64_t n = std::min<uint64_t>(dst_elems, per * ctx->max_batch);
        BN_CUDA(cudaSetDevice(ctx->eng->device));
        BN_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->eng->tc_mode && p.tensors[i].H * p.tensors[i].W > 1) {
            // hi/lo fp16 planes -> FP32 (x = hi + lo)
            const size_t plane = (size_t)std::max<uint64_t>(ctx->max_batch, 1) * per;
            std::vector<__half> hi(n), lo(n);
            const __half* base = reinterpret_cast<const __half*>(ctx->d_tensor[i]);
            BN_CUDA(cudaMemcpy(hi.data(), base, n * sizeof(__half), cudaMemcpyDeviceToHost));
            BN_CUDA(cudaMemcpy(lo.data(), base + plane, n * sizeof(__half), cudaMemcpyDeviceToHost));
            for (uint64_t j = 0; j < n; ++j) dst[j] = __half2float(hi[j]) + __half2float(lo[j]);
            return BN_OK;
        }
        BN_CUDA(cudaMemcpy(dst, ctx->d_tensor[i], n * sizeof(float), cudaMemcpyDeviceToHost));
        return BN_OK;
    }
    return set_error(BN_ERR_INVALID_ARGUMENT, std::string("no tensor named '") + name + "'");
}

int bn_ctx_read_normalized(bn_ctx* ctx, float* dst, uint64_t dst_elems) {
    if (!ctx || !dst) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (!ctx->d_norm) return set_error(BN_ERR_INVALID_ARGUMENT, "model has no normaliser");
    if (!ctx->keep_normalized) return set_error(BN_ERR_INVALID_ARGUMENT, "normalised audio is not kept: set BN_KEEP_NORMALIZED=1 before creating the context");
    uint64_t n = std::min<uint64_t>(dst_elems, ctx->max_batch * (uint64_t)ctx->eng->plan.sample_count);
    BN_CUDA(cudaSetDevice(ctx->eng->device));
    BN_CUDA(cudaStreamSynchronize(ctx->stream));
    BN_CUDA(cudaMemcpy(dst, ctx->d_norm, n * sizeof(float), cudaMemcpyDeviceToHost));
    return BN_OK;
}

uint64_t bn_ctx_last_launch_count(const bn_ctx* ctx) { return ctx ? ctx->last_launches : 0; }
int bn_ctx_last_run_in_place(const bn_ctx* ctx) { return ctx && ctx->last_in_place ? 1 : 0; }
uint64_t bn_ctx_nonfinite_segments(const bn_ctx* ctx) {
    return ctx && ctx->h_count ? ctx->h_count[std::max<uint64_t>(ctx->max_batch, 1)] : 0;
}
int bn_ctx_set_profiling(bn_ctx* ctx, int32_t enabled) {
    if (!ctx) return set_error(BN_ERR_INVALID_ARGUMENT, "null ctx");
    ctx->profiling = enabled != 0;
    return BN_OK;
}
int bn_ctx_stage_times(const bn_ctx* ctx, float* ms_out, char (*names_out)[48], uint64_t cap, uint64_t* n_out) {
    if (!ctx || !n_out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    uint64_t n = ctx->prof_ms.size();
    *n_out = n;
    for (uint64_t i = 0; i < std::min(n, cap); ++i) {
        if (ms_out) ms_out[i] = ctx->prof_ms[i];
        if (names_out) { memset(names_out[i], 0, 48); strncpy(names_out[i], ctx->prof_names[i].c_str(), 47); }
    }
    return BN_OK;
}
void* bn_ctx_stream(bn_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int32_t bn_engine_compute_lanes(const bn_engine* engine) { return engine ? engine->n_lanes : 0; }


int bn_range_filter_apply(bn_engine* engine, const bn_pred* in, const uint32_t* in_count, uint64_t rows,
                          uint64_t stride, const uint8_t* state, const float* score, uint64_t n, int32_t rerank,
                          bn_pred* out, uint32_t* out_count) {
    if (!out || !out_count) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    const int device = engine ? engine->device : 0;      // NULL engine: device 0
    if (rows == 0) return BN_OK;
    if (stride == 0) { memset(out_count, 0, rows * sizeof(uint32_t)); return BN_OK; }
    if (!in || !in_count) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    BN_CUDA(cudaSetDevice(device));
    BN_CUDA(init_kernels_for_device());
    std::shared_ptr<RangeDev> r;
    if (state) { int st = make_range(device, state, score, n, rerank, r); if (st != BN_OK) return st; }
    DevBuf b_in, b_out, b_ic, b_oc;
    size_t bytes = rows * stride * sizeof(Pred);
    BN_CUDA(b_in.alloc(bytes));
    BN_CUDA(b_out.alloc(bytes));
    BN_CUDA(b_ic.alloc(rows * sizeof(uint32_t)));
    BN_CUDA(b_oc.alloc(rows * sizeof(uint32_t)));
    Pred *d_in = b_in.as<Pred>(), *d_out = b_out.as<Pred>();
    uint32_t *d_ic = b_ic.as<uint32_t>(), *d_oc = b_oc.as<uint32_t>();
    cudaError_t ce = cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_ic, in_count, rows * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = launch_range_filter((const Pred*)d_in, d_ic, (int)rows, (int)stride, r ? r->state : nullptr, r ? r->score : nullptr,
                                                    (int)n, rerank ? 1 : 0, d_out, d_oc, 0);
    if (ce == cudaSuccess) ce = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(out_count, d_oc, rows * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return cuda_fail(ce, "bn_range_filter_apply");
    return BN_OK;
}

int bn_topk_apply(bn_engine* engine, const float* logits, uint64_t rows, uint64_t n, uint64_t top_k,
                  int32_t has_min_confidence, float min_confidence, const uint8_t* state, const float* score,
                  int32_t rerank, bn_pred* out, uint32_t* out_count) {
    if (!out_count) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    const int device = engine ? engine->device : 0;      // NULL engine: device 0
    if (rows == 0) return BN_OK;
    uint64_t k = std::min(top_k, n);                      // postprocess.rs:46-50
    if (k == 0 || n == 0) { memset(out_count, 0, rows * sizeof(uint32_t)); return BN_OK; }
    if (!logits || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    BN_CUDA(cudaSetDevice(device));
    BN_CUDA(init_kernels_for_device());
    std::shared_ptr<RangeDev> r;
    if (state) { int st = make_range(device, state, score, n, rerank, r); if (st != BN_OK) return st; }
    DevBuf b_l, b_out, b_oc;
    BN_CUDA(b_l.alloc(rows * n * sizeof(float)));
    BN_CUDA(b_out.alloc(rows * k * sizeof(Pred)));
    BN_CUDA(b_oc.alloc(rows * sizeof(uint32_t)));
    float* d_l = b_l.as<float>();
    Pred* d_out = b_out.as<Pred>();
    uint32_t* d_oc = b_oc.as<uint32_t>();
    cudaError_t ce = cudaMemcpy(d_l, logits, rows * n * sizeof(float), cudaMemcpyHostToDevice);
    TopkParams tp{};
    tp.logits = d_l; tp.batch = (int)rows; tp.n = (int)n; tp.k = (uint32_t)k;
    tp.has_min_conf = has_min_confidence ? 1 : 0; tp.min_conf = min_confidence;
    tp.range_state = r ? r->state : nullptr; tp.range_score = r ? r->score : nullptr; tp.rerank = rerank ? 1 : 0;
    tp.out = d_out; tp.out_count = d_oc;
    if (ce == cudaSuccess) ce = launch_topk(tp, 0);
    if (ce == cudaSuccess) ce = cudaMemcpy(out, d_out, rows * k * sizeof(Pred), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(out_count, d_oc, rows * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return cuda_fail(ce, "bn_topk_apply");
    return BN_OK;
}

}  // extern "C"
