// Engine: ONNX -> plan -> device weights; per-context staging slabs, streams and the
// enqueue of the whole hot path (H2D, front-end, CNN, epilogue, D2H).
//
// Reference call sites replaced: src/classifier.rs:340-383 (build), 504-574 (run_inference
// monitor), 676-727 (predict_batch), 826-867 (predict_batch_with_context);
// src/batch_context.rs:102-133, 188-338.
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>

namespace bn {

// --------------------------------------------------------------------------------------
// error channel
// --------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local uint64_t g_detail[3] = {0, 0, 0};

int set_error(int status, const std::string& msg) {
    g_err = msg;
    g_detail[0] = g_detail[1] = g_detail[2] = 0;
    return status;
}
int set_error_detail(int status, const std::string& msg, uint64_t a, uint64_t b, uint64_t c) {
    g_err = msg;
    g_detail[0] = a; g_detail[1] = b; g_detail[2] = c;
    return status;
}
const std::string& last_error() { return g_err; }
const uint64_t* last_detail() { return g_detail; }

int cuda_fail(cudaError_t e, const char* what) {
    std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " (" + what + ")";
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice ||
        e == cudaErrorInitializationError)
        return set_error(BN_ERR_RUNTIME_INIT, msg);
    return set_error(BN_ERR_INFERENCE, msg);
}

RangeDev::~RangeDev() {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    if (state) cudaFree(state);
    if (score) cudaFree(score);
    cudaSetDevice(prev);
}

// --------------------------------------------------------------------------------------
// io info / detection glue
// --------------------------------------------------------------------------------------
static void copy_tensor_info(const std::string& name, const std::vector<int64_t>& dims, bn_tensor_info* ti) {
    memset(ti, 0, sizeof(*ti));
    strncpy(ti->name, name.c_str(), sizeof(ti->name) - 1);
    ti->rank = (int32_t)std::min<size_t>(dims.size(), BN_MAX_DIMS);
    for (int i = 0; i < ti->rank; ++i) ti->dims[i] = dims[i];
}

int fill_io_info(const Plan& plan, bn_io_info* out) {
    memset(out, 0, sizeof(*out));
    copy_tensor_info(plan.input_name, plan.input_dims, &out->input);
    out->n_outputs = (int32_t)std::min<size_t>(plan.outputs.size(), BN_MAX_OUTPUTS);
    for (int i = 0; i < out->n_outputs; ++i) copy_tensor_info(plan.outputs[i].name, plan.outputs[i].dims, &out->outputs[i]);
    out->model_type = plan.model_type;
    out->sample_rate = plan.model_type == MT_BIRDNET_V24 ? 48000u : 32000u;         // types.rs:17-22
    out->segment_duration = plan.model_type == MT_BIRDNET_V24 ? 3.0f : 5.0f;        // types.rs:26-31
    out->sample_count = (uint64_t)plan.sample_count;
    out->num_species = (uint64_t)plan.num_species;
    out->embedding_dim = (uint64_t)plan.embedding_dim;
    return BN_OK;
}

static int finish_plan(Plan& plan, int override_type) {
    std::vector<std::vector<int64_t>> oshapes;
    for (auto& o : plan.outputs) oshapes.push_back(o.dims);
    std::string reason;
    if (!detect_model_type(plan.input_dims, oshapes, override_type, &plan.model_type, &plan.sample_count,
                           &plan.num_species, &plan.embedding_dim, &reason))
        return set_error(BN_ERR_MODEL_DETECTION, reason);
    // output index map of src/classifier.rs:917-934
    int li = plan.model_type == MT_BIRDNET_V24 ? 0 : plan.model_type == MT_BIRDNET_V30 ? 1 : 3;
    plan.logits_tensor = plan.outputs[li].tensor;
    plan.embedding_tensor = plan.model_type == MT_BIRDNET_V24 ? -1 : plan.outputs[0].tensor;
    const TensorInfo& lt = plan.tensors[plan.logits_tensor];
    if ((int)lt.elems() != plan.num_species)
        return set_error(BN_ERR_MODEL_LOAD, "declared logits shape does not match the graph");
    if (plan.embedding_tensor >= 0 && (int)plan.tensors[plan.embedding_tensor].elems() != plan.embedding_dim)
        return set_error(BN_ERR_MODEL_LOAD, "declared embedding shape does not match the graph");
    if (plan.sample_count != plan.fe.sample_count)
        return set_error(BN_ERR_MODEL_LOAD, "input sample count does not match the front-end");
    return BN_OK;
}

int load_plan(const char* path, int override_type, Plan& plan) {
    if (!path) return set_error(BN_ERR_MODEL_PATH_REQUIRED, "model path required");
    try {
        OnnxModel m;
        load_onnx(path, m);
        build_plan(m, plan);
    } catch (const std::exception& ex) {
        return set_error(BN_ERR_MODEL_LOAD, ex.what());
    }
    return finish_plan(plan, override_type);
}

// --------------------------------------------------------------------------------------
// front-end basis: window[n] * cos(2*pi*f*n/N) projected on the mel matrix (float64 maths)
// --------------------------------------------------------------------------------------
static void build_real_mel_basis(const SpecBranch& br, std::vector<float>& basis, int& ldb) {
    const int N = br.n_fft, F = br.n_bins, Mm = br.n_mels;
    ldb = (Mm + 3) / 4 * 4;
    basis.assign((size_t)N * ldb, 0.f);
    std::vector<double> costab(N);
    for (int i = 0; i < N; ++i) costab[i] = cos(2.0 * M_PI * (double)i / (double)N);
    // sparse mel rows
    std::vector<std::vector<std::pair<int, double>>> nz(F);
    for (int f = 0; f < F; ++f)
        for (int m = 0; m < Mm; ++m) {
            float w = br.mel[(size_t)f * Mm + m];
            if (w != 0.f) nz[f].push_back({br.flip ? Mm - 1 - m : m, (double)w});
        }
    std::vector<double> acc(Mm);
    for (int n = 0; n < N; ++n) {
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int f = 0; f < F; ++f) {
            if (nz[f].empty()) continue;
            double c = costab[(int)(((long long)f * n) % N)];
            for (auto& pr : nz[f]) acc[pr.first] += c * pr.second;
        }
        for (int m = 0; m < Mm; ++m) basis[(size_t)n * ldb + m] = (float)(acc[m] * (double)br.window[n]);
    }
}

// --------------------------------------------------------------------------------------
// engine
// --------------------------------------------------------------------------------------
int engine_create(const char* path, const bn_device_cfg* cfg, bn_engine** out) {
    if (!out) return set_error(BN_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    std::unique_ptr<bn_engine> e(new bn_engine());
    e->device = cfg ? cfg->device_id : 0;
    int ov = cfg ? cfg->model_type_override : BN_MODEL_AUTO;
    int st = load_plan(path, ov, e->plan);
    if (st != BN_OK) return st;
    fill_io_info(e->plan, &e->info);
    unsigned hw = std::thread::hardware_concurrency();
    e->pack_threads = cfg && cfg->pack_threads > 0 ? cfg->pack_threads : (int)std::max(1u, std::min(16u, hw / 2));

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_error(BN_ERR_RUNTIME_INIT, std::string("no CUDA device available: ") + cudaGetErrorString(ce));
    if (e->device < 0 || e->device >= ndev)
        return set_error(BN_ERR_RUNTIME_INIT, "device_id " + std::to_string(e->device) + " out of range (" + std::to_string(ndev) + " devices)");
    BN_CUDA(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    BN_CUDA(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major < 10)
        return set_error(BN_ERR_RUNTIME_INIT, std::string("device '") + prop.name + "' is not sm_100-class; this engine is built for sm_100a only");
    BN_CUDA(init_kernels_for_device());

    Plan& p = e->plan;
    e->dev_ops.resize(p.ops.size());
    for (size_t i = 0; i < p.ops.size(); ++i) {
        PlanOp& op = p.ops[i];
        if (op.kind == OP_GAP) continue;
        BN_CUDA(cudaMalloc(&e->dev_ops[i].weight, op.weight.size() * sizeof(float)));
        BN_CUDA(cudaMemcpy(e->dev_ops[i].weight, op.weight.data(), op.weight.size() * sizeof(float), cudaMemcpyHostToDevice));
        BN_CUDA(cudaMalloc(&e->dev_ops[i].bias, op.bias.size() * sizeof(float)));
        BN_CUDA(cudaMemcpy(e->dev_ops[i].bias, op.bias.data(), op.bias.size() * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<float>().swap(op.weight);   // host copy no longer needed
    }
    if (p.fe.kind == FE_BIRDNET_V24) {
        for (auto& br : p.fe.branches) {
            std::vector<float> basis;
            int ldb = 0;
            build_real_mel_basis(br, basis, ldb);
            float* d = nullptr;
            BN_CUDA(cudaMalloc(&d, basis.size() * sizeof(float)));
            BN_CUDA(cudaMemcpy(d, basis.data(), basis.size() * sizeof(float), cudaMemcpyHostToDevice));
            e->d_basis.push_back(d);
            e->ldb.push_back(ldb);
        }
    } else {
        return set_error(BN_ERR_MODEL_LOAD, "log-mel front-end is not implemented in this engine build yet");
    }
    *out = e.release();
    return BN_OK;
}

}  // namespace bn

bn_engine::~bn_engine() {
    cudaSetDevice(device);
    for (auto& kv : thread_ctx) delete kv.second;
    for (auto& d : dev_ops) {
        if (d.weight) cudaFree(d.weight);
        if (d.bias) cudaFree(d.bias);
    }
    for (auto* b : d_basis) cudaFree(b);
}

bn_ctx::~bn_ctx() {
    if (!eng) return;
    cudaSetDevice(eng->device);
    if (stream) cudaStreamSynchronize(stream);
    if (h_in) cudaFreeHost(h_in);
    if (d_in) cudaFree(d_in);
    if (d_norm) cudaFree(d_norm);
    for (size_t i = 0; i < d_tensor.size(); ++i)
        if (d_tensor[i] && eng->plan.tensors[i].alias_of < 0 && eng->plan.tensors[i].scale_base < 0) cudaFree(d_tensor[i]);
    if (h_logits) cudaFreeHost(h_logits);
    if (h_emb) cudaFreeHost(h_emb);
    if (d_topk) cudaFree(d_topk);
    if (d_count) cudaFree(d_count);
    if (h_topk) cudaFreeHost(h_topk);
    if (h_count) cudaFreeHost(h_count);
    for (auto ev : prof_events) cudaEventDestroy(ev);
    if (done) cudaEventDestroy(done);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
}

namespace bn {

int ctx_create(bn_engine* e, uint64_t max_batch, bn_ctx** out) {
    if (!e || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    BN_CUDA(cudaSetDevice(e->device));
    std::unique_ptr<bn_ctx> c(new bn_ctx());
    c->eng = e;
    c->max_batch = max_batch;
    const Plan& p = e->plan;
    const size_t S = (size_t)p.sample_count;
    const size_t mb = std::max<uint64_t>(max_batch, 1);
    BN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    BN_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    BN_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    BN_CUDA(cudaHostAlloc(&c->h_in, mb * S * sizeof(float), cudaHostAllocDefault));
    memset(c->h_in, 0, mb * S * sizeof(float));    // vec![0.0f32; max*sample_count], batch_context.rs:122
    BN_CUDA(cudaMalloc(&c->d_in, mb * S * sizeof(float)));
    if (p.fe.normalize) BN_CUDA(cudaMalloc(&c->d_norm, mb * S * sizeof(float)));
    c->d_tensor.assign(p.tensors.size(), nullptr);
    for (size_t i = 0; i < p.tensors.size(); ++i) {
        const TensorInfo& t = p.tensors[i];
        if (t.alias_of >= 0 || t.scale_base >= 0) continue;
        BN_CUDA(cudaMalloc(&c->d_tensor[i], mb * t.elems() * sizeof(float)));
    }
    for (size_t i = 0; i < p.tensors.size(); ++i)
        if (p.tensors[i].alias_of >= 0) c->d_tensor[i] = c->d_tensor[p.root((int)i)];
    BN_CUDA(cudaHostAlloc(&c->h_logits, mb * p.num_species * sizeof(float), cudaHostAllocDefault));
    if (p.embedding_dim > 0) BN_CUDA(cudaHostAlloc(&c->h_emb, mb * p.embedding_dim * sizeof(float), cudaHostAllocDefault));
    BN_CUDA(cudaMalloc(&c->d_count, mb * sizeof(uint32_t)));
    BN_CUDA(cudaHostAlloc(&c->h_count, mb * sizeof(uint32_t), cudaHostAllocDefault));
    *out = c.release();
    return BN_OK;
}

static int ensure_topk_capacity(bn_ctx* c, uint64_t k) {
    if (k <= c->topk_cap && c->d_topk) return BN_OK;
    const size_t mb = std::max<uint64_t>(c->max_batch, 1);
    if (c->d_topk) { cudaFree(c->d_topk); c->d_topk = nullptr; }
    if (c->h_topk) { cudaFreeHost(c->h_topk); c->h_topk = nullptr; }
    uint64_t cap = std::max<uint64_t>(k, 1);
    BN_CUDA(cudaMalloc(&c->d_topk, mb * cap * sizeof(Pred)));
    BN_CUDA(cudaHostAlloc(&c->h_topk, mb * cap * sizeof(Pred), cudaHostAllocDefault));
    c->topk_cap = cap;
    return BN_OK;
}

// ---- profiling helpers -------------------------------------------------------------------
static void prof_mark(bn_ctx* c, const char* name) {
    if (!c->profiling) return;
    size_t i = c->prof_names.size();
    if (i >= c->prof_events.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        c->prof_events.push_back(ev);
    }
    cudaEventRecord(c->prof_events[i], c->stream);
    c->prof_names.push_back(name);
}

// ---- the forward pass (device side), input already in c->d_in or `d_audio` -----------------
static int enqueue_forward(bn_ctx* c, const float* d_audio, int B, const PostCfg& post, uint64_t k_eff) {
    bn_engine* e = c->eng;
    const Plan& p = e->plan;
    cudaStream_t s = c->stream;
    uint64_t launches = 0;
    const float* fe_in = d_audio;
    prof_mark(c, "normalize");
    if (p.fe.normalize) {
        BN_CUDA(launch_minmax_normalize(d_audio, c->d_norm, B, p.sample_count, p.fe.eps, p.fe.half, p.fe.two, s));
        ++launches;
        fe_in = c->d_norm;
    }
    float* spec = c->d_tensor[p.fe.out_tensor];
    for (size_t bi = 0; bi < p.fe.branches.size(); ++bi) {
        const SpecBranch& br = p.fe.branches[bi];
        prof_mark(c, bi == 0 ? "spectrogram0" : "spectrogram1");
        BN_CUDA(launch_spectrogram_v24(fe_in, e->d_basis[bi], e->ldb[bi], spec, B, p.sample_count, br.n_fft, br.hop,
                                       br.n_frames, br.n_mels, (int)p.fe.branches.size(), (int)bi, br.exponent, s));
        ++launches;
    }
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const PlanOp& op = p.ops[i];
        prof_mark(c, op.name.c_str());
        if (op.kind == OP_GAP) {
            BN_CUDA(launch_gap(c->d_tensor[op.in], c->d_tensor[op.out], B, op.hin * op.win, op.cin, s));
        } else if (op.kind == OP_DWCONV) {
            DwParams d{c->d_tensor[op.in], e->dev_ops[i].weight, e->dev_ops[i].bias, c->d_tensor[op.out],
                       B, op.hin, op.win, op.cout, op.hout, op.wout, op.k, op.stride, op.pad, op.act};
            BN_CUDA(launch_dwconv(d, s));
        } else {
            ConvParams cp{c->d_tensor[op.in], op.in_scale >= 0 ? c->d_tensor[op.in_scale] : nullptr,
                          e->dev_ops[i].weight, e->dev_ops[i].bias,
                          op.residual >= 0 ? c->d_tensor[op.residual] : nullptr, c->d_tensor[op.out],
                          B, op.hin, op.win, op.cin, op.hout, op.wout, op.cout, op.ldw, op.k, op.stride, op.pad, op.act};
            BN_CUDA(launch_conv_igemm(cp, s));
        }
        ++launches;
    }
    prof_mark(c, "topk_epilogue");
    TopkParams tp{};
    tp.logits = c->d_tensor[p.logits_tensor];
    tp.batch = B;
    tp.n = p.num_species;
    tp.k = (uint32_t)k_eff;
    tp.has_min_conf = post.has_min_conf;
    tp.min_conf = post.min_conf;
    tp.range_state = post.range ? post.range->state : nullptr;
    tp.range_score = post.range ? post.range->score : nullptr;
    tp.rerank = post.range ? post.range->rerank : 0;
    tp.out = c->d_topk;
    tp.out_count = c->d_count;
    BN_CUDA(launch_topk(tp, s));
    ++launches;
    prof_mark(c, "d2h");
    c->last_launches = launches;
    return BN_OK;
}

static int enqueue_fetch(bn_ctx* c, int B, uint64_t k_eff) {
    const Plan& p = c->eng->plan;
    cudaStream_t s = c->stream;
    BN_CUDA(cudaMemcpyAsync(c->h_logits, c->d_tensor[p.logits_tensor], (size_t)B * p.num_species * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (p.embedding_tensor >= 0)
        BN_CUDA(cudaMemcpyAsync(c->h_emb, c->d_tensor[p.embedding_tensor], (size_t)B * p.embedding_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (k_eff > 0) BN_CUDA(cudaMemcpyAsync(c->h_topk, c->d_topk, (size_t)B * k_eff * sizeof(Pred), cudaMemcpyDeviceToHost, s));
    BN_CUDA(cudaMemcpyAsync(c->h_count, c->d_count, (size_t)B * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return BN_OK;
}

// Monitor loop of Classifier::run_inference (src/classifier.rs:527-554): completed first, then
// cancellation, then timeout.  A finished run wins over a fired monitor (classifier.rs:569).
static int wait_done(bn_ctx* c, const bn_run_opts* opts) {
    using clock = std::chrono::steady_clock;
    const bool monitor = opts && (opts->cancel_flag || opts->has_timeout);
    if (!monitor) {
        BN_CUDA(cudaEventSynchronize(c->done));
        return BN_OK;
    }
    const auto t0 = clock::now();
    int spins = 0;
    while (true) {
        cudaError_t q = cudaEventQuery(c->done);
        if (q == cudaSuccess) return BN_OK;
        if (q != cudaErrorNotReady) return cuda_fail(q, "cudaEventQuery");
        if (opts->cancel_flag && *opts->cancel_flag != 0) {
            c->draining = true;
            return set_error(BN_ERR_CANCELLED, "inference was cancelled");
        }
        if (opts->has_timeout) {
            uint64_t el = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(clock::now() - t0).count();
            if (el >= opts->timeout_ns) {
                c->draining = true;
                return set_error_detail(BN_ERR_TIMEOUT, "inference timed out", opts->timeout_ns, 0, 0);
            }
        }
        if (++spins < 200) std::this_thread::yield();
        else std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

static void fill_outputs(bn_ctx* c, uint64_t B, uint64_t k_eff, bn_outputs* out) {
    const Plan& p = c->eng->plan;
    memset(out, 0, sizeof(*out));
    out->batch = B;
    out->num_species = (uint64_t)p.num_species;
    out->logits = c->h_logits;
    out->embedding_dim = (uint64_t)p.embedding_dim;
    out->embeddings = p.embedding_tensor >= 0 ? c->h_emb : nullptr;
    out->topk_stride = k_eff;
    out->topk_count = c->h_count;
    out->topk = reinterpret_cast<const bn_pred*>(c->h_topk);
}

static int begin_run(bn_ctx* c, PostCfg& post, uint64_t& k_eff, const bn_run_opts* opts) {
    bn_engine* e = c->eng;
    BN_CUDA(cudaSetDevice(e->device));
    if (c->draining) {                       // a timed-out / cancelled run may still be in flight
        BN_CUDA(cudaStreamSynchronize(c->stream));
        c->draining = false;
    }
    {
        std::lock_guard<std::mutex> lk(e->post_mu);
        post = e->post;
    }
    if (post.range && post.range->n != (uint64_t)e->plan.num_species)
        return set_error(BN_ERR_INFERENCE, "range filter has " + std::to_string(post.range->n) + " classes, model has " + std::to_string(e->plan.num_species));
    k_eff = std::min<uint64_t>(post.top_k, (uint64_t)e->plan.num_species);   // postprocess.rs:50
    int st = ensure_topk_capacity(c, k_eff);
    if (st != BN_OK) return st;
    c->range_in_flight = post.range;
    c->prof_names.clear();
    // a token cancelled before the run starts terminates it at the first poll (classifier.rs:536-541)
    if (opts && opts->cancel_flag && *opts->cancel_flag != 0) return set_error(BN_ERR_CANCELLED, "inference was cancelled");
    return BN_OK;
}

int ctx_enqueue_device(bn_ctx* c, const float* d_audio, uint64_t batch, bool fetch, const bn_run_opts* opts) {
    if (!c) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) return BN_OK;
    if (batch > c->max_batch)
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    PostCfg post;
    uint64_t k_eff = 0;
    int st = begin_run(c, post, k_eff, opts);
    if (st != BN_OK) return st;
    st = enqueue_forward(c, d_audio, (int)batch, post, k_eff);
    if (st != BN_OK) return st;
    if (fetch) { st = enqueue_fetch(c, (int)batch, k_eff); if (st != BN_OK) return st; }
    prof_mark(c, "end");
    BN_CUDA(cudaEventRecord(c->done, c->stream));
    c->pending_batch = batch;
    c->pending_k = k_eff;
    return BN_OK;
}

int ctx_wait(bn_ctx* c, const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    BN_CUDA(cudaSetDevice(c->eng->device));
    int st = wait_done(c, opts);
    if (st != BN_OK) return st;
    fill_outputs(c, c->pending_batch, c->pending_k, out);
    if (c->profiling) {
        size_t n = c->prof_names.size();
        c->prof_ms.assign(n ? n - 1 : 0, 0.f);
        for (size_t i = 0; i + 1 < n; ++i) cudaEventElapsedTime(&c->prof_ms[i], c->prof_events[i], c->prof_events[i + 1]);
    }
    return BN_OK;
}

int ctx_run_device(bn_ctx* c, const float* d_audio, uint64_t batch, bool fetch, const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) { memset(out, 0, sizeof(*out)); return BN_OK; }
    int st = ctx_enqueue_device(c, d_audio, batch, fetch, opts);
    if (st != BN_OK) return st;
    return ctx_wait(c, opts, out);
}

// ---- host staging: gather caller slices into the pinned slab and ship them chunk by chunk ----
static int stage_input(bn_ctx* c, const float* const* seg_ptrs, uint64_t B) {
    bn_engine* e = c->eng;
    const size_t S = (size_t)e->plan.sample_count;
    const size_t seg_bytes = S * sizeof(float);
    const uint64_t chunk = 8;
    const uint64_t n_chunks = (B + chunk - 1) / chunk;
    int T = (int)std::min<uint64_t>((uint64_t)e->pack_threads, n_chunks);
    if (T <= 1) {
        for (uint64_t ci = 0; ci < n_chunks; ++ci) {
            uint64_t lo = ci * chunk, hi = std::min(B, lo + chunk);
            for (uint64_t i = lo; i < hi; ++i) memcpy(c->h_in + i * S, seg_ptrs[i], seg_bytes);   // batch_context.rs:209-211
            BN_CUDA(cudaMemcpyAsync(c->d_in + lo * S, c->h_in + lo * S, (hi - lo) * seg_bytes, cudaMemcpyHostToDevice, c->stream));
        }
        return BN_OK;
    }
    std::atomic<uint64_t> next{0};
    std::atomic<int> err{(int)cudaSuccess};
    auto worker = [&]() {
        cudaSetDevice(e->device);
        while (true) {
            uint64_t ci = next.fetch_add(1);
            if (ci >= n_chunks) break;
            uint64_t lo = ci * chunk, hi = std::min(B, lo + chunk);
            for (uint64_t i = lo; i < hi; ++i) memcpy(c->h_in + i * S, seg_ptrs[i], seg_bytes);
            cudaError_t ce = cudaMemcpyAsync(c->d_in + lo * S, c->h_in + lo * S, (hi - lo) * seg_bytes, cudaMemcpyHostToDevice, c->stream);
            if (ce != cudaSuccess) err.store((int)ce);
        }
    };
    std::vector<std::thread> th;
    th.reserve(T - 1);
    for (int t = 0; t < T - 1; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    if (err.load() != (int)cudaSuccess) return cuda_fail((cudaError_t)err.load(), "cudaMemcpyAsync(H2D)");
    return BN_OK;
}

int ctx_run_host(bn_ctx* c, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
                 bool check_max_first, const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) { memset(out, 0, sizeof(*out)); return BN_OK; }          // classifier.rs:681-683, 832-834
    if (!seg_ptrs || !seg_lens) return set_error(BN_ERR_INVALID_ARGUMENT, "null segment array");
    const uint64_t S = (uint64_t)c->eng->plan.sample_count;
    if (check_max_first && batch > c->max_batch)                               // batch_context.rs:191-196
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    for (uint64_t i = 0; i < batch; ++i)                                       // classifier.rs:688-696, batch_context.rs:199-206
        if (seg_lens[i] != S)
            return set_error_detail(BN_ERR_BATCH_INPUT_SIZE, "batch input size mismatch", i, S, seg_lens[i]);
    if (batch > c->max_batch)
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    PostCfg post;
    uint64_t k_eff = 0;
    int st = begin_run(c, post, k_eff, opts);
    if (st != BN_OK) return st;
    prof_mark(c, "h2d");
    st = stage_input(c, seg_ptrs, batch);
    if (st != BN_OK) return st;
    st = enqueue_forward(c, c->d_in, (int)batch, post, k_eff);
    if (st != BN_OK) return st;
    st = enqueue_fetch(c, (int)batch, k_eff);
    if (st != BN_OK) return st;
    prof_mark(c, "end");
    BN_CUDA(cudaEventRecord(c->done, c->stream));
    st = wait_done(c, opts);
    if (st != BN_OK) return st;
    fill_outputs(c, batch, k_eff, out);
    if (c->profiling) {
        size_t n = c->prof_names.size();
        c->prof_ms.assign(n ? n - 1 : 0, 0.f);
        for (size_t i = 0; i + 1 < n; ++i) cudaEventElapsedTime(&c->prof_ms[i], c->prof_events[i], c->prof_events[i + 1]);
    }
    return BN_OK;
}

}  // namespace bn
