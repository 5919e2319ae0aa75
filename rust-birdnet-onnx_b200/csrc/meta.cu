// Range-filter meta model on the device: replaces the ONNX Runtime session that
// RangeFilter::predict drives (src/rangefilter.rs:239-267 load + validation, 451-479 run) and builds the dense
// per-class tri-state the fused epilogue consumes without a round trip through host HashMaps
// (src/rangefilter.rs:340-343 builds one per call).
//
// The graph is a small MLP over [latitude, longitude, week]: {Mul by a constant vector}? -> (Gemm|MatMul+Add) ->
// Relu ... -> Sigmoid.  One CTA evaluates the whole net for the single input row: a warp per output neuron,
// lanes stride the input vector, shuffle reduction.
#include "engine.h"
#include "onnx_reader.h"

#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>

using namespace bn;

struct bn_meta {
    int device = 0;
    struct Layer { float* w = nullptr; float* b = nullptr; int in = 0, out = 0, act = 0; };   // act: 0 none, 1 relu, 2 sigmoid
    std::vector<Layer> layers;
    float in_scale[3] = {1.f, 1.f, 1.f};
    int n_out = 0;
    float* d_act[2] = {nullptr, nullptr};     // ping-pong activations
    float* d_in = nullptr;
    float* h_scores = nullptr;                 // pinned
    cudaStream_t stream = nullptr;
    std::mutex mu;                             // the reference keeps its session in a Mutex (src/rangefilter.rs:271, 390, 461-466)
    ~bn_meta() {
        cudaSetDevice(device);
        for (auto& l : layers) { if (l.w) cudaFree(l.w); if (l.b) cudaFree(l.b); }
        for (auto* p : d_act) if (p) cudaFree(p);
        if (d_in) cudaFree(d_in);
        if (h_scores) cudaFreeHost(h_scores);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

// y[o] = act(b[o] + sum_i w[o][i] * x[i]); one warp per output neuron
__global__ void __launch_bounds__(256) k_meta_fc(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                 float* __restrict__ y, int n_in, int n_out, int act) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= n_out) return;
    const float* wr = w + (size_t)o * n_in;
    float acc = 0.f;
    for (int i = lane; i < n_in; i += 32) acc = fmaf(wr[i], x[i], acc);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        float v = acc + b[o];
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) v = 1.0f / (1.0f + expf(-v));
        y[o] = v;
    }
}

// dense tri-state of filter_predictions_impl (rangefilter.rs:333-386) straight from the scores: a species is in the
// map iff score >= predict_threshold (predict() only returns those, 482-496); in the map, score >= filter_threshold
// keeps (x score when reranking), below drops; not in the map keeps unchanged.
__global__ void k_meta_state(const float* __restrict__ scores, int n, float predict_thr, float filter_thr,
                             uint8_t* __restrict__ state, float* __restrict__ out_score) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = scores[i];
    out_score[i] = s;
    state[i] = !(s >= predict_thr) ? 0 : (s >= filter_thr ? 1 : 2);
}

int meta_fail(int code, const std::string& msg) { return set_error(code, msg); }

}  // namespace

extern "C" {

int bn_meta_create(const char* onnx_path, int32_t device_id, bn_meta** out) {
    if (!out) return set_error(BN_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (!onnx_path) return set_error(BN_ERR_MODEL_PATH_REQUIRED, "model path required");
    OnnxModel m;
    try {
        load_onnx(onnx_path, m);
    } catch (const std::exception& ex) {
        return meta_fail(BN_ERR_MODEL_LOAD, ex.what());
    }
    if (m.outputs.size() != 1)                                              // rangefilter.rs:254-258
        return meta_fail(BN_ERR_MODEL_DETECTION, "meta model expects 1 output, got " + std::to_string(m.outputs.size()));
    if (m.outputs[0].dims.empty()) return meta_fail(BN_ERR_MODEL_DETECTION, "empty output shape");
    if (m.inputs.empty() || m.inputs[0].dims.empty() || m.inputs[0].dims.back() != 3)
        return meta_fail(BN_ERR_MODEL_DETECTION, "meta model input must be [1, 3] = [latitude, longitude, week]");
    std::unique_ptr<bn_meta> mm(new bn_meta());
    mm->device = device_id;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) { cudaGetLastError(); return meta_fail(BN_ERR_RUNTIME_INIT, "no CUDA device available"); }
    if (device_id < 0 || device_id >= ndev) return meta_fail(BN_ERR_RUNTIME_INIT, "device_id out of range");
    BN_CUDA(cudaSetDevice(device_id));
    // walk the chain from the graph input
    std::string cur = m.inputs[0].name;
    int cur_dim = 3;
    for (size_t ni = 0; ni < m.nodes.size(); ++ni) {
        const OnnxNode& n = m.nodes[ni];
        if (n.inputs.empty() || n.inputs[0] != cur) {
            if (n.op == "Mul" && n.inputs.size() == 2 && n.inputs[1] == cur) {}      // commuted scale
            else return meta_fail(BN_ERR_MODEL_LOAD, "meta model: unsupported graph topology at node " + n.op);
        }
        if (n.op == "Mul") {
            const OnnxTensor* t = m.init(n.inputs[0] == cur ? n.inputs[1] : n.inputs[0]);
            if (!t || t->numel() != 3 || !mm->layers.empty()) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: only an input scale of 3 constants is supported for Mul");
            for (int i = 0; i < 3; ++i) mm->in_scale[i] *= t->f32_at(i);
        } else if (n.op == "Gemm" || n.op == "MatMul") {
            const OnnxTensor* w = m.init(n.inputs[1]);
            if (!w || w->dims.size() != 2) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: weight initializer missing");
            const bool transB = n.op == "Gemm" && n.attr_i("transB", 0) != 0;
            const int wi = (int)(transB ? w->dims[1] : w->dims[0]), wo = (int)(transB ? w->dims[0] : w->dims[1]);
            if (wi != cur_dim) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: layer input width mismatch");
            std::vector<float> wt((size_t)wo * wi), bias((size_t)wo, 0.f);
            for (int o = 0; o < wo; ++o)
                for (int i = 0; i < wi; ++i) wt[(size_t)o * wi + i] = transB ? w->f32_at((size_t)o * wi + i) : w->f32_at((size_t)i * wo + o);
            if (n.op == "Gemm" && n.inputs.size() > 2) {
                const OnnxTensor* b = m.init(n.inputs[2]);
                if (!b || (int)b->numel() != wo) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: bias shape mismatch");
                for (int o = 0; o < wo; ++o) bias[o] = b->f32_at(o);
            }
            bn_meta::Layer L;
            L.in = wi; L.out = wo;
            BN_CUDA(cudaMalloc(&L.w, wt.size() * sizeof(float)));
            BN_CUDA(cudaMemcpy(L.w, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
            BN_CUDA(cudaMalloc(&L.b, bias.size() * sizeof(float)));
            BN_CUDA(cudaMemcpy(L.b, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
            mm->layers.push_back(L);
            cur_dim = wo;
        } else if (n.op == "Add") {
            const OnnxTensor* b = m.init(n.inputs[1]);
            if (!b || mm->layers.empty() || (int)b->numel() != cur_dim) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: unsupported Add");
            std::vector<float> bias((size_t)cur_dim);
            for (int o = 0; o < cur_dim; ++o) bias[o] = b->f32_at(o);
            BN_CUDA(cudaMemcpy(mm->layers.back().b, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
        } else if (n.op == "Relu" || n.op == "Sigmoid") {
            if (mm->layers.empty() || mm->layers.back().act != 0) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: activation without a layer");
            mm->layers.back().act = n.op == "Relu" ? 1 : 2;
        } else {
            return meta_fail(BN_ERR_MODEL_LOAD, "meta model: unsupported op " + n.op);
        }
        cur = n.outputs[0];
    }
    if (mm->layers.empty() || cur != m.outputs[0].name) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: the op chain does not end at the graph output");
    mm->n_out = cur_dim;
    const int64_t declared = m.outputs[0].dims.back();
    if (declared > 0 && declared != cur_dim) return meta_fail(BN_ERR_MODEL_LOAD, "meta model: declared output shape does not match the graph");
    int widest = 3;
    for (auto& l : mm->layers) widest = std::max(widest, l.out);
    BN_CUDA(cudaMalloc(&mm->d_act[0], widest * sizeof(float)));
    BN_CUDA(cudaMalloc(&mm->d_act[1], widest * sizeof(float)));
    BN_CUDA(cudaMalloc(&mm->d_in, 3 * sizeof(float)));
    BN_CUDA(cudaHostAlloc(&mm->h_scores, (size_t)mm->n_out * sizeof(float), cudaHostAllocDefault));
    BN_CUDA(cudaStreamCreateWithFlags(&mm->stream, cudaStreamNonBlocking));
    *out = mm.release();
    return BN_OK;
}

void bn_meta_destroy(bn_meta* m) { delete m; }

uint64_t bn_meta_num_outputs(const bn_meta* m) { return m ? (uint64_t)m->n_out : 0; }

static int meta_forward(bn_meta* m, float latitude, float longitude, float week, const float** d_scores) {
    BN_CUDA(cudaSetDevice(m->device));
    const float x[3] = {latitude * m->in_scale[0], longitude * m->in_scale[1], week * m->in_scale[2]};
    BN_CUDA(cudaMemcpyAsync(m->d_in, x, sizeof(x), cudaMemcpyHostToDevice, m->stream));
    const float* cur = m->d_in;
    int pp = 0;
    for (auto& l : m->layers) {
        k_meta_fc<<<(l.out + 7) / 8, 256, 0, m->stream>>>(cur, l.w, l.b, m->d_act[pp], l.in, l.out, l.act);
        BN_CUDA(cudaGetLastError());
        cur = m->d_act[pp];
        pp ^= 1;
    }
    *d_scores = cur;
    return BN_OK;
}

// RangeFilter::predict's session.run (rangefilter.rs:451-479): scores[n_out] for [latitude, longitude, week]
int bn_meta_predict(bn_meta* m, float latitude, float longitude, float week, float* scores, uint64_t n) {
    if (!m || !scores) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (n != (uint64_t)m->n_out) return set_error(BN_ERR_INVALID_ARGUMENT, "scores buffer must hold " + std::to_string(m->n_out) + " floats");
    std::lock_guard<std::mutex> run_lock(m->mu);
    const float* d = nullptr;
    int st = meta_forward(m, latitude, longitude, week, &d);
    if (st != BN_OK) return st;
    BN_CUDA(cudaMemcpyAsync(m->h_scores, d, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    BN_CUDA(cudaStreamSynchronize(m->stream));
    memcpy(scores, m->h_scores, n * sizeof(float));
    return BN_OK;
}

// predict + the dense mask of filter_predictions_impl, built on the device and installed as the engine's fused range
// filter.  Valid when the classifier's labels and the meta model's labels are the same list in the same order (true for
// BirdNET v2.4: both use the 6,522-line label file), which the caller asserts by calling this.
int bn_meta_install_range_filter(bn_meta* m, bn_engine* engine, float latitude, float longitude, float week,
                                 float predict_threshold, float filter_threshold, int32_t rerank) {
    if (!m || !engine) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (m->device != engine->device) return set_error(BN_ERR_INVALID_ARGUMENT, "meta model and engine live on different devices");
    if ((uint64_t)m->n_out != (uint64_t)engine->plan.num_species)
        return set_error(BN_ERR_INFERENCE, "meta model has " + std::to_string(m->n_out) + " classes, model has " + std::to_string(engine->plan.num_species));
    std::lock_guard<std::mutex> run_lock(m->mu);
    const float* d = nullptr;
    int st = meta_forward(m, latitude, longitude, week, &d);
    if (st != BN_OK) return st;
    auto r = std::make_shared<RangeDev>();
    r->device = m->device;
    r->n = (uint64_t)m->n_out;
    r->rerank = rerank ? 1 : 0;
    BN_CUDA(cudaMalloc(&r->state, r->n));
    BN_CUDA(cudaMalloc(&r->score, r->n * sizeof(float)));
    k_meta_state<<<(m->n_out + 255) / 256, 256, 0, m->stream>>>(d, m->n_out, predict_threshold, filter_threshold, r->state, r->score);
    BN_CUDA(cudaGetLastError());
    BN_CUDA(cudaStreamSynchronize(m->stream));
    std::lock_guard<std::mutex> lk(engine->post_mu);
    engine->post.range = r;
    return BN_OK;
}

}  // extern "C"
