// Multi-GPU dispatcher: per device one engine replica and `depth` contexts, each driven by its
// own host thread; all threads pull whole batches from ONE shared queue (an atomic batch counter), so a device's
// H2D copy, kernels and D2H of consecutive batches overlap (one batch per context in flight) and a slower device
// simply takes fewer batches.  Results are gathered on the host in caller order.  There is no cross-segment
// reduction, hence no collective.  The reference's only multi-GPU hook is CUDAConfig::with_device_id
// (src/cuda_config.rs:179-182).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "engine.h"

using namespace bn;

struct bn_pool {
    std::vector<bn_engine*> engines;
    std::vector<bn_ctx*> ctxs;          // [device][depth]
    int depth = 3;
    uint64_t ctx_batch = 0;
    ~bn_pool() {
        for (auto* c : ctxs) delete c;
        for (auto* e : engines) delete e;
    }
};

extern "C" {

int bn_pool_create_ex(const char* onnx_path, const int32_t* device_ids, int32_t n_devices, int32_t model_type_override,
                      uint64_t ctx_batch, int32_t depth, bn_pool** out) {
    if (!out || n_devices <= 0 || !device_ids || ctx_batch == 0) return set_error(BN_ERR_INVALID_ARGUMENT, "bad pool arguments");
    if (depth < 1 || depth > 8) return set_error(BN_ERR_INVALID_ARGUMENT, "pool depth must be in [1, 8]");
    *out = nullptr;
    std::unique_ptr<bn_pool> p(new bn_pool());
    p->ctx_batch = ctx_batch;
    p->depth = depth;
    unsigned hw = std::thread::hardware_concurrency();
    // gather threads per call: the box's cores are shared by every context of every device
    const int pack = (int)std::max(1u, std::min(8u, hw / (unsigned)(2 * n_devices * depth)));
    for (int i = 0; i < n_devices; ++i) {
        bn_device_cfg cfg{device_ids[i], model_type_override, pack, 0};
        bn_engine* e = nullptr;
        int st = engine_create(onnx_path, &cfg, &e);
        if (st != BN_OK) return st;
        p->engines.push_back(e);
        for (int d = 0; d < depth; ++d) {
            bn_ctx* c = nullptr;
            st = ctx_create(e, ctx_batch, &c);
            if (st != BN_OK) return st;
            p->ctxs.push_back(c);
        }
    }
    *out = p.release();
    return BN_OK;
}

int bn_pool_create(const char* onnx_path, const int32_t* device_ids, int32_t n_devices, int32_t model_type_override,
                   uint64_t ctx_batch, bn_pool** out) {
    int depth = 3;
    if (const char* ev = getenv("BN_POOL_DEPTH")) { const int v = atoi(ev); if (v >= 1 && v <= 8) depth = v; }
    return bn_pool_create_ex(onnx_path, device_ids, n_devices, model_type_override, ctx_batch, depth, out);
}

void bn_pool_destroy(bn_pool* pool) { delete pool; }

int bn_pool_set_postprocess(bn_pool* pool, uint64_t top_k, int32_t has_min_confidence, float min_confidence) {
    if (!pool) return set_error(BN_ERR_INVALID_ARGUMENT, "null pool");
    for (auto* e : pool->engines) {
        int st = bn_engine_set_postprocess(e, top_k, has_min_confidence, min_confidence);
        if (st != BN_OK) return st;
    }
    return BN_OK;
}

int bn_pool_set_range_filter(bn_pool* pool, const uint8_t* state, const float* score, uint64_t n, int32_t rerank) {
    if (!pool) return set_error(BN_ERR_INVALID_ARGUMENT, "null pool");
    for (auto* e : pool->engines) {
        int st = state ? bn_engine_set_range_filter(e, state, score, n, rerank) : bn_engine_clear_range_filter(e);
        if (st != BN_OK) return st;
    }
    return BN_OK;
}

int bn_pool_run(bn_pool* pool, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t n_segments,
                const bn_run_opts* opts, float* logits, float* embeddings, bn_pred* topk, uint32_t* topk_count,
                uint64_t topk_stride) {
    if (!pool) return set_error(BN_ERR_INVALID_ARGUMENT, "null pool");
    if (n_segments == 0) return BN_OK;
    if (!seg_ptrs || !seg_lens) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    const uint64_t S = pool->engines[0]->info.sample_count;
    const uint64_t N = pool->engines[0]->info.num_species;
    const uint64_t E = pool->engines[0]->info.embedding_dim;
    for (uint64_t i = 0; i < n_segments; ++i)
        if (seg_lens[i] != S) return set_error_detail(BN_ERR_BATCH_INPUT_SIZE, "batch input size mismatch", i, S, seg_lens[i]);
    const uint64_t W = pool->ctxs.size();                                       // workers = devices x depth
    const uint64_t nb = (n_segments + pool->ctx_batch - 1) / pool->ctx_batch;   // whole batches, last one ragged
    std::vector<int> status(W, BN_OK);
    std::vector<std::string> msgs(W);
    std::vector<uint64_t> detail(W * 3, 0);
    std::atomic<uint64_t> next{0};
    std::atomic<int> abort_flag{0};
    auto work = [&](uint64_t w) {
        bn_ctx* c = pool->ctxs[w];
        while (!abort_flag.load(std::memory_order_relaxed)) {
            const uint64_t b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= nb) break;
            const uint64_t lo = b * pool->ctx_batch, hi = std::min(n_segments, lo + pool->ctx_batch);
            bn_outputs o;
            int st = ctx_run_host(c, seg_ptrs + lo, seg_lens + lo, hi - lo, true, opts, &o);
            if (st != BN_OK) {
                status[w] = st;
                msgs[w] = last_error();
                const uint64_t* dt = last_detail();
                detail[w * 3] = dt[0] + (st == BN_ERR_BATCH_INPUT_SIZE ? lo : 0);
                detail[w * 3 + 1] = dt[1]; detail[w * 3 + 2] = dt[2];
                abort_flag.store(1);
                return;
            }
            if (logits) memcpy(logits + lo * N, o.logits, (hi - lo) * N * sizeof(float));
            if (embeddings && o.embeddings) memcpy(embeddings + lo * E, o.embeddings, (hi - lo) * E * sizeof(float));
            if (topk && topk_count) {
                uint64_t ks = std::min<uint64_t>(o.topk_stride, topk_stride);
                for (uint64_t i = lo; i < hi; ++i) {
                    uint32_t cnt = std::min<uint32_t>(o.topk_count[i - lo], (uint32_t)ks);
                    topk_count[i] = cnt;
                    memcpy(topk + i * topk_stride, o.topk + (i - lo) * o.topk_stride, cnt * sizeof(bn_pred));
                }
            }
        }
    };
    const uint64_t n_workers = std::min<uint64_t>(W, nb);
    std::vector<std::thread> th;
    // worker w drives context w; start them device-interleaved so a short job still spreads over all devices
    std::vector<uint64_t> order;
    const uint64_t D = pool->engines.size();
    for (int d = 0; d < pool->depth; ++d)
        for (uint64_t g = 0; g < D; ++g) order.push_back(g * pool->depth + d);
    for (uint64_t i = 1; i < n_workers; ++i) th.emplace_back(work, order[i]);
    work(order[0]);
    for (auto& t : th) t.join();
    for (uint64_t w = 0; w < W; ++w)
        if (status[w] != BN_OK) return set_error_detail(status[w], msgs[w], detail[w * 3], detail[w * 3 + 1], detail[w * 3 + 2]);
    return BN_OK;
}

}  // extern "C"
