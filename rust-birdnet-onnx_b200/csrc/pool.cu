// Multi-GPU dispatcher: one engine replica + context + host thread per
// device, contiguous block partition of the segment range, host-side gather into caller order.
// There is no cross-segment reduction, hence no collective.  The reference's only multi-GPU
// hook is CUDAConfig::with_device_id (src/cuda_config.rs:179-182).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "engine.h"

using namespace bn;

struct bn_pool {
    std::vector<bn_engine*> engines;
    std::vector<bn_ctx*> ctxs;
    uint64_t ctx_batch = 0;
    ~bn_pool() {
        for (auto* c : ctxs) delete c;
        for (auto* e : engines) delete e;
    }
};

extern "C" {

int bn_pool_create(const char* onnx_path, const int32_t* device_ids, int32_t n_devices, int32_t model_type_override,
                   uint64_t ctx_batch, bn_pool** out) {
    if (!out || n_devices <= 0 || !device_ids || ctx_batch == 0) return set_error(BN_ERR_INVALID_ARGUMENT, "bad pool arguments");
    *out = nullptr;
    std::unique_ptr<bn_pool> p(new bn_pool());
    p->ctx_batch = ctx_batch;
    for (int i = 0; i < n_devices; ++i) {
        bn_device_cfg cfg{device_ids[i], model_type_override, 0, 0};
        bn_engine* e = nullptr;
        int st = engine_create(onnx_path, &cfg, &e);
        if (st != BN_OK) return st;
        p->engines.push_back(e);
        bn_ctx* c = nullptr;
        st = ctx_create(e, ctx_batch, &c);
        if (st != BN_OK) return st;
        p->ctxs.push_back(c);
    }
    *out = p.release();
    return BN_OK;
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
    if (!seg_ptrs || !seg_lens || !logits) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    const uint64_t S = pool->engines[0]->info.sample_count;
    const uint64_t N = pool->engines[0]->info.num_species;
    const uint64_t E = pool->engines[0]->info.embedding_dim;
    for (uint64_t i = 0; i < n_segments; ++i)
        if (seg_lens[i] != S) return set_error_detail(BN_ERR_BATCH_INPUT_SIZE, "batch input size mismatch", i, S, seg_lens[i]);
    const uint64_t D = pool->engines.size();
    const uint64_t nb = (n_segments + pool->ctx_batch - 1) / pool->ctx_batch;   // batches, block-partitioned
    std::vector<int> status(D, BN_OK);
    std::vector<std::string> msgs(D);
    std::vector<uint64_t> detail(D * 3, 0);
    std::atomic<int> abort_flag{0};
    auto work = [&](uint64_t d) {
        uint64_t b0 = d * nb / D, b1 = (d + 1) * nb / D;
        for (uint64_t b = b0; b < b1 && !abort_flag.load(); ++b) {
            uint64_t lo = b * pool->ctx_batch, hi = std::min(n_segments, lo + pool->ctx_batch);
            bn_outputs o;
            int st = ctx_run_host(pool->ctxs[d], seg_ptrs + lo, seg_lens + lo, hi - lo, true, opts, &o);
            if (st != BN_OK) {
                status[d] = st;
                msgs[d] = last_error();
                const uint64_t* dt = last_detail();
                detail[d * 3] = dt[0] + (st == BN_ERR_BATCH_INPUT_SIZE ? lo : 0);
                detail[d * 3 + 1] = dt[1]; detail[d * 3 + 2] = dt[2];
                abort_flag.store(1);
                return;
            }
            memcpy(logits + lo * N, o.logits, (hi - lo) * N * sizeof(float));
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
    std::vector<std::thread> th;
    for (uint64_t d = 1; d < D; ++d) th.emplace_back(work, d);
    work(0);
    for (auto& t : th) t.join();
    for (uint64_t d = 0; d < D; ++d)
        if (status[d] != BN_OK) return set_error_detail(status[d], msgs[d], detail[d * 3], detail[d * 3 + 1], detail[d * 3 + 2]);
    return BN_OK;
}

}  // extern "C"
