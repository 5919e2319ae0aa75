/* End-to-end driver in C over the C ABI (include/birdnet_b200.h): what a Rust / C host pays, without the
 * Python facade.  Host slices in, host results out, every copy inside the timed region.
 *
 *   gcc -O2 -std=c11 -o tools/_build/e2e_driver tools/e2e_driver.c -Iinclude \
 *       -Lrust-birdnet-onnx_b200/lib -lbirdnet_b200 -Wl,-rpath,'$ORIGIN/../../rust-birdnet-onnx_b200/lib' -lpthread -lm
 *   tools/_build/e2e_driver --model models/birdnet_v24_seed0.onnx [--mode ctx|pool] [--devices 0,1,..]
 *       [--batch 256] [--batches 64] [--depth 5] [--pinned 0|1] [--range 0|1] [--reps 5]
 *
 * mode ctx : per device one engine + `depth` contexts, one host thread per context calling bn_ctx_run
 *            (= predict_batch_with_context, src/classifier.rs:826-867) on `batches` batches in total per device.
 * mode pool: bn_pool_run over all devices (shared batch queue, `depth` contexts per device) on
 *            devices x batches x batch segments in ONE call (BASELINE.json configs[4] shape).
 * Prints one JSON line: segments/s = segments processed / wall time (median of --reps runs). */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "birdnet_b200.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void die(const char* what, int st) {
    fprintf(stderr, "%s failed: status %d: %s\n", what, st, bn_last_error());
    exit(2);
}

/* the reference's test generators: LCG noise (testutil.rs:115-118) and sines (integration_test.rs:57-67) */
static void fill_audio(float* a, uint64_t nseg, uint64_t S) {
    for (uint64_t s = 0; s < nseg; ++s) {
        float* x = a + s * S;
        uint64_t st = 1000 + s;
        const int kind = (int)(s % 4);
        for (uint64_t i = 0; i < S; ++i) {
            st = st * 1103515245ull + 12345ull;
            const float noise = (float)((st >> 16) & 0xFFFF) * (2.0f / 65535.0f) - 1.0f;
            const float tone = 0.5f * sinf(6.2831853f * (440.0f + 37.0f * (float)s) * (float)i / 48000.0f);
            x[i] = kind == 0 ? noise : kind == 1 ? tone : kind == 2 ? 0.3f * noise + tone : 1e-3f * noise;
        }
    }
}

typedef struct {
    bn_ctx* ctx;
    const float* const* ptrs;
    const uint64_t* lens;
    uint64_t batch, n_batches;
    uint64_t sink;
} worker_t;

static void* ctx_worker(void* arg) {
    worker_t* w = (worker_t*)arg;
    for (uint64_t b = 0; b < w->n_batches; ++b) {
        bn_outputs out;
        int st = bn_ctx_run(w->ctx, w->ptrs, w->lens, w->batch, NULL, &out);
        if (st != BN_OK) die("bn_ctx_run", st);
        /* touch every result the reference would materialise (raw_scores + predictions) */
        w->sink += out.topk_count[0] + (uint64_t)(out.logits[out.num_species * (w->batch - 1)] > 0.f);
    }
    return NULL;
}

static int cmp_d(const void* a, const void* b) { return (*(const double*)a > *(const double*)b) - (*(const double*)a < *(const double*)b); }

int main(int argc, char** argv) {
    const char* model = NULL;
    const char* mode = "ctx";
    const char* devs = "0";
    uint64_t batch = 256, batches = 64;
    int depth = 5, pinned = 0, range = 0, reps = 5;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--model")) model = argv[i + 1];
        else if (!strcmp(argv[i], "--mode")) mode = argv[i + 1];
        else if (!strcmp(argv[i], "--devices")) devs = argv[i + 1];
        else if (!strcmp(argv[i], "--batch")) batch = strtoull(argv[i + 1], NULL, 10);
        else if (!strcmp(argv[i], "--batches")) batches = strtoull(argv[i + 1], NULL, 10);
        else if (!strcmp(argv[i], "--depth")) depth = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--pinned")) pinned = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--range")) range = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
    }
    if (!model) { fprintf(stderr, "usage: e2e_driver --model file.onnx [...]\n"); return 2; }
    int32_t ids[16];
    int nd = 0;
    { char* tmp = strdup(devs); for (char* t = strtok(tmp, ","); t && nd < 16; t = strtok(NULL, ",")) ids[nd++] = atoi(t); free(tmp); }
    bn_io_info info;
    int st = bn_model_inspect(model, BN_MODEL_AUTO, &info);
    if (st != BN_OK) die("bn_model_inspect", st);
    const uint64_t S = info.sample_count, N = info.num_species;
    /* one batch of host audio, reused by every call (147 MB for v2.4: larger than the CPU's last-level cache) */
    const size_t bytes = (size_t)batch * S * sizeof(float);
    float* audio = pinned ? (float*)bn_host_alloc(bytes) : (float*)malloc(bytes);
    if (!audio) { fprintf(stderr, "host allocation of %zu bytes failed\n", bytes); return 2; }
    fill_audio(audio, batch, S);
    uint8_t* rstate = NULL;
    float* rscore = NULL;
    if (range) {          /* dense tri-state mask from a seeded score vector */
        rstate = (uint8_t*)malloc(N);
        rscore = (float*)malloc(N * sizeof(float));
        uint64_t s = 42;
        for (uint64_t i = 0; i < N; ++i) {
            s = s * 1103515245ull + 12345ull;
            rscore[i] = (float)((s >> 16) & 0xFFFF) / 65535.0f;
            rstate[i] = (i % 10) < 3 ? 0 : (rscore[i] >= 0.01f ? 1 : 2);
        }
    }
    double* rates = (double*)calloc((size_t)reps, sizeof(double));
    uint64_t total_segments = 0;
    if (!strcmp(mode, "pool")) {
        bn_pool* pool = NULL;
        st = bn_pool_create_ex(model, ids, nd, BN_MODEL_AUTO, batch, depth, &pool);
        if (st != BN_OK) die("bn_pool_create_ex", st);
        bn_pool_set_postprocess(pool, 5, 1, 0.1f);
        if (range) { st = bn_pool_set_range_filter(pool, rstate, rscore, N, 1); if (st != BN_OK) die("bn_pool_set_range_filter", st); }
        const uint64_t n = (uint64_t)nd * batches * batch;
        const float** ptrs = (const float**)malloc(n * sizeof(*ptrs));
        uint64_t* lens = (uint64_t*)malloc(n * sizeof(*lens));
        for (uint64_t i = 0; i < n; ++i) { ptrs[i] = audio + (i % batch) * S; lens[i] = S; }
        float* logits = (float*)malloc(n * N * sizeof(float));
        bn_pred* topk = (bn_pred*)malloc(n * 5 * sizeof(bn_pred));
        uint32_t* cnt = (uint32_t*)malloc(n * sizeof(uint32_t));
        memset(logits, 0, n * N * sizeof(float));           /* fault the pages in before the clock starts */
        st = bn_pool_run(pool, ptrs, lens, (uint64_t)nd * depth * batch, NULL, logits, NULL, topk, cnt, 5);   /* warm-up */
        if (st != BN_OK) die("bn_pool_run", st);
        for (int r = 0; r < reps; ++r) {
            const double t0 = now_s();
            st = bn_pool_run(pool, ptrs, lens, n, NULL, logits, NULL, topk, cnt, 5);
            if (st != BN_OK) die("bn_pool_run", st);
            rates[r] = (double)n / (now_s() - t0);
        }
        total_segments = n;
        bn_pool_destroy(pool);
    } else {
        const int nw = nd * depth;
        bn_engine** eng = (bn_engine**)calloc((size_t)nd, sizeof(*eng));
        worker_t* w = (worker_t*)calloc((size_t)nw, sizeof(*w));
        const float** ptrs = (const float**)malloc(batch * sizeof(*ptrs));
        uint64_t* lens = (uint64_t*)malloc(batch * sizeof(*lens));
        for (uint64_t i = 0; i < batch; ++i) { ptrs[i] = audio + i * S; lens[i] = S; }
        for (int d = 0; d < nd; ++d) {
            bn_device_cfg cfg = {ids[d], BN_MODEL_AUTO, 0, 0};
            st = bn_engine_create(model, &cfg, &eng[d]);
            if (st != BN_OK) die("bn_engine_create", st);
            bn_engine_set_postprocess(eng[d], 5, 1, 0.1f);
            if (range) { st = bn_engine_set_range_filter(eng[d], rstate, rscore, N, 1); if (st != BN_OK) die("bn_engine_set_range_filter", st); }
            for (int k = 0; k < depth; ++k) {
                worker_t* x = &w[d * depth + k];
                st = bn_ctx_create(eng[d], batch, &x->ctx);
                if (st != BN_OK) die("bn_ctx_create", st);
                x->ptrs = ptrs; x->lens = lens; x->batch = batch;
                x->n_batches = 2;
                ctx_worker(x);                               /* warm-up */
                x->n_batches = (batches + depth - 1) / depth;
            }
        }
        pthread_t* th = (pthread_t*)calloc((size_t)nw, sizeof(*th));
        for (int r = 0; r < reps; ++r) {
            const double t0 = now_s();
            for (int i = 0; i < nw; ++i) pthread_create(&th[i], NULL, ctx_worker, &w[i]);
            for (int i = 0; i < nw; ++i) pthread_join(th[i], NULL);
            total_segments = (uint64_t)nw * w[0].n_batches * batch;
            rates[r] = (double)total_segments / (now_s() - t0);
        }
        for (int i = 0; i < nw; ++i) bn_ctx_destroy(w[i].ctx);
        for (int d = 0; d < nd; ++d) bn_engine_destroy(eng[d]);
    }
    qsort(rates, (size_t)reps, sizeof(double), cmp_d);
    printf("{\"driver\": \"c\", \"mode\": \"%s\", \"devices\": %d, \"depth\": %d, \"batch\": %llu, \"segments_per_run\": %llu, "
           "\"inputs\": \"%s\", \"range_filter\": %d, \"h2d_bytes_per_batch\": %zu, \"d2h_bytes_per_batch\": %llu, "
           "\"segments_per_s_median\": %.1f, \"segments_per_s_min\": %.1f, \"segments_per_s_max\": %.1f}\n",
           mode, nd, depth, (unsigned long long)batch, (unsigned long long)total_segments,
           pinned ? "page-locked (bn_host_alloc)" : "pageable malloc", range, bytes,
           (unsigned long long)(batch * N * 4 + batch * 5 * 8 + batch * 4), rates[reps / 2], rates[0], rates[reps - 1]);
    return 0;
}
