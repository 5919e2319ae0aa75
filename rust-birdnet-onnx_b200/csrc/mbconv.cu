// Fused MBConv block (rows A8 of SURVEY.md section 8): 1x1 expand + SiLU -> depthwise kxk + SiLU -> squeeze-excite
// -> gated 1x1 projection (+ residual), ONE kernel per block, one CTA (12 worker warps + 1 control warp) per segment at a time.
//
//   E = silu(X * We + be)                    [npix][cexp]   tcgen05 into TMEM, never leaves the SM
//   D = silu(dw_kxk(E) + bd)                 [npix][cexp]   CUDA cores, E read from a shared-memory patch
//   g = sigmoid(W2^T silu(W1^T mean_p(D) + b1) + b2)        [cexp]
//   Y = (D * g) * Wp + bp (+ X)              [npix][cout]   tcgen05; the gate is applied to D's tiles in shared memory
//
// What used to be three launches (expand conv writing E as FP32, depthwise + SE kernel re-reading it and rescaling D
// in place, projection conv) with E and D round-tripping through HBM is now: X read once (TMA), D written once and
// read back once as the projection's A operand (it is needed only after the gate, which depends on all of D; 0.2-0.5 MB
// per segment, written microseconds earlier by the same SM: it is served by L2), Y written once.
//
// Phase A, per group of G expanded channels (double-buffered in TMEM, so the MMAs of group g+1 run under the
// depthwise arithmetic of group g):
//   control warp (one lane)     E_g[npix][G] = X[npix][cin] * We[cin][G]      (hi/lo operands: 3 MMAs per K step, main |
//                                                                             correction accumulators as in tc_conv.cu)
//   12 worker warps             TMEM -> +bias, SiLU -> FP32 patch [G/2 channel pairs][zero-halo pixels] in smem
//   384 worker threads          depthwise conv from the patch (thread = channel pair x XB x YB output block, packed
//                               FFMA2), SiLU, D -> global hi/lo planes, pooled partial sums (fixed order: deterministic)
// tcgen05.mma issue blocks the issuing thread while the tensor queue is full (measured: ~2.7 k cycles per group when a
// worker issued), hence the dedicated control warp: the workers never wait behind the tensor pipe.
// Gate: the two tiny FCs by the workers.
// Phase B, per 64-channel K chunk: TMA lands D's hi/lo tiles (2-3 stage ring) and a bulk copy lands the pre-packed
//   [W_hi | W_lo] projection weights of the chunk; the workers multiply the D tile by the gate in place (hi + lo -> FP32 ->
//   x g[c] -> hi/lo), the control lane issues the MMAs; the epilogue adds bias and residual through an FP32 staging tile.
// Small images: a CTA pass takes `segs` consecutive segments (rows = segs * npix <= 256) so the 128-row MMA tiles, the
//   FC weights and the projection weights are amortised over them (3x16 images: two segments per pass).
#include "mbconv.h"
#include "tc_common.cuh"
#include "fast_act.cuh"

#include <cstdio>
#include <cstdlib>

namespace bn {
using namespace tc;

namespace {

constexpr int MB_WARPS = 12;                  // worker warps (TMEM reads, depthwise arithmetic, weight images, epilogue)
constexpr int MB_THREADS = MB_WARPS * 32;
constexpr int MB_BLOCK = MB_THREADS + 32;     // + one control warp: TMA / bulk loads and every tcgen05.mma issue
// 13 warps: one SM sub-partition hosts 4 of them, so 128 registers per thread is the hardware limit (4 x 32 x 128 = its 16 K
// registers; __maxnreg__(144) fails at launch).  Register spills in the depthwise loop cost 20 % (measured), so the per-segment
// arrays of the gate phase are sized by the instantiation (PM) and the weights are loaded right before the loop.
constexpr uint32_t BOX_BYTES = 64 * 128;      // one TMA box: 64 rows x 64 fp16 channels, SWIZZLE_128B
#ifndef MB_SERIAL_TMEM
#define MB_SERIAL_TMEM 0                      // 1: issue group g+1 only after the workers' TMEM reads of group g (measured: no gain)
#endif

__device__ __forceinline__ float silu_f(float v) { return silu_approx(v); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MB_THREADS) : "memory"); }     // the 12 worker warps only
__device__ __forceinline__ void arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

}  // namespace

// shared-memory carve-up (bytes from the 1024-aligned base); host and device agree through this function
__host__ __device__ inline MbLayout mb_layout(int h, int w, int k, int cin, int cexp, int cout, int r, int G, int segs) {
    MbLayout L;
    const int rows = segs * h * w, pad = k / 2;
    L.n_box = (rows + 63) / 64;
    L.n_mt = (rows + 127) / 128;
    L.kc_e = (cin + 63) / 64;
    L.kc_p = (cexp + 63) / 64;
    L.npixp = (segs * (h + 2 * pad) * (w + 2 * pad) + 15) / 16 * 16 + 1;   // = 1 mod 16: channel pairs land on distinct banks
    L.xa_bytes = (uint32_t)L.kc_e * 2u * (uint32_t)L.n_box * BOX_BYTES;
    L.we_stage = (uint32_t)L.kc_e * 2u * (uint32_t)G * 128u;
    L.da_stage = 2u * (uint32_t)L.n_box * BOX_BYTES;
    L.wp_stage = 2u * (uint32_t)cout * 128u;
    L.off_xa = 0;
    L.off_we = L.xa_bytes;
    const uint32_t a_end = L.xa_bytes + 3u * L.we_stage + BOX_BYTES;     // three weight stages; + one box: the last M tile may read past its 64 rows
    // phase A: [X tiles | 2 expand weight stages | patch]; phase B reuses ALL of it (the patch is re-zeroed per pass):
    // [n_da D stages | 2 projection weight stages], then the FP32 staging tile of the epilogue over the same bytes
    L.off_patch = (a_end + 1023u) & ~1023u;
    L.patch_bytes = (uint32_t)(G / 2) * (uint32_t)L.npixp * 8u;
    const uint32_t pa_end = L.off_patch + ((L.patch_bytes + 15u) & ~15u);
    L.stg_pitch = (uint32_t)cout * 4u + 16u;                             // FP32 row + 16 B: conflict-free 16-byte lane stores
    const uint32_t stg_end = (uint32_t)rows * L.stg_pitch;
    L.off_da = 0;
    L.n_da = 3;
    uint32_t b_end = 3u * L.da_stage + 2u * L.wp_stage + BOX_BYTES;
    if (b_end > (pa_end > 184u * 1024u ? pa_end : 184u * 1024u)) { L.n_da = 2; b_end = 2u * L.da_stage + 2u * L.wp_stage + BOX_BYTES; }     // a third D stage only when it costs no extra shared memory
    L.off_wp = (uint32_t)L.n_da * L.da_stage;
    uint32_t o = pa_end > b_end ? pa_end : b_end;
    if (stg_end > o) o = stg_end;
    o = (o + 1023u) & ~1023u;
    L.off_wd = o;    o += 2u * (uint32_t)(k * k + 2) * (uint32_t)G * 4u;                // depthwise weights + both biases of a group, double buffered
    L.off_part = o;  o += 2u * (uint32_t)(MB_THREADS / (G / 2)) * (uint32_t)G * 4u;     // pooled partials per block slot, double buffered
    L.off_pool = o;  o += (uint32_t)segs * (uint32_t)cexp * 4u;
    L.off_gate = L.off_pool;                                             // the gate overwrites the pooled means (dead after FC1)
    L.off_r = o;     o += (uint32_t)segs * (uint32_t)((r + 3) & ~3) * 4u;
    L.off_fc = o;    o += (uint32_t)MB_WARPS * (uint32_t)segs * (uint32_t)r * 4u;
    L.total = o + 1024u;                                                 // alignment slack
    return L;
}

MbLayout mb_layout_host(int h, int w, int k, int cin, int cexp, int cout, int r, int G, int segs) { return mb_layout(h, w, k, cin, cexp, cout, r, G, segs); }

constexpr int MB_PMAX = 2;                    // segments per CTA pass (small images), at most

// K: depthwise kernel size; G: expanded channels per group; XB x YB: output pixels per worker thread;
// GATEA: the squeeze-excite gate is applied to D's tiles in shared memory (any number of segments per pass, pre-packed
// projection weights) instead of to the projection weights (one segment per pass, cout <= 144: fewer elements to scale
// when the image has more pixels than the projection has output channels)
template <int K, int G, int XB, int YB, bool GATEA>
__global__ void __launch_bounds__(MB_BLOCK, 1) k_mbconv(const __grid_constant__ MbconvParams p) {
    extern __shared__ __align__(1024) uint8_t mb_smem_raw[];
    __shared__ __align__(8) uint64_t bar_x;          // block input landed
    __shared__ __align__(8) uint64_t bar_w[3];       // expand weights of a group landed (stage = group % 3)
    __shared__ __align__(8) uint64_t bar_e[2];       // expand MMAs of a group complete (TMEM buffer = group & 1)
    __shared__ __align__(8) uint64_t bar_d[3];       // projection: D chunk landed (stage = chunk % n_da)
    __shared__ __align__(8) uint64_t bar_q[2];       // projection: weights of a chunk landed (stage = chunk & 1)
    __shared__ __align__(8) uint64_t bar_m[2];       // projection: MMAs of a chunk complete
    // worker warps -> control lane (one arrive per worker warp); the control warp never joins the workers' barriers
    __shared__ __align__(8) uint64_t bar_tf[2];      // E buffer read out of TMEM (buffer = group & 1)
    __shared__ __align__(8) uint64_t bar_g[3];       // projection: D tiles of a chunk gated (stage = chunk % n_da)      [GATEA]
    __shared__ __align__(8) uint64_t bar_wr[2];      // projection: gated weight image of a chunk built (stage = chunk & 1) [!GATEA]
    __shared__ __align__(8) uint64_t bar_dr;         // phase A of the pass done: D stored and fenced
    __shared__ __align__(8) uint64_t bar_sd;         // pass done: TMEM and the staging tile read by the epilogue
    __shared__ uint32_t tmem_holder;

    constexpr int PAD = K / 2;
    constexpr int NP = G / 2;                        // channel pairs per group
    constexpr int NSLOT = MB_THREADS / NP;           // pixel-block slots
    constexpr int NCOL = XB - 1 + K, NROW = YB - 1 + K;
    constexpr int SL = G / 16;                       // 16-channel slices per group
    constexpr int WROWS = K * K + 2;                 // depthwise weight rows + depthwise bias row + expand bias row
    constexpr int NG4 = MB_WARPS / 4;                // warp groups (one warp per TMEM lane quarter in each)
    constexpr int PM = GATEA ? MB_PMAX : 1;          // segments per pass this instantiation can take
    static_assert(MB_WARPS % 4 == 0 && G % 16 == 0, "group = whole 16-channel slices");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_ctl = warp == MB_WARPS;            // control warp: one elected lane issues, the rest only keep the barriers
    const int H = p.h, W = p.w, P = p.segs, npix1 = H * W, npix = P * npix1, wp = W + 2 * PAD, hpwp = (H + 2 * PAD) * wp;
    const int cin = p.cin, cexp = p.cexp, cout = p.cout, R = p.r, R4 = (R + 3) & ~3;
    const MbLayout L = mb_layout(H, W, K, cin, cexp, cout, R, G, P);
    // 1024-byte alignment by pointer arithmetic ON the shared array: a round trip through uintptr_t loses the address
    // space and every access becomes a generic LD / ST (measured: the patch stores alone cost 2.8 k cycles per group)
    uint8_t* sm = mb_smem_raw + ((1024u - (smem_u32(mb_smem_raw) & 1023u)) & 1023u);
    float* s_wd = reinterpret_cast<float*>(sm + L.off_wd);
    float* s_part = reinterpret_cast<float*>(sm + L.off_part);
    float* s_pool = reinterpret_cast<float*>(sm + L.off_pool);
    float* s_gate = reinterpret_cast<float*>(sm + L.off_gate);
    float* s_r = reinterpret_cast<float*>(sm + L.off_r);
    float* s_fc = reinterpret_cast<float*>(sm + L.off_fc);
    const uint32_t wd_u32 = smem_u32(s_wd);
    const int n_grp = cexp / G;
    const int ks_e = (cin + 15) >> 4;                // K steps of the expand GEMM
    const uint32_t e_cols = (uint32_t)L.n_mt * 2u * (uint32_t)G;     // TMEM columns of one E buffer
    const uint32_t p_cols = 2u * (uint32_t)cout;                     // per M tile: projection main | correction
    const bool fused_n = 2 * cout <= 256;
    const uint32_t nda = (uint32_t)L.n_da;

    if (tid == 0) {
        mbar_init(&bar_x, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_w[i], 1); mbar_init(&bar_e[i], 1); mbar_init(&bar_q[i], 1); mbar_init(&bar_m[i], 1);
            mbar_init(&bar_tf[i], MB_WARPS); mbar_init(&bar_wr[i], MB_WARPS);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(&bar_d[i], 1); mbar_init(&bar_g[i], MB_WARPS); }
        mbar_init(&bar_w[2], 1);
        mbar_init(&bar_dr, MB_WARPS);
        mbar_init(&bar_sd, MB_WARPS);
        fence_barrier_init();
        tma_prefetch_desc(&p.xmap);
        tma_prefetch_desc(&p.dmap);
    }
    if (is_ctl) tmem_alloc(&tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    // depthwise thread mapping: channel pair x pixel block (block -> segment of the pass, block row, block column)
    const int cp = tid % NP, blk = tid / NP;
    const int xblocks = W / XB;
    const int nblk1 = (H / YB) * xblocks;            // blocks per segment
    const int bs = blk / nblk1, b1 = blk - bs * nblk1;
    const int by = b1 / xblocks, oy0 = by * YB, ox0 = (b1 - by * xblocks) * XB;
    const bool dw_mapped = !is_ctl && blk < P * nblk1;
    const float inv_np = 1.0f / (float)npix1;
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    // patch position of this lane's row for each of the warp's (at most 4) TMEM units: fixed for the whole kernel
    int ppix_u[4];
#pragma unroll
    for (int ui = 0; ui < 4; ++ui) {
        const int u = (warp >> 2) + ui * NG4;
        const int pix = min((u / SL) * 128 + q * 32 + lane, npix - 1);
        const int sg = pix / npix1, l = pix - sg * npix1;
        const int y = l / W, x = l - y * W;
        ppix_u[ui] = sg * hpwp + (y + PAD) * wp + x + PAD;
    }

    // development aid (p.prof != nullptr): cycles a worker thread of CTA 0 spends per phase, summed over its passes
    //  [0] pass start  [1] wait E_g  [2] TMEM -> patch + barrier A  [3] depthwise + barrier B  [4] FC2  [5] projection chunk loop
    //  [6] wait last MMAs  [7] epilogue  [8] pooled + fence + barrier  [9] FC1  [10] passes  [11] total
    const bool prof = p.prof != nullptr && blockIdx.x == 0 && tid == 32;
    unsigned long long pc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = prof ? clock64() : 0;
    const long long t_start = t_prev;
    auto tick = [&](int slot) { if (prof) { const long long t = clock64(); pc[slot] += (unsigned long long)(t - t_prev); t_prev = t; } };
    uint32_t gi = 0;        // running group counter: buffers / stages = gi & 1, barrier parity = (gi >> 1) & 1
    uint32_t ck = 0;        // running projection chunk counter
    uint32_t pass_it = 0;

    auto load_we = [&](int g, uint32_t stage) {      // expand weights of group g -> stage (control lane)
        arrive_expect_tx(&bar_w[stage], L.we_stage);
        bulk_copy_g2s(sm + L.off_we + stage * L.we_stage, reinterpret_cast<const uint8_t*>(p.we_pack) + (size_t)g * L.we_stage,
                      L.we_stage, &bar_w[stage]);
    };
    auto issue_expand = [&](uint32_t gcount) {       // MMAs of the group with running index gcount (control lane)
        const uint32_t b = gcount & 1u;
        const uint32_t idesc2 = umma_idesc_f16(128, 2 * G), idesc1 = umma_idesc_f16(128, G);
        for (int m = 0; m < L.n_mt; ++m) {
            const uint32_t acc = tmem_base + b * e_cols + (uint32_t)m * 2u * (uint32_t)G;
            for (int ks = 0; ks < ks_e; ++ks) {
                const int kc = ks >> 2, j = ks & 3;
                const uint32_t a_hi = smem_u32(sm + L.off_xa) + ((uint32_t)(kc * 2 + 0) * (uint32_t)L.n_box + 2u * (uint32_t)m) * BOX_BYTES;
                const uint32_t a_lo = smem_u32(sm + L.off_xa) + ((uint32_t)(kc * 2 + 1) * (uint32_t)L.n_box + 2u * (uint32_t)m) * BOX_BYTES;
                const uint32_t wb = smem_u32(sm + L.off_we) + (gcount % 3u) * L.we_stage + (uint32_t)kc * 2u * (uint32_t)G * 128u;
                const uint64_t db = umma_desc_sw128(wb) + (uint64_t)(kDescKStep * j);
                umma_f16(acc, umma_desc_sw128(a_hi) + (uint64_t)(kDescKStep * j), db, idesc2, ks != 0 ? 1u : 0u);
                umma_f16(acc + (uint32_t)G, umma_desc_sw128(a_lo) + (uint64_t)(kDescKStep * j), db, idesc1, 1u);
            }
        }
        umma_commit(&bar_e[b]);
    };
    // depthwise weights + both biases of group g -> s_wd[buf] (cp.async, 16-byte units; every worker commits a group)
    auto load_wd = [&](int g, uint32_t buf) {
        constexpr int UNITS = WROWS * (G / 4);
        for (int i = tid; i < UNITS; i += MB_THREADS) {
            const int row = i / (G / 4), ch = i - row * (G / 4);
            const float* src = (row < K * K ? p.wd + (size_t)row * cexp : (row == K * K ? p.bd : p.be)) + g * G + ch * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wd_u32 + (uint32_t)((buf * WROWS + row) * G + ch * 4) * 4u), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_d = [&](int seg0, int kc, uint32_t st) {   // D chunk kc of this pass -> stage st (control lane)
        arrive_expect_tx(&bar_d[st], L.da_stage);
        for (int pl = 0; pl < 2; ++pl)
            for (int bx = 0; bx < L.n_box; ++bx)
                tma_load_5d(sm + L.off_da + st * L.da_stage + ((uint32_t)pl * (uint32_t)L.n_box + (uint32_t)bx) * BOX_BYTES, &p.dmap,
                            kc * 64, seg0 * npix1 + bx * 64, 0, 0, pl, &bar_d[st]);
    };
    auto load_wq = [&](int kc, uint32_t st) {            // packed projection weights of chunk kc -> stage st (control lane)
        arrive_expect_tx(&bar_q[st], L.wp_stage);
        bulk_copy_g2s(sm + L.off_wp + st * L.wp_stage, reinterpret_cast<const uint8_t*>(p.wp_pack) + (size_t)kc * L.wp_stage, L.wp_stage, &bar_q[st]);
    };
    // !GATEA: projection weights of chunk kc for this thread's units -> registers (the L2 latency is spent before the waits)
    constexpr int UPT = 3;                           // units (8 K values of one output channel) per worker thread: cout * 8 <= 1152
    auto load_wp = [&](int kc, float4 (&wr)[UPT][2]) {
#pragma unroll
        for (int i = 0; i < UPT; ++i) {
            const int u = tid + i * MB_THREADS;
            const int n = u >> 3, c0 = kc * 64 + (u & 7) * 8;
            if (u < cout * 8 && c0 < cexp) {                 // cexp is a multiple of 8
                const float4* src = reinterpret_cast<const float4*>(p.wpT + (size_t)n * cexp + c0);
                wr[i][0] = __ldg(src);
                wr[i][1] = __ldg(src + 1);
            } else {
                wr[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                wr[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    // ... scaled by the gate -> [W_hi | W_lo] operand image of stage st (row n = output channel, 64 K values)
    auto store_wp = [&](int kc, uint32_t st, const float4 (&wr)[UPT][2]) {
        uint8_t* wst = sm + L.off_wp + st * L.wp_stage;
#pragma unroll
        for (int i = 0; i < UPT; ++i) {
            const int u = tid + i * MB_THREADS;
            if (u < cout * 8) {
                const int n = u >> 3, ku = u & 7;
                const int c0 = kc * 64 + ku * 8;
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (c0 < cexp) {
                    const float4 g0 = *reinterpret_cast<const float4*>(s_gate + c0), g1 = *reinterpret_cast<const float4*>(s_gate + c0 + 4);
                    v[0] = wr[i][0].x * g0.x; v[1] = wr[i][0].y * g0.y; v[2] = wr[i][0].z * g0.z; v[3] = wr[i][0].w * g0.w;
                    v[4] = wr[i][1].x * g1.x; v[5] = wr[i][1].y * g1.y; v[6] = wr[i][1].z * g1.z; v[7] = wr[i][1].w * g1.w;
                }
                uint4 hi, lo;
                split8(v, hi, lo);
                const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)ku);
                *reinterpret_cast<uint4*>(wst + off) = hi;
                *reinterpret_cast<uint4*>(wst + (uint32_t)cout * 128u + off) = lo;
            }
        }
    };
    const uint32_t idescP2 = umma_idesc_f16(128, fused_n ? 2 * cout : cout), idescP1 = umma_idesc_f16(128, cout);
    auto issue_project = [&](int kc, uint32_t st, uint32_t sd) {   // MMAs of projection chunk kc: weight stage st, D stage sd (control lane)
        const int ksn = min(4, (cexp - kc * 64 + 15) >> 4);
        const uint32_t wb = smem_u32(sm + L.off_wp) + st * L.wp_stage;
        for (int m = 0; m < L.n_mt; ++m) {
            const uint32_t acc = tmem_base + (uint32_t)m * p_cols;
            const uint32_t a_hi = smem_u32(sm + L.off_da) + sd * L.da_stage + (2u * (uint32_t)m) * BOX_BYTES;
            const uint32_t a_lo = a_hi + (uint32_t)L.n_box * BOX_BYTES;
            for (int j = 0; j < ksn; ++j) {
                const uint64_t dk = (uint64_t)(kDescKStep * j);
                const uint32_t accf = (kc | j) != 0 ? 1u : 0u;
                if (fused_n) {
                    umma_f16(acc, umma_desc_sw128(a_hi) + dk, umma_desc_sw128(wb) + dk, idescP2, accf);
                } else {
                    umma_f16(acc, umma_desc_sw128(a_hi) + dk, umma_desc_sw128(wb) + dk, idescP1, accf);
                    umma_f16(acc + (uint32_t)cout, umma_desc_sw128(a_hi) + dk, umma_desc_sw128(wb + (uint32_t)cout * 128u) + dk, idescP1, accf);
                }
                umma_f16(acc + (uint32_t)cout, umma_desc_sw128(a_lo) + dk, umma_desc_sw128(wb) + dk, idescP1, 1u);
            }
        }
        umma_commit(&bar_m[st]);
    };
    // pooled means of a finished group from the per-block-slot partials: 4 threads per (segment, channel) add a fixed
    // quarter of the block slots each and combine by shuffles - a fixed order, spread over the last warps
    auto pooled_mean = [&](int g, uint32_t buf) {
        const int first = MB_THREADS - 4 * P * G;                    // 4 * P * G <= 256 worker threads
        if (!is_ctl && tid >= first) {
            const int idx = (tid - first) >> 2, part4 = (tid - first) & 3;
            const int sg = idx / G, ch = idx - sg * G;
            const float* part = s_part + (size_t)buf * NSLOT * G + (size_t)sg * nblk1 * G + ch;
            float sum = 0.f;
            for (int i = part4; i < nblk1; i += 4) sum += part[i * G];
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (part4 == 0) s_pool[sg * cexp + g * G + ch] = sum * inv_np;
        }
    };

    if (is_ctl) {
        // =========================== control warp: TMA, bulk copies, every tcgen05.mma ===========================
        if (elect_one()) {
            // development aid: control-lane cycles of CTA 0 -> prof[12] wait TMEM free, [13] wait weights, [14] issue expand, [15] phase B
            const bool cprof = p.prof != nullptr && blockIdx.x == 0;
            unsigned long long cc[4] = {0, 0, 0, 0};
            long long ct = cprof ? clock64() : 0;
            auto ctick = [&](int i) { if (cprof) { const long long t = clock64(); cc[i] += (unsigned long long)(t - ct); ct = t; } };
            for (int pass = blockIdx.x; pass * P < p.batch; pass += gridDim.x, ++pass_it) {
                const int seg0 = pass * P;
                // previous pass: its epilogue has read TMEM and the staging tile (bar_sd), all its MMAs have retired
                if (pass_it > 0) mbar_wait(&bar_sd, (pass_it - 1u) & 1u);
                tc_fence_after();
                arrive_expect_tx(&bar_x, L.xa_bytes);
                for (int kc = 0; kc < L.kc_e; ++kc)
                    for (int pl = 0; pl < 2; ++pl)
                        for (int bx = 0; bx < L.n_box; ++bx)
                            tma_load_5d(sm + L.off_xa + ((uint32_t)(kc * 2 + pl) * (uint32_t)L.n_box + (uint32_t)bx) * BOX_BYTES, &p.xmap,
                                        kc * 64, seg0 * npix1 + bx * 64, 0, 0, pl, &bar_x);
                for (int i = 0; i < 3 && i < n_grp; ++i) load_we(i, (gi + (uint32_t)i) % 3u);
                mbar_wait(&bar_x, pass_it & 1u);
                for (int g = 0; g < n_grp; ++g, ++gi) {
                    const uint32_t b = gi & 1u;
                    // TMEM buffer b must have been read out by the workers (group gi - 2) before it is overwritten
                    // (debug bit 64: wait for group gi - 1 instead, i.e. no MMA in flight while the workers read TMEM)
                    ctick(3);
                    if (p.debug & 64) { if (gi >= 1u) mbar_wait(&bar_tf[b ^ 1u], ((gi - 1u) >> 1) & 1u); }
                    else if (gi >= 2u) mbar_wait(&bar_tf[b], ((gi - 2u) >> 1) & 1u);
                    ctick(0);
                    mbar_wait(&bar_w[gi % 3u], (gi / 3u) & 1u);
                    ctick(1);
                    tc_fence_after();
                    issue_expand(gi);
                    ctick(2);
                    // the weight stage of group gi - 1 is free once its MMAs have retired: fetch group g + 2 into it, a whole
                    // group period before it is needed
                    if (g >= 1 && g + 2 < n_grp) {
                        mbar_wait(&bar_e[b ^ 1u], ((gi - 1u) >> 1) & 1u);
                        load_we(g + 2, (gi - 1u) % 3u);
                    }
                }
                // every expand MMA retired -> the operand region may take the projection's stages
                mbar_wait(&bar_e[(gi - 1u) & 1u], ((gi - 1u) >> 1) & 1u);
                if (n_grp > 1) mbar_wait(&bar_e[gi & 1u], ((gi - 2u) >> 1) & 1u);
                mbar_wait(&bar_dr, pass_it & 1u);                    // D stored and fenced, patch no longer read, by every worker
                if (GATEA) for (int i = 0; i < 2 && i < L.kc_p; ++i) load_wq(i, (ck + (uint32_t)i) & 1u);
                for (uint32_t i = 0; i < nda && (int)i < L.kc_p; ++i) load_d(seg0, (int)i, (ck + i) % nda);
                for (int kc = 0; kc < L.kc_p; ++kc, ++ck) {
                    const uint32_t st = ck & 1u, sd = ck % nda;
                    if (GATEA) {
                        mbar_wait(&bar_g[sd], (ck / nda) & 1u);      // D tiles of chunk kc landed and gated by the workers
                        mbar_wait(&bar_q[st], (ck >> 1) & 1u);       // weights of chunk kc landed
                    } else {
                        mbar_wait(&bar_wr[st], (ck >> 1) & 1u);      // gated weight image of chunk kc built by the workers
                        mbar_wait(&bar_d[sd], (ck / nda) & 1u);      // D tiles of chunk kc landed
                    }
                    tc_fence_after();
                    issue_project(kc, st, sd);
                    // chunk kc - 1 retired -> its stages take the D tiles of chunk kc - 1 + n_da and the weights of chunk kc + 1
                    if (kc >= 1) {
                        mbar_wait(&bar_m[st ^ 1u], ((ck - 1u) >> 1) & 1u);
                        if (kc - 1 + (int)nda < L.kc_p) load_d(seg0, kc - 1 + (int)nda, (ck - 1u) % nda);
                        if (GATEA && kc + 1 < L.kc_p) load_wq(kc + 1, st ^ 1u);
                    }
                }
            }
            if (cprof) for (int i = 0; i < 4; ++i) p.prof[12 + i] = cc[i];
        }
        __syncwarp();
    } else {
    // ================================================ workers ================================================
    for (int pass = blockIdx.x; pass * P < p.batch; pass += gridDim.x, ++pass_it) {
        const int seg0 = pass * P;
        const int nseg = min(P, p.batch - seg0);                     // the last pass may hold fewer segments
        const int rows_ok = nseg * npix1;
        // ================================ phase A ================================
        load_wd(0, gi & 1u);
        // the projection stages and the epilogue tile of the previous pass lay over the patch: zero it (halo included)
        for (uint32_t i = tid; i < L.patch_bytes / 16u; i += MB_THREADS)
            reinterpret_cast<uint4*>(sm + L.off_patch)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        worker_sync();                                               // patch zeroed, group 0's depthwise weights + biases visible
        tick(0);
        for (int g = 0; g < n_grp; ++g, ++gi) {
            const uint32_t b = gi & 1u;
            if (g + 1 < n_grp) load_wd(g + 1, b ^ 1u);               // next group's depthwise weights + biases
            const float* wgrp = s_wd + (size_t)b * WROWS * G;
            tick(0);
            mbar_wait(&bar_e[b], (gi >> 1) & 1u);
            tc_fence_after();
            tick(1);
            // TMEM -> + expand bias, SiLU -> patch.  warp (q, unit): rows m*128 + q*32 + lane, channels slice*16 .. +15
#pragma unroll
            for (int ui = 0; ui < 4; ++ui) {
                const int u = (warp >> 2) + ui * NG4;
                if (ui * NG4 >= 2 * SL) break;                       // compile-time bound: n_mt <= 2
                const int m = u / SL, sl = u - m * SL;
                const int pix = m * 128 + q * 32 + lane;
                if (u < L.n_mt * SL && m * 128 + q * 32 < npix) {    // warp-uniform: a real unit whose quarter holds real rows
                    uint32_t rm[16], rc[16];
                    const uint32_t t = tmem_base + b * e_cols + (uint32_t)m * 2u * (uint32_t)G + (uint32_t)(sl * 16) + ((uint32_t)(q * 32) << 16);
                    tmem_ld16_nowait(t, rm);
                    tmem_ld16_nowait(t + (uint32_t)G, rc);
                    tmem_ld_wait();
                    if (pix < npix) {
                        const int ppix = ppix_u[ui];
                        const float2* be = reinterpret_cast<const float2*>(wgrp + (K * K + 1) * G + sl * 16);
                        float2* dst = reinterpret_cast<float2*>(sm + L.off_patch) + (size_t)(sl * 8) * L.npixp + ppix;
                        float2 bq[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) bq[j] = be[j];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v0 = silu_f(__uint_as_float(rm[2 * j]) + __uint_as_float(rc[2 * j]) + bq[j].x);
                            const float v1 = silu_f(__uint_as_float(rm[2 * j + 1]) + __uint_as_float(rc[2 * j + 1]) + bq[j].y);
                            dst[(size_t)j * L.npixp] = make_float2(v0, v1);
                        }
                    }
                }
            }
            // this warp has read its share of E_g out of TMEM: tell the control lane (buffer b may be overwritten)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tf[b]);
            worker_sync();                                           // A: patch complete
            tick(2);
            if (g > 0) pooled_mean(g - 1, b ^ 1u);
            // depthwise conv + SiLU -> D planes (un-gated), pooled partial sums
            float2 pool = make_float2(0.f, 0.f);
            if (dw_mapped && bs < nseg) {
                const int c = g * G + 2 * cp;
                // this thread's depthwise weights (visible since the previous barrier B)
                unsigned long long wk[K * K];
                {
                    const float* wb = wgrp + 2 * cp;
#pragma unroll
                    for (int i = 0; i < K * K; ++i) wk[i] = *reinterpret_cast<const unsigned long long*>(wb + i * G);
                }
                const unsigned long long bias = *reinterpret_cast<const unsigned long long*>(wgrp + K * K * G + 2 * cp);
                unsigned long long acc[YB][XB];
#pragma unroll
                for (int y = 0; y < YB; ++y)
#pragma unroll
                    for (int x = 0; x < XB; ++x) acc[y][x] = bias;
                const unsigned long long* base = reinterpret_cast<const unsigned long long*>(sm + L.off_patch) + (size_t)cp * L.npixp + bs * hpwp + oy0 * wp + ox0;
#pragma unroll
                for (int r = 0; r < NROW; ++r) {
#pragma unroll
                    for (int x = 0; x < NCOL; ++x) {
                        const unsigned long long v = base[r * wp + x];
#pragma unroll
                        for (int y = 0; y < YB; ++y) {
                            const int ky = r - y;
                            if (ky < 0 || ky >= K) continue;
#pragma unroll
                            for (int j = 0; j < XB; ++j) {
                                const int kx = x - j;
                                if (kx < 0 || kx >= K) continue;
                                acc[y][j] = ffma2(v, wk[ky * K + kx], acc[y][j]);
                            }
                        }
                    }
                }
                __half* dh = p.d_hi + ((size_t)(seg0 + bs) * npix1) * cexp + c;
#pragma unroll
                for (int y = 0; y < YB; ++y)
#pragma unroll
                    for (int j = 0; j < XB; ++j) {
                        const float v0 = silu_f(lo_f(acc[y][j])), v1 = silu_f(hi_f(acc[y][j]));
                        pool.x += v0;
                        pool.y += v1;
                        const __half2 hh = __floats2half2_rn(v0, v1);
                        const float2 bk = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(v0 - bk.x, v1 - bk.y);
                        const size_t o = (size_t)((oy0 + y) * W + ox0 + j) * cexp;
                        *reinterpret_cast<__half2*>(dh + o) = hh;
                        *reinterpret_cast<__half2*>(dh + p.d_plane + o) = ll;
                    }
            }
            // pooled partial of this thread's block, one slot per block: summed per segment in a fixed order later
            reinterpret_cast<float2*>(s_part + (size_t)b * NSLOT * G)[blk * NP + cp] = pool;
            asm volatile("cp.async.wait_group 0;" ::: "memory");     // next group's weights have landed (this thread's copies)
            worker_sync();                                           // B: patch free, partials + next weights visible
            tick(3);
        }
        pooled_mean(n_grp - 1, (gi - 1u) & 1u);
        fence_proxy_async_all();                                     // this thread's D stores -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_dr);                         // control: D of this pass may be fetched
        float4 wr[UPT][2];
        if (!GATEA) load_wp(0, wr);
        worker_sync();                                               // pooled vectors complete
        tick(8);

        // ================================ gate (per segment of the pass) ================================
        {
            // FC1: r = silu(W1^T pooled + b1).  A warp takes a slice of the channels; 8 lanes cover one row of W1 with
            // float4 loads (R <= 32 of its ldw1 floats per sweep), 4 rows per step, up to 64 rows of the slice in flight.
            const int cpw = (cexp + MB_WARPS - 1) / MB_WARPS;
            const int cbeg = warp * cpw, cend = min(cexp, cbeg + cpw);
            const int l8 = lane & 7, rsub = lane >> 3;
            for (int j0 = 0; j0 < R; j0 += 32) {
                float4 a4[PM];
#pragma unroll
                for (int sg = 0; sg < PM; ++sg) a4[sg] = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool col_ok = j0 + 4 * l8 < p.ldw1;            // ldw1 is R rounded up to 4: the pad columns are zero
                for (int c0 = cbeg; c0 < cend; c0 += 64) {
                    float4 wv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int cc = c0 + 4 * i + rsub;
                        wv[i] = (col_ok && cc < cend) ? __ldg(reinterpret_cast<const float4*>(p.w1 + (size_t)cc * p.ldw1 + j0) + l8) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int cc = c0 + 4 * i + rsub;
#pragma unroll
                        for (int sg = 0; sg < PM; ++sg) {
                            const float pv = (cc < cend && sg < P) ? s_pool[sg * cexp + cc] : 0.f;
                            a4[sg].x = fmaf(pv, wv[i].x, a4[sg].x); a4[sg].y = fmaf(pv, wv[i].y, a4[sg].y);
                            a4[sg].z = fmaf(pv, wv[i].z, a4[sg].z); a4[sg].w = fmaf(pv, wv[i].w, a4[sg].w);
                        }
                    }
                }
#pragma unroll
                for (int sg = 0; sg < PM; ++sg) {
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {              // add the 4 row lanes (fixed order)
                        a4[sg].x += __shfl_xor_sync(0xffffffffu, a4[sg].x, o); a4[sg].y += __shfl_xor_sync(0xffffffffu, a4[sg].y, o);
                        a4[sg].z += __shfl_xor_sync(0xffffffffu, a4[sg].z, o); a4[sg].w += __shfl_xor_sync(0xffffffffu, a4[sg].w, o);
                    }
                    if (rsub == 0 && sg < P) {
                        const int j = j0 + 4 * l8;
                        float* dst = s_fc + (size_t)(warp * P + sg) * R;
                        if (j + 0 < R) dst[j + 0] = a4[sg].x;
                        if (j + 1 < R) dst[j + 1] = a4[sg].y;
                        if (j + 2 < R) dst[j + 2] = a4[sg].z;
                        if (j + 3 < R) dst[j + 3] = a4[sg].w;
                    }
                }
            }
            worker_sync();
            for (int i = tid; i < P * R; i += MB_THREADS) {
                const int sg = i / R, j = i - sg * R;
                float v = __ldg(p.b1 + j);
#pragma unroll
                for (int wi = 0; wi < MB_WARPS; ++wi) v += s_fc[(size_t)(wi * P + sg) * R + j];
                s_r[sg * R4 + j] = v * (1.0f / (1.0f + expf(-v)));
            }
            worker_sync();
            tick(9);
            // FC2: g = sigmoid(W2^T r + b2): a thread takes channels tid and tid + 384 together (rows of W2 are contiguous in
            // c: coalesced), 2 x 12 loads in flight per step, each weight used for every segment of the pass
            for (int cb = 0; cb < cexp; cb += 2 * MB_THREADS) {
                const int c0 = cb + tid, c1 = cb + tid + MB_THREADS;
                const bool ok0 = c0 < cexp, ok1 = c1 < cexp;
                float v0[PM], v1[PM];
                const float b20 = ok0 ? __ldg(p.b2 + c0) : 0.f, b21 = ok1 ? __ldg(p.b2 + c1) : 0.f;
#pragma unroll
                for (int sg = 0; sg < PM; ++sg) { v0[sg] = b20; v1[sg] = b21; }
                for (int j = 0; j < R; j += 12) {
                    float w0[12], w1v[12];
#pragma unroll
                    for (int u = 0; u < 12; ++u) {
                        const bool jo = j + u < R;
                        w0[u] = (jo && ok0) ? __ldg(p.w2 + (size_t)(j + u) * p.ldw2 + c0) : 0.f;
                        w1v[u] = (jo && ok1) ? __ldg(p.w2 + (size_t)(j + u) * p.ldw2 + c1) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 12; ++u) {
#pragma unroll
                        for (int sg = 0; sg < PM; ++sg) {
                            const float rv = (j + u < R && sg < P) ? s_r[sg * R4 + j + u] : 0.f;
                            v0[sg] = fmaf(rv, w0[u], v0[sg]);
                            v1[sg] = fmaf(rv, w1v[u], v1[sg]);
                        }
                    }
                }
#pragma unroll
                for (int sg = 0; sg < PM; ++sg) {
                    if (sg < P && ok0) s_gate[sg * cexp + c0] = 1.0f / (1.0f + expf(-v0[sg]));
                    if (sg < P && ok1) s_gate[sg * cexp + c1] = 1.0f / (1.0f + expf(-v1[sg]));
                }
            }
            worker_sync();
        }
        tick(4);

        // ================================ phase B: Y = (D * g) * Wp ================================
        // chunk kc: D stage sd = ck % n_da.  The workers multiply the landed D tiles by the gate in place (x = hi + lo,
        // x * g[segment][channel], split again); the control lane issues the chunk once tiles and weights are ready.
        if (!GATEA) {
            // the workers build the gated weight image of chunk kc while the MMAs of chunk kc - 1 run
            for (int kc = 0; kc < L.kc_p; ++kc, ++ck) {
                const uint32_t st = ck & 1u;
                if (kc >= 2) mbar_wait(&bar_m[st], ((ck - 2u) >> 1) & 1u);   // the stage's previous chunk has retired
                store_wp(kc, st, wr);
                if (kc + 1 < L.kc_p) load_wp(kc + 1, wr);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_wr[st]);
            }
        } else
        for (int kc = 0; kc < L.kc_p; ++kc, ++ck) {
            const uint32_t sd = ck % nda;
            mbar_wait(&bar_d[sd], (ck / nda) & 1u);
            uint8_t* tile = sm + L.off_da + sd * L.da_stage;
            const uint32_t lo_off = (uint32_t)L.n_box * BOX_BYTES;
            for (int u = tid; u < rows_ok * 8; u += MB_THREADS) {
                const int row = u >> 3, ku = u & 7;
                const int c0 = kc * 64 + ku * 8;
                if (c0 >= cexp) continue;                            // zero-filled tail of the last chunk
                const int sg = row / npix1;
                const uint32_t off = (uint32_t)(row >> 6) * BOX_BYTES + sw128_offset((uint32_t)(row & 63), (uint32_t)ku);
                const uint4 qh = *reinterpret_cast<const uint4*>(tile + off);
                const uint4 ql = *reinterpret_cast<const uint4*>(tile + lo_off + off);
                const float4 g0 = *reinterpret_cast<const float4*>(s_gate + sg * cexp + c0);
                const float4 g1 = *reinterpret_cast<const float4*>(s_gate + sg * cexp + c0 + 4);
                const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const __half2* h2 = reinterpret_cast<const __half2*>(&qh);
                const __half2* l2 = reinterpret_cast<const __half2*>(&ql);
                float v[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = __half22float2(h2[e]), d = __half22float2(l2[e]);
                    v[2 * e] = (a.x + d.x) * gv[2 * e];
                    v[2 * e + 1] = (a.y + d.y) * gv[2 * e + 1];
                }
                uint4 hi, lo;
                split8(v, hi, lo);
                *reinterpret_cast<uint4*>(tile + off) = hi;
                *reinterpret_cast<uint4*>(tile + lo_off + off) = lo;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_g[sd]);
        }
        tick(5);
        // every chunk's MMAs complete (commits retire in order; both stages are waited so the parities stay in step)
        {
            const uint32_t last = ck - 1u;
            if (L.kc_p >= 2) { const uint32_t prev = ck - 2u; mbar_wait(&bar_m[prev & 1u], (prev >> 1) & 1u); }
            mbar_wait(&bar_m[last & 1u], (last >> 1) & 1u);
            tc_fence_after();
        }
        tick(6);
        // epilogue, pass 1: main + correction + bias -> FP32 staging tile [rows][cout] in the (now idle) operand region.
        // warp = (q, unit = M tile x 16-column slice), lane = row; row pitch + 16 B keeps the 16-byte lane stores conflict-free
        {
            const int n_sl = cout >> 4;
            for (int u = warp >> 2; u < L.n_mt * n_sl; u += NG4) {
                const int m = u / n_sl, sl = u - m * n_sl;
                if (m * 128 + q * 32 >= rows_ok) continue;           // warp-uniform
                const int pix = m * 128 + q * 32 + lane;
                const int n0 = sl * 16;
                uint32_t rm[16], rc[16];
                const uint32_t t = tmem_base + (uint32_t)m * p_cols + (uint32_t)(sl * 16) + ((uint32_t)(q * 32) << 16);
                tmem_ld16_nowait(t, rm);
                tmem_ld16_nowait(t + (uint32_t)cout, rc);
                tmem_ld_wait();
                if (pix < rows_ok) {
                    float4* dst = reinterpret_cast<float4*>(sm + (size_t)pix * L.stg_pitch + (size_t)n0 * 4);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bp + n0) + j4);
                        dst[j4] = make_float4(__uint_as_float(rm[4 * j4 + 0]) + __uint_as_float(rc[4 * j4 + 0]) + bv.x,
                                              __uint_as_float(rm[4 * j4 + 1]) + __uint_as_float(rc[4 * j4 + 1]) + bv.y,
                                              __uint_as_float(rm[4 * j4 + 2]) + __uint_as_float(rc[4 * j4 + 2]) + bv.z,
                                              __uint_as_float(rm[4 * j4 + 3]) + __uint_as_float(rc[4 * j4 + 3]) + bv.w);
                    }
                }
            }
        }
        tc_fence_before();
        worker_sync();
        // pass 2: (+ residual) -> hi/lo split -> planes, 8 channels (16 bytes per plane) per thread, rows contiguous: coalesced
        {
            const int upr = cout >> 3;                               // 8-channel units per row
            const int total = rows_ok * upr;
            for (int u = tid; u < total; u += MB_THREADS) {
                const int pix = u / upr, n0 = (u - pix * upr) * 8;
                const float4* src = reinterpret_cast<const float4*>(sm + (size_t)pix * L.stg_pitch + (size_t)n0 * 4);
                const float4 a = src[0], bq4 = src[1];
                float v[8] = {a.x, a.y, a.z, a.w, bq4.x, bq4.y, bq4.z, bq4.w};
                const size_t o = ((size_t)seg0 * npix1 + pix) * cout + n0;
                if (p.res_hi) {
                    const uint4 qh = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o));
                    const uint4 ql = __ldg(reinterpret_cast<const uint4*>(p.res_hi + p.res_plane + o));
                    const __half2* h2 = reinterpret_cast<const __half2*>(&qh);
                    const __half2* l2 = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 x = __half22float2(h2[e]), d = __half22float2(l2[e]);
                        v[2 * e] += x.x + d.x;
                        v[2 * e + 1] += x.y + d.y;
                    }
                }
                uint4 hq, lq;
                split8(v, hq, lq);
                *reinterpret_cast<uint4*>(p.out_hi + o) = hq;
                *reinterpret_cast<uint4*>(p.out_hi + p.out_plane + o) = lq;
            }
        }
        // this warp is done with TMEM, the staging tile and the pass: the control lane may start the next one
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_sd);
        worker_sync();                                               // nobody zeroes the patch while a slower warp still reads the tile
        tick(7);
        if (prof) pc[10] += 1;
    }
    }
    if (prof) {
        pc[11] = (unsigned long long)(clock64() - t_start);
        for (int i = 0; i < 12; ++i) p.prof[i] = pc[i];
    }
    tc_fence_before();
    __syncthreads();
    if (is_ctl) tmem_dealloc(tmem_base, 512);
}

namespace {

constexpr size_t MB_SMEM_MAX = 226 * 1024;      // dynamic part: the 227 KB per-CTA limit includes the ~200 B of static barriers

struct MbVariant { int k, g, xb, yb, segs; };

// segments per CTA pass: small images are taken two at a time
int mb_segs(int h, int w) { return h * w < 128 ? 2 : 1; }

bool mb_fits(int h, int w, int k, int cin, int cexp, int cout, int r, int g, int segs) {
    if ((cin & 7) || (cexp % g) || (cexp & 7) || (cout & 15) || cout > 256 || r < 1 || r > 256) return false;
    if (segs * h * w > 256) return false;                                    // two 128-row MMA tiles per pass
    const MbLayout L = mb_layout(h, w, k, cin, cexp, cout, r, g, segs);
    if (L.total > MB_SMEM_MAX) return false;
    if ((uint32_t)L.n_mt * 2u * (uint32_t)g * 2u > 512u) return false;       // two E buffers in TMEM
    if ((uint32_t)L.n_mt * 2u * (uint32_t)cout > 512u) return false;         // projection accumulators in TMEM
    return true;
}

// the channel group and pixel block this build instantiates for a block (stride 1 only), or g = 0:
// 384 worker threads = G/2 channel pairs x block slots, every block of the pass's segments in its own slot, and the
// shared-memory / TMEM budget of the layout must hold
MbVariant mb_pick(int h, int w, int k, int cin, int cexp, int cout, int r) {
    if (k != 3 && k != 5) return MbVariant{0, 0, 0, 0, 0};
    const int segs = mb_segs(h, w);
    constexpr int NC = 6;
    const int cand[NC][3] = {{32, 4, 2}, {64, 8, 1}, {32, 4, 1}, {64, 4, 1}, {32, 2, 2}, {64, 2, 1}};     // G, XB, YB
    int best = -1;
    double best_u = 0.0;
    for (int i = 0; i < NC; ++i) {
        const int g = cand[i][0], xb = cand[i][1], yb = cand[i][2];
        if (w % xb || h % yb) continue;
        const int slots = MB_THREADS / (g / 2), nblk = segs * (h / yb) * (w / xb);
        if (nblk > slots || !mb_fits(h, w, k, cin, cexp, cout, r, g, segs)) continue;
        // larger pixel blocks reuse more of each patch read; larger groups mean fewer barriers
        const double u = (double)nblk / slots * (xb * yb >= 8 ? 1.0 : (xb * yb >= 4 ? 0.9 : 0.75)) * (g == 64 ? 1.05 : 1.0);
        if (u > best_u) { best_u = u; best = i; }
    }
    if (best < 0) return MbVariant{0, 0, 0, 0, 0};
    return MbVariant{k, cand[best][0], cand[best][1], cand[best][2], segs};
}

// gate on the weights when one segment is processed per pass and the projection is narrow enough for 3 units per thread
bool mb_gate_on_d(int segs, int cout) { return segs > 1 || cout * 8 > 3 * MB_THREADS; }

template <int K, int G, int XB, int YB>
cudaError_t mb_launch_v(const MbconvParams& p, int grid, size_t smem, cudaStream_t stream) {
    if (mb_gate_on_d(p.segs, p.cout)) k_mbconv<K, G, XB, YB, true><<<grid, MB_BLOCK, smem, stream>>>(p);
    else k_mbconv<K, G, XB, YB, false><<<grid, MB_BLOCK, smem, stream>>>(p);
    return cudaGetLastError();
}
template <int K, int G, int XB, int YB>
cudaError_t mb_attr_v() {
    cudaError_t e = cudaFuncSetAttribute(k_mbconv<K, G, XB, YB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mbconv<K, G, XB, YB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM_MAX);
    return e;
}
template <int K>
cudaError_t mb_attr_k() {
    cudaError_t e = mb_attr_v<K, 32, 4, 2>();
    if (e == cudaSuccess) e = mb_attr_v<K, 64, 8, 1>();
    if (e == cudaSuccess) e = mb_attr_v<K, 32, 4, 1>();
    if (e == cudaSuccess) e = mb_attr_v<K, 64, 4, 1>();
    if (e == cudaSuccess) e = mb_attr_v<K, 32, 2, 2>();
    if (e == cudaSuccess) e = mb_attr_v<K, 64, 2, 1>();
    return e;
}
template <int K>
cudaError_t mb_launch_k(const MbVariant& v, const MbconvParams& p, int grid, size_t smem, cudaStream_t stream) {
    if (v.g == 32 && v.xb == 4 && v.yb == 2) return mb_launch_v<K, 32, 4, 2>(p, grid, smem, stream);
    if (v.g == 32 && v.xb == 4) return mb_launch_v<K, 32, 4, 1>(p, grid, smem, stream);
    if (v.g == 64 && v.xb == 8) return mb_launch_v<K, 64, 8, 1>(p, grid, smem, stream);
    if (v.g == 64 && v.xb == 4) return mb_launch_v<K, 64, 4, 1>(p, grid, smem, stream);
    if (v.g == 32) return mb_launch_v<K, 32, 2, 2>(p, grid, smem, stream);
    return mb_launch_v<K, 64, 2, 1>(p, grid, smem, stream);
}

}  // namespace

int mbconv_group(int h, int w, int k, int cin, int cexp, int cout, int r) { return mb_pick(h, w, k, cin, cexp, cout, r).g; }

bool mbconv_supported(int h, int w, int k, int stride, int cin, int cexp, int cout, int r) {
    if (stride != 1) return false;
    // Images under 128 pixels (two segments per pass, gate applied to D): correct, but measured SLOWER than the layered
    // path on 3x16 (0.175 vs 0.159 ms per block: 36 groups of barriers, FC weights and a 2-stage D ring per pass) - off
    // unless BN_MBCONV_SMALL=1
    static const bool small_on = [] { const char* ev = getenv("BN_MBCONV_SMALL"); return ev && ev[0] == '1'; }();
    if (h * w < 128 && !small_on) return false;
    return mb_pick(h, w, k, cin, cexp, cout, r).g != 0;
}

cudaError_t mbconv_init_device() {
    cudaError_t e = mb_attr_k<3>();
    if (e == cudaSuccess) e = mb_attr_k<5>();
    return e;
}

cudaError_t launch_mbconv(const MbconvParams& pin, int num_sms, cudaStream_t stream) {
    if (pin.batch <= 0) return cudaSuccess;
    if (!mbconv_supported(pin.h, pin.w, pin.k, 1, pin.cin, pin.cexp, pin.cout, pin.r)) return cudaErrorInvalidValue;
    const MbVariant v = mb_pick(pin.h, pin.w, pin.k, pin.cin, pin.cexp, pin.cout, pin.r);
    MbconvParams p = pin;
    p.segs = v.segs;
    const MbLayout L = mb_layout(p.h, p.w, p.k, p.cin, p.cexp, p.cout, p.r, v.g, v.segs);
    const int passes = (p.batch + v.segs - 1) / v.segs;
    const int grid = passes < num_sms ? passes : num_sms;
    if (v.k == 3) return mb_launch_k<3>(v, p, grid, L.total, stream);
    return mb_launch_k<5>(v, p, grid, L.total, stream);
}

}  // namespace bn
