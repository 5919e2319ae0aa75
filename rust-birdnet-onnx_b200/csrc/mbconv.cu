// Fused MBConv block (rows A8 of SURVEY.md section 8): 1x1 expand + SiLU -> depthwise kxk + SiLU -> squeeze-excite
// -> gated 1x1 projection (+ residual), ONE kernel per block, one CTA per segment at a time.
//
//   E = silu(X * We + be)                    [npix][cexp]   tcgen05 into TMEM, never leaves the SM
//   D = silu(dw_kxk(E) + bd)                 [npix][cexp]   CUDA cores, E read from a shared-memory patch
//   g = sigmoid(W2^T silu(W1^T mean_p(D) + b1) + b2)        [cexp]
//   Y = (D * g) * Wp + bp (+ X)              [npix][cout]   tcgen05; the gate is folded into the WEIGHT operand
//
// What used to be three launches (expand conv writing E as FP32, depthwise + SE kernel re-reading it and rescaling D
// in place, projection conv) with E and D round-tripping through HBM is now: X read once (TMA), D written once and
// read back once as the projection's A operand (it is needed only after the gate, which depends on all of D; 0.2-0.5 MB
// per segment, written microseconds earlier by the same SM: it is served by L2), Y written once.
//
// Phase A, per group of G expanded channels (double-buffered in TMEM, so the MMAs of group g+1 run under the
// depthwise arithmetic of group g):
//   one elected thread issues   E_g[npix][G] = X[npix][cin] * We[cin][G]      (hi/lo operands: 3 MMAs per K step, main |
//                                                                             correction accumulators as in tc_conv.cu)
//   all 16 warps                TMEM -> +bias, SiLU -> FP32 patch [G/2 channel pairs][zero-halo pixels] in smem
//   all 512 threads             depthwise conv from the patch (thread = channel pair x XB x YB output block, packed
//                               FFMA2), SiLU, D -> global hi/lo planes, pooled partial sums (fixed order: deterministic)
// Gate: the two tiny FCs by the whole CTA.
// Phase B, per 64-channel K chunk (two-stage ring): TMA lands D's hi/lo tiles, all threads build the chunk of
//   Wp^T * diag(g) as the SWIZZLE_128B [W_hi | W_lo] operand image, one thread issues the MMAs; epilogue adds bias and
//   the residual and writes the block's output planes.
#include "mbconv.h"
#include "tc_common.cuh"

#include <cstdio>
#include <cstdlib>

namespace bn {
using namespace tc;

namespace {

constexpr int MB_THREADS = 512;
constexpr int MB_WARPS = MB_THREADS / 32;
constexpr uint32_t BOX_BYTES = 64 * 128;      // one TMA box: 64 rows x 64 fp16 channels, SWIZZLE_128B

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

}  // namespace

// shared-memory carve-up (bytes from the 1024-aligned base); host and device agree through this function
__host__ __device__ inline MbLayout mb_layout(int h, int w, int k, int cin, int cexp, int cout, int r, int G) {
    MbLayout L;
    const int npix = h * w, pad = k / 2;
    L.n_box = (npix + 63) / 64;
    L.n_mt = (npix + 127) / 128;
    L.kc_e = (cin + 63) / 64;
    L.kc_p = (cexp + 63) / 64;
    L.npixp = ((h + 2 * pad) * (w + 2 * pad) + 15) / 16 * 16 + 1;      // = 1 mod 16: channel pairs land on distinct banks
    L.xa_bytes = (uint32_t)L.kc_e * 2u * (uint32_t)L.n_box * BOX_BYTES;
    L.we_stage = (uint32_t)L.kc_e * 2u * (uint32_t)G * 128u;
    L.da_stage = 2u * (uint32_t)L.n_box * BOX_BYTES;
    L.wp_stage = 2u * (uint32_t)cout * 128u;
    L.off_xa = 0;
    L.off_we = L.xa_bytes;
    const uint32_t a_end = L.xa_bytes + 2u * L.we_stage + BOX_BYTES;     // + one box: the last M tile may read past its 64 rows
    L.off_da = 0;
    L.off_wp = 2u * L.da_stage;
    const uint32_t b_end = 2u * (L.da_stage + L.wp_stage) + BOX_BYTES;
    uint32_t o = a_end > b_end ? a_end : b_end;
    o = (o + 1023u) & ~1023u;
    L.off_patch = o;
    L.patch_bytes = (uint32_t)(G / 2) * (uint32_t)L.npixp * 8u;
    o += (L.patch_bytes + 15u) & ~15u;
    L.off_part = o;  o += 2u * (uint32_t)(MB_THREADS / (G / 2)) * (uint32_t)G * 4u;     // pooled partials, double buffered
    L.off_pool = o;  o += (uint32_t)cexp * 4u;
    L.off_gate = o;  o += (uint32_t)cexp * 4u;
    L.off_r = o;     o += (uint32_t)((r + 3) & ~3) * 4u;
    L.off_fc = o;    o += (uint32_t)MB_WARPS * (uint32_t)r * 4u;
    L.total = o + 1024u;                                                 // alignment slack
    return L;
}

template <int K, int G, int XB, int YB>
__global__ void __launch_bounds__(MB_THREADS, 1) k_mbconv(const __grid_constant__ MbconvParams p) {
    extern __shared__ __align__(1024) uint8_t mb_smem_raw[];
    __shared__ __align__(8) uint64_t bar_x;          // block input landed
    __shared__ __align__(8) uint64_t bar_w[2];       // expand weights of a group landed (stage = group & 1)
    __shared__ __align__(8) uint64_t bar_e[2];       // expand MMAs of a group complete (TMEM buffer = group & 1)
    __shared__ __align__(8) uint64_t bar_d[2];       // projection: D chunk landed (stage = chunk & 1)
    __shared__ __align__(8) uint64_t bar_m[2];       // projection: MMAs of a chunk complete
    __shared__ uint32_t tmem_holder;

    constexpr int PAD = K / 2;
    constexpr int NP = G / 2;                        // channel pairs per group
    constexpr int NB = MB_THREADS / NP;              // pixel-block slots
    constexpr int NCOL = XB - 1 + K, NROW = YB - 1 + K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.h, W = p.w, npix = H * W, wp = W + 2 * PAD;
    const int cin = p.cin, cexp = p.cexp, cout = p.cout, R = p.r;
    const MbLayout L = mb_layout(H, W, K, cin, cexp, cout, R, G);
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mb_smem_raw) + 1023) & ~(uintptr_t)1023);
    float* s_part = reinterpret_cast<float*>(sm + L.off_part);
    float* s_pool = reinterpret_cast<float*>(sm + L.off_pool);
    float* s_gate = reinterpret_cast<float*>(sm + L.off_gate);
    float* s_r = reinterpret_cast<float*>(sm + L.off_r);
    float* s_fc = reinterpret_cast<float*>(sm + L.off_fc);
    const int n_grp = cexp / G;
    const int ks_e = (cin + 15) >> 4;                // K steps of the expand GEMM
    const uint32_t e_cols = (uint32_t)L.n_mt * 2u * (uint32_t)G;     // TMEM columns of one E buffer
    const uint32_t p_cols = 2u * (uint32_t)cout;                     // per M tile: projection main | correction
    const bool fused_n = 2 * cout <= 256;

    if (tid == 0) {
        mbar_init(&bar_x, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_w[i], 1); mbar_init(&bar_e[i], 1); mbar_init(&bar_d[i], 1); mbar_init(&bar_m[i], 1); }
        fence_barrier_init();
        tma_prefetch_desc(&p.xmap);
        tma_prefetch_desc(&p.dmap);
    }
    if (warp == 1) tmem_alloc(&tmem_holder, 512);
    // the patch halo is never written again: zero the whole patch once
    for (uint32_t i = tid; i < L.patch_bytes / 16u; i += MB_THREADS)
        reinterpret_cast<uint4*>(sm + L.off_patch)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    // depthwise thread mapping: channel pair x pixel block
    const int cp = tid % NP, blk = tid / NP;
    const int xblocks = W / XB;
    const int nblk = (H / YB) * xblocks;
    const bool dw_active = blk < nblk;
    const int by = blk / xblocks, oy0 = by * YB, ox0 = (blk - by * xblocks) * XB;
    const float inv_np = 1.0f / (float)npix;
    // TMEM -> patch mapping: warp = (lane quarter q, M tile m, 16-channel slice of the group)
    const int q = warp & 3;
    constexpr int SL = G / 16;                       // 16-channel slices per group
    static_assert(MB_WARPS % 4 == 0 && G % 16 == 0, "group = whole 16-channel slices");

    uint32_t gi = 0;        // running group counter: buffers / stages = gi & 1, barrier parity = (gi >> 1) & 1
    uint32_t ck = 0;        // running projection chunk counter
    uint32_t seg_it = 0;

    auto load_we = [&](int g, uint32_t stage) {      // expand weights of group g -> stage (one elected thread)
        arrive_expect_tx(&bar_w[stage], L.we_stage);
        bulk_copy_g2s(sm + L.off_we + stage * L.we_stage, reinterpret_cast<const uint8_t*>(p.we_pack) + (size_t)g * L.we_stage,
                      L.we_stage, &bar_w[stage]);
    };
    auto issue_expand = [&](uint32_t gcount) {       // MMAs of the group with running index gcount (one elected thread)
        const uint32_t b = gcount & 1u;
        const uint32_t idesc2 = umma_idesc_f16(128, 2 * G), idesc1 = umma_idesc_f16(128, G);
        for (int m = 0; m < L.n_mt; ++m) {
            const uint32_t acc = tmem_base + b * e_cols + (uint32_t)m * 2u * (uint32_t)G;
            for (int ks = 0; ks < ks_e; ++ks) {
                const int kc = ks >> 2, j = ks & 3;
                const uint32_t a_hi = smem_u32(sm + L.off_xa) + ((uint32_t)(kc * 2 + 0) * (uint32_t)L.n_box + 2u * (uint32_t)m) * BOX_BYTES;
                const uint32_t a_lo = smem_u32(sm + L.off_xa) + ((uint32_t)(kc * 2 + 1) * (uint32_t)L.n_box + 2u * (uint32_t)m) * BOX_BYTES;
                const uint32_t wb = smem_u32(sm + L.off_we) + b * L.we_stage + (uint32_t)kc * 2u * (uint32_t)G * 128u;
                const uint64_t db = umma_desc_sw128(wb) + (uint64_t)(kDescKStep * j);
                umma_f16(acc, umma_desc_sw128(a_hi) + (uint64_t)(kDescKStep * j), db, idesc2, ks != 0 ? 1u : 0u);
                umma_f16(acc + (uint32_t)G, umma_desc_sw128(a_lo) + (uint64_t)(kDescKStep * j), db, idesc1, 1u);
            }
        }
        umma_commit(&bar_e[b]);
    };

    for (int seg = blockIdx.x; seg < p.batch; seg += gridDim.x, ++seg_it) {
        // ================================ phase A ================================
        if (warp == 0) {
            if (elect_one()) {
                arrive_expect_tx(&bar_x, L.xa_bytes);
                for (int kc = 0; kc < L.kc_e; ++kc)
                    for (int pl = 0; pl < 2; ++pl)
                        for (int bx = 0; bx < L.n_box; ++bx)
                            tma_load_5d(sm + L.off_xa + ((uint32_t)(kc * 2 + pl) * (uint32_t)L.n_box + (uint32_t)bx) * BOX_BYTES, &p.xmap,
                                        kc * 64, seg * npix + bx * 64, 0, 0, pl, &bar_x);
                load_we(0, gi & 1u);
                if (n_grp > 1) load_we(1, (gi + 1u) & 1u);
                mbar_wait(&bar_x, seg_it & 1u);
                mbar_wait(&bar_w[gi & 1u], (gi >> 1) & 1u);
                tc_fence_after();
                issue_expand(gi);
            }
            __syncwarp();
        }
        for (int g = 0; g < n_grp; ++g, ++gi) {
            const uint32_t b = gi & 1u;
            // (0) the next group's MMAs go to the other TMEM buffer (free since the barrier after the previous group's
            //     TMEM reads) and run under this group's depthwise arithmetic
            if (warp == 0) {
                if (g + 1 < n_grp && elect_one()) {
                    mbar_wait(&bar_w[b ^ 1u], ((gi + 1u) >> 1) & 1u);
                    tc_fence_after();
                    issue_expand(gi + 1u);
                }
                __syncwarp();
            }
            // (1) E_g complete -> its weight stage is free again: fetch group g + 2 into it
            mbar_wait(&bar_e[b], (gi >> 1) & 1u);
            tc_fence_after();
            if (warp == 0) {
                if (g + 2 < n_grp && elect_one()) load_we(g + 2, b);
                __syncwarp();
            }
            // TMEM -> bias, SiLU -> patch.  warp (q, m, slice): rows m*128 + q*32 + lane, channels slice*16 .. +15
            for (int u = warp >> 2; u < L.n_mt * SL; u += MB_WARPS / 4) {
                const int m = u / SL, sl = u - m * SL;
                const int pix = m * 128 + q * 32 + lane;
                if (m * 128 + q * 32 < npix) {                       // warp-uniform: this quarter holds real rows
                    uint32_t rm[16], rc[16];
                    const uint32_t t = tmem_base + b * e_cols + (uint32_t)m * 2u * (uint32_t)G + (uint32_t)(sl * 16) + ((uint32_t)(q * 32) << 16);
                    tmem_ld16_nowait(t, rm);
                    tmem_ld16_nowait(t + (uint32_t)G, rc);
                    tmem_ld_wait();
                    if (pix < npix) {
                        const int y = pix / W, x = pix - y * W;
                        const int ppix = (y + PAD) * wp + x + PAD;
                        const float* be = p.be + g * G + sl * 16;
                        float2* dst = reinterpret_cast<float2*>(sm + L.off_patch) + (size_t)(sl * 8) * L.npixp + ppix;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 bb = __ldg(reinterpret_cast<const float2*>(be) + j);
                            const float v0 = silu_f(__uint_as_float(rm[2 * j]) + __uint_as_float(rc[2 * j]) + bb.x);
                            const float v1 = silu_f(__uint_as_float(rm[2 * j + 1]) + __uint_as_float(rc[2 * j + 1]) + bb.y);
                            dst[(size_t)j * L.npixp] = make_float2(v0, v1);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();                                         // (2) patch complete, TMEM buffer b free
            if (g > 0 && tid < G) {                                  // pooled mean of the previous group, fixed order
                const float* part = s_part + (size_t)(b ^ 1u) * NB * G;
                float s = 0.f;
                for (int i = 0; i < NB; ++i) s += part[i * G + tid];
                s_pool[(g - 1) * G + tid] = s * inv_np;
            }
            // depthwise conv + SiLU -> D planes (un-gated), pooled partial sums
            float2 pool = make_float2(0.f, 0.f);
            if (dw_active) {
                const int c = g * G + 2 * cp;
                unsigned long long wk[K * K];
#pragma unroll
                for (int i = 0; i < K * K; ++i) wk[i] = __ldg(reinterpret_cast<const unsigned long long*>(p.wd + (size_t)i * cexp + c));
                const unsigned long long bias = __ldg(reinterpret_cast<const unsigned long long*>(p.bd + c));
                unsigned long long acc[YB][XB];
#pragma unroll
                for (int y = 0; y < YB; ++y)
#pragma unroll
                    for (int x = 0; x < XB; ++x) acc[y][x] = bias;
                const unsigned long long* base = reinterpret_cast<const unsigned long long*>(sm + L.off_patch) + (size_t)cp * L.npixp + oy0 * wp + ox0;
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
                __half* dh = p.d_hi + ((size_t)seg * npix) * cexp + c;
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
            reinterpret_cast<float2*>(s_part + (size_t)b * NB * G)[blk * NP + cp] = pool;
            __syncthreads();                                         // (3) patch free, partials visible
        }
        if (tid < G) {                                               // last group's pooled mean
            const float* part = s_part + (size_t)((gi - 1u) & 1u) * NB * G;
            float s = 0.f;
            for (int i = 0; i < NB; ++i) s += part[i * G + tid];
            s_pool[(n_grp - 1) * G + tid] = s * inv_np;
        }
        fence_proxy_async_all();                                     // this thread's D stores -> visible to the TMA engine
        __syncthreads();

        // ================================ gate ================================
        {
            const int cpw = (cexp + MB_WARPS - 1) / MB_WARPS;       // FC1: warps split the channels, lanes = output j
            const int cbeg = warp * cpw, cend = min(cexp, cbeg + cpw);
            for (int j0 = 0; j0 < R; j0 += 32) {
                const int j = j0 + lane;
                if (j < R) {
                    float a = 0.f;
                    int cc = cbeg;
                    for (; cc + 8 <= cend; cc += 8) {
                        float wv[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) wv[u] = __ldg(p.w1 + (size_t)(cc + u) * p.ldw1 + j);
#pragma unroll
                        for (int u = 0; u < 8; ++u) a = fmaf(s_pool[cc + u], wv[u], a);
                    }
                    for (; cc < cend; ++cc) a = fmaf(s_pool[cc], __ldg(p.w1 + (size_t)cc * p.ldw1 + j), a);
                    s_fc[warp * R + j] = a;
                }
            }
            __syncthreads();
            for (int j = tid; j < R; j += MB_THREADS) {
                float v = p.b1[j];
#pragma unroll
                for (int wi = 0; wi < MB_WARPS; ++wi) v += s_fc[wi * R + j];
                s_r[j] = v * (1.0f / (1.0f + expf(-v)));
            }
            __syncthreads();
            for (int cc = tid; cc < cexp; cc += MB_THREADS) {       // FC2: thread per channel, rows of W2 contiguous in c
                float v = p.b2[cc];
                int j = 0;
                for (; j + 8 <= R; j += 8) {
                    float wv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) wv[u] = __ldg(p.w2 + (size_t)(j + u) * p.ldw2 + cc);
#pragma unroll
                    for (int u = 0; u < 8; ++u) v = fmaf(s_r[j + u], wv[u], v);
                }
                for (; j < R; ++j) v = fmaf(s_r[j], __ldg(p.w2 + (size_t)j * p.ldw2 + cc), v);
                s_gate[cc] = 1.0f / (1.0f + expf(-v));
            }
            __syncthreads();
        }

        // ================================ phase B: Y = (D * g) * Wp ================================
        const uint32_t idescP2 = umma_idesc_f16(128, fused_n ? 2 * cout : cout), idescP1 = umma_idesc_f16(128, cout);
        for (int kc = 0; kc < L.kc_p; ++kc, ++ck) {
            const uint32_t s = ck & 1u;
            // stage s is free when the MMAs of chunk ck - 2 have retired
            if (ck >= 2u) mbar_wait(&bar_m[s], ((ck >> 1) - 1u) & 1u);
            if (warp == 0) {
                if (elect_one()) {
                    arrive_expect_tx(&bar_d[s], L.da_stage);
                    for (int pl = 0; pl < 2; ++pl)
                        for (int bx = 0; bx < L.n_box; ++bx)
                            tma_load_5d(sm + L.off_da + s * L.da_stage + ((uint32_t)pl * (uint32_t)L.n_box + (uint32_t)bx) * BOX_BYTES, &p.dmap,
                                        kc * 64, seg * npix + bx * 64, 0, 0, pl, &bar_d[s]);
                }
                __syncwarp();
            }
            // Wp^T chunk scaled by the gate -> [W_hi | W_lo] operand image (row n = output channel, 64 K values)
            uint8_t* wst = sm + L.off_wp + s * L.wp_stage;
            for (int u = tid; u < cout * 8; u += MB_THREADS) {
                const int n = u >> 3, ku = u & 7;
                const int c0 = kc * 64 + ku * 8;
                float v[8];
                if (c0 < cexp) {                                     // cexp is a multiple of 8
                    const float4* src = reinterpret_cast<const float4*>(p.wpT + (size_t)n * cexp + c0);
                    const float4 a = __ldg(src), bq = __ldg(src + 1);
                    const float4 g0 = *reinterpret_cast<const float4*>(s_gate + c0), g1 = *reinterpret_cast<const float4*>(s_gate + c0 + 4);
                    v[0] = a.x * g0.x; v[1] = a.y * g0.y; v[2] = a.z * g0.z; v[3] = a.w * g0.w;
                    v[4] = bq.x * g1.x; v[5] = bq.y * g1.y; v[6] = bq.z * g1.z; v[7] = bq.w * g1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = 0.f;
                }
                uint4 hi, lo;
                split8(v, hi, lo);
                const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)ku);
                *reinterpret_cast<uint4*>(wst + off) = hi;
                *reinterpret_cast<uint4*>(wst + (uint32_t)cout * 128u + off) = lo;
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (warp == 0) {
                if (elect_one()) {
                    mbar_wait(&bar_d[s], (ck >> 1) & 1u);
                    tc_fence_after();
                    const int ksn = min(4, (cexp - kc * 64 + 15) >> 4);
                    for (int m = 0; m < L.n_mt; ++m) {
                        const uint32_t acc = tmem_base + (uint32_t)m * p_cols;
                        const uint32_t a_hi = smem_u32(sm + L.off_da) + s * L.da_stage + (2u * (uint32_t)m) * BOX_BYTES;
                        const uint32_t a_lo = a_hi + (uint32_t)L.n_box * BOX_BYTES;
                        const uint32_t wb = smem_u32(wst);
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
                    umma_commit(&bar_m[s]);
                }
                __syncwarp();
            }
        }
        // every chunk's MMAs complete (commits arrive in order: the last one covers all)
        {
            const uint32_t last = ck - 1u;
            mbar_wait(&bar_m[last & 1u], (last >> 1) & 1u);
            if (L.kc_p >= 2) { const uint32_t prev = ck - 2u; mbar_wait(&bar_m[prev & 1u], (prev >> 1) & 1u); }
            tc_fence_after();
        }
        // epilogue: main + correction + bias (+ residual) -> hi/lo planes.  warp = (q, m, column slices)
        {
            const int n_sl = cout >> 4;
            for (int u = warp >> 2; u < L.n_mt * n_sl; u += MB_WARPS / 4) {
                const int m = u / n_sl, sl = u - m * n_sl;
                if (m * 128 + q * 32 >= npix) continue;              // warp-uniform
                const int pix = m * 128 + q * 32 + lane;
                uint32_t rm[16], rc[16];
                const uint32_t t = tmem_base + (uint32_t)m * p_cols + (uint32_t)(sl * 16) + ((uint32_t)(q * 32) << 16);
                tmem_ld16_nowait(t, rm);
                tmem_ld16_nowait(t + (uint32_t)cout, rc);
                tmem_ld_wait();
                if (pix < npix) {
                    const int n0 = sl * 16;
                    float v[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bp + n0) + j4);
                        v[4 * j4 + 0] = __uint_as_float(rm[4 * j4 + 0]) + __uint_as_float(rc[4 * j4 + 0]) + bv.x;
                        v[4 * j4 + 1] = __uint_as_float(rm[4 * j4 + 1]) + __uint_as_float(rc[4 * j4 + 1]) + bv.y;
                        v[4 * j4 + 2] = __uint_as_float(rm[4 * j4 + 2]) + __uint_as_float(rc[4 * j4 + 2]) + bv.z;
                        v[4 * j4 + 3] = __uint_as_float(rm[4 * j4 + 3]) + __uint_as_float(rc[4 * j4 + 3]) + bv.w;
                    }
                    const size_t o = ((size_t)seg * npix + pix) * cout + n0;
                    if (p.res_hi) {
#pragma unroll
                        for (int j2 = 0; j2 < 2; ++j2) {
                            const uint4 qh = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o) + j2);
                            const uint4 ql = __ldg(reinterpret_cast<const uint4*>(p.res_hi + p.res_plane + o) + j2);
                            const __half2* h2 = reinterpret_cast<const __half2*>(&qh);
                            const __half2* l2 = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 a = __half22float2(h2[e]), d = __half22float2(l2[e]);
                                v[8 * j2 + 2 * e] += a.x + d.x;
                                v[8 * j2 + 2 * e + 1] += a.y + d.y;
                            }
                        }
                    }
                    uint4 hq[2], lq[2];
                    split8(v, hq[0], lq[0]);
                    split8(v + 8, hq[1], lq[1]);
                    uint4* oh = reinterpret_cast<uint4*>(p.out_hi + o);
                    uint4* ol = reinterpret_cast<uint4*>(p.out_hi + p.out_plane + o);
                    oh[0] = hq[0]; oh[1] = hq[1];
                    ol[0] = lq[0]; ol[1] = lq[1];
                }
            }
        }
        tc_fence_before();
        __syncthreads();               // TMEM and the shared operand region are free for the next segment
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

namespace {

constexpr size_t MB_SMEM_MAX = 224 * 1024;      // dynamic part: the 227 KB per-CTA limit includes the static barriers

struct MbVariant { int k, g, xb, yb; };

// the pixel block and channel group this build instantiates for a layer shape (stride 1 only), or g = 0
MbVariant mb_pick(int h, int w, int k) {
    if ((k != 3 && k != 5) || w % 2 || h % 3) return MbVariant{0, 0, 0, 0};
    const int nblk = (h / 3) * (w / 2);              // 2 x 3 output blocks
    if (nblk <= 16) return MbVariant{k, 64, 2, 3};   // 32 channel pairs x 16 block slots
    if (nblk <= 32) return MbVariant{k, 32, 2, 3};   // 16 channel pairs x 32 block slots
    return MbVariant{0, 0, 0, 0};
}

template <int K, int G>
cudaError_t mb_launch_kg(const MbconvParams& p, int grid, size_t smem, cudaStream_t stream) {
    k_mbconv<K, G, 2, 3><<<grid, MB_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

int mbconv_group(int h, int w, int k) { return mb_pick(h, w, k).g; }

bool mbconv_supported(int h, int w, int k, int stride, int cin, int cexp, int cout, int r) {
    if (stride != 1) return false;
    const MbVariant v = mb_pick(h, w, k);
    if (v.g == 0) return false;
    if ((cin & 7) || (cexp % v.g) || (cout & 15) || cout > 256 || r < 1 || r > 256) return false;
    const MbLayout L = mb_layout(h, w, k, cin, cexp, cout, r, v.g);
    if (L.total > MB_SMEM_MAX) return false;
    if ((uint32_t)L.n_mt * 2u * (uint32_t)v.g * 2u > 512u) return false;     // two E buffers in TMEM
    if ((uint32_t)L.n_mt * 2u * (uint32_t)cout > 512u) return false;         // projection accumulators in TMEM
    return true;
}

cudaError_t mbconv_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_mbconv<3, 32, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mbconv<5, 32, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mbconv<3, 64, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mbconv<5, 64, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM_MAX);
    return e;
}

cudaError_t launch_mbconv(const MbconvParams& p, int num_sms, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    if (!mbconv_supported(p.h, p.w, p.k, 1, p.cin, p.cexp, p.cout, p.r)) return cudaErrorInvalidValue;
    const MbVariant v = mb_pick(p.h, p.w, p.k);
    const MbLayout L = mb_layout(p.h, p.w, p.k, p.cin, p.cexp, p.cout, p.r, v.g);
    const int grid = p.batch < num_sms ? p.batch : num_sms;
    if (v.k == 3 && v.g == 32) return mb_launch_kg<3, 32>(p, grid, L.total, stream);
    if (v.k == 5 && v.g == 32) return mb_launch_kg<5, 32>(p, grid, L.total, stream);
    if (v.k == 3 && v.g == 64) return mb_launch_kg<3, 64>(p, grid, L.total, stream);
    return mb_launch_kg<5, 64>(p, grid, L.total, stream);
}

}  // namespace bn
