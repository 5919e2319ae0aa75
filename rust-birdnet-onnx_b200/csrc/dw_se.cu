// Depthwise conv + SiLU + squeeze-excite, ONE kernel per MBConv block (rows A8 of SURVEY.md section 8).
//
//   D[b][p][c]  = silu(dw_kxk(E)[b][p][c] + bias[c])
//   pooled[b][c] = mean_p D[b][p][c];  r = silu(W1^T pooled + b1);  g = sigmoid(W2^T r + b2)
//   out[b][p][c] = D[b][p][c] * g[b][c]                      (what the projection conv consumes)
//
// One CTA owns one segment.  It walks the channel groups of the expanded tensor E (zero-halo FP32
// patch of CG channels in shared memory), writes the un-gated D planes, keeps the pooled sums in
// shared memory, runs the two tiny FCs, and then rescales ITS OWN D in place.  D of a segment is
// 0.2-0.5 MB and was written microseconds earlier by the same CTA, so the rescale pass reads it
// mostly from L2: DRAM traffic of the whole dw + SE tail approaches "E read once, D*g written once".
// The previous design (dw kernel, SE kernel) moved E + 3 x D through HBM and paid two launches.
//
// Input E comes in one of two forms:
//   F32IN  (normal case: E is produced by the 1x1 expand conv for this kernel alone, so that conv
//          writes plain FP32): cp.async lands the next group's pixels straight in the interior of
//          the other patch buffer while the current group is being convolved - one barrier per group.
//   planes (generic: hi/lo fp16 planes): cp.async lands the raw planes, a conversion pass adds
//          hi + lo into the FP32 patch - two barriers per group.
//
// Thread mapping in the conv phase: a thread owns a channel PAIR (64-bit shared-memory reads, packed
// FFMA2 = fma.rn.f32x2, half2 plane stores) and an XB x YB block of output pixels, so every staged
// input value feeds several FMAs.  Sums over pixels are reduced in a fixed order -> deterministic.
#include "kernels.h"
#include "fast_act.cuh"

#include <cstdio>
#include <cstdlib>

namespace bn {

namespace {

__device__ __forceinline__ float silu_fast(float v) { return silu_approx(v); }

__device__ __forceinline__ void store_pair(__half* hi, size_t plane, size_t o, float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 bk = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - bk.x, b - bk.y);
    *reinterpret_cast<__half2*>(hi + o) = h;
    *reinterpret_cast<__half2*>(hi + plane + o) = l;
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }

constexpr int DW_THREADS = 256;
constexpr int DW_WARPS = DW_THREADS / 32;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory carve-up, in floats (host and device agree through these functions)
__host__ __device__ inline int dw_part_floats(int r) {
    const int n = DW_WARPS * r > 4 * DW_THREADS ? DW_WARPS * r : 4 * DW_THREADS;      // FC1 partials | 2 x pooled partials
    return (n + 3) & ~3;
}
__host__ __device__ inline size_t dw_fixed_floats(int c, int r) { return (size_t)2 * c + ((r + 3) & ~3) + dw_part_floats(r); }
// per channel of the group: patch buffers + staging + 2 x (weights + bias)
__host__ __device__ inline size_t dw_per_channel_floats(int hin, int win, int pad, int k, bool f32in) {
    const size_t patch = (size_t)(hin + 2 * pad) * (win + 2 * pad);
    return (f32in ? 2 * patch : patch + (size_t)hin * win) + 2 * (size_t)(k * k + 1);
}

#ifndef BN_DW_MINB
#define BN_DW_MINB 2
#endif
template <int K, int S, int XB, int YB, int CG_SHIFT, bool F32IN>
__global__ void __launch_bounds__(DW_THREADS, BN_DW_MINB) k_dw_se(const DwSeParams p) {
    constexpr int cg_shift = CG_SHIFT;
    extern __shared__ __align__(16) float smem_dw[];
    constexpr int NCOL = (XB - 1) * S + K;
    constexpr int NROW = (YB - 1) * S + K;
    constexpr int CG = 1 << cg_shift, NP = CG >> 1, PG = DW_THREADS / NP;
    constexpr int WROWS = K * K + 1;                    // weight rows + the bias row
    const int C = p.c, R = p.r;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int hp = p.hin + 2 * p.pad, wp = p.win + 2 * p.pad;
    const int npin = p.hin * p.win;
    const int npout = p.hout * p.wout;
    const int patch_floats = (hp * wp) << cg_shift;
    float* s_pool = smem_dw;                            // [C]
    float* s_gate = s_pool + C;                         // [C]
    float* s_r = s_gate + C;                            // [R rounded up to 4]
    float* s_part = s_r + ((R + 3) & ~3);               // FC1 partials [DW_WARPS][R] | pooled partials [2][PG][CG]
    float* s_in = s_part + dw_part_floats(R);           // F32IN: [2][hp][wp][CG]; planes: [hp][wp][CG] + raw [hi|lo][npin][CG] fp16
    float* s_w = s_in + (F32IN ? 2 * (size_t)patch_floats : (size_t)patch_floats + ((size_t)npin << cg_shift));   // [2][WROWS][CG]
    __half* s_raw = reinterpret_cast<__half*>(s_in + patch_floats);
    const int cp = tid & (NP - 1), pg = tid / NP;
    const int xblocks = p.wout / XB;
    const int nblk = (p.hout / YB) * xblocks;
    __half* out_hi = p.out.hi + (size_t)b * npout * C;
    const float inv_np = 1.0f / (float)npout;
    const uint32_t in_u32 = (uint32_t)__cvta_generic_to_shared(s_in);
    const uint32_t raw_u32 = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t w_u32 = (uint32_t)__cvta_generic_to_shared(s_w);

    // next channel group -> shared memory, asynchronously: weights + bias, and the input pixels
    auto prefetch = [&](int c0, int buf) {
        constexpr int WCH = WROWS * (CG / 4);
        for (int i = tid; i < WCH; i += DW_THREADS) {
            const int row = i / (CG / 4), ch = i - row * (CG / 4);
            const float* src = (row < K * K ? p.weight + (size_t)row * C : p.bias) + c0 + ch * 4;
            cp_async16(w_u32 + (uint32_t)((buf * WROWS + row) * CG + ch * 4) * 4u, src);
        }
        if (F32IN) {
            constexpr int UPP = CG / 4;                  // 16-byte units per pixel
            const float* src0 = p.in_f32 + (size_t)b * npin * C + c0;
            const uint32_t d0 = in_u32 + (uint32_t)buf * (uint32_t)patch_floats * 4u;
            for (int u = tid; u < npin * UPP; u += DW_THREADS) {
                const int pix = u / UPP, cu = u - pix * UPP;
                const int y = pix / p.win, x = pix - y * p.win;
                cp_async16(d0 + (uint32_t)((((y + p.pad) * wp + x + p.pad) << cg_shift) + cu * 4) * 4u, src0 + (size_t)pix * C + cu * 4);
            }
        } else {
            constexpr int UPP = CG / 8;                  // 16-byte units per pixel and plane
            const __half* src0 = p.in.hi + (size_t)b * npin * C + c0;
            const uint32_t plane_bytes = (uint32_t)npin << (cg_shift + 1);
            for (int u = tid; u < npin * UPP; u += DW_THREADS) {
                const int pix = u / UPP, cu = u - pix * UPP;
                const size_t o = (size_t)pix * C + cu * 8;
                const uint32_t d = raw_u32 + (uint32_t)((pix << cg_shift) + cu * 8) * 2u;
                cp_async16(d, src0 + o);
                cp_async16(d + plane_bytes, src0 + p.in.plane + o);
            }
        }
    };
    prefetch(0, 0);
    // the halo never changes: zero it once (the interior is overwritten by every group).  F32IN: the cp.async of
    // group 0 targets interior cells, and a zero store to an interior cell could race with it - only halo cells
    // are zeroed there.
    if (F32IN) {
        const int per_buf = hp * wp * (CG / 4);
        for (int i = tid; i < 2 * per_buf; i += DW_THREADS) {
            const int bufi = i / per_buf, rem = i - bufi * per_buf;
            const int pix = rem / (CG / 4);
            const int y = pix / wp, x = pix - y * wp;
            if (y < p.pad || y >= p.hin + p.pad || x < p.pad || x >= p.win + p.pad)
                reinterpret_cast<float4*>(s_in)[(size_t)bufi * (patch_floats >> 2) + rem] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int i = tid; i < patch_floats >> 2; i += DW_THREADS) reinterpret_cast<float4*>(s_in)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    int g = 0;
    for (int c0 = 0; c0 < C; c0 += CG, ++g) {
        const int buf = g & 1;
        cp_async_wait_all();
        __syncthreads();                                // group landed; every thread is past the previous group's conv
        if (g > 0 && tid < CG) {                        // pooled mean of the previous group, fixed summation order
            const float* part = s_part + (buf ^ 1) * (2 * DW_THREADS);
            float s = 0.f;
            for (int q = 0; q < PG; ++q) s += part[q * CG + tid];
            s_pool[c0 - CG + tid] = s * inv_np;
        }
        const float* patch = s_in;
        if (F32IN) {
            patch = s_in + (size_t)buf * patch_floats;
            if (c0 + CG < C) prefetch(c0 + CG, buf ^ 1);
        } else {
            // ---- raw fp16 hi + lo -> FP32 interior of the patch ----
            constexpr int UPP = CG / 8;
            for (int u = tid; u < npin * UPP; u += DW_THREADS) {
                const int pix = u / UPP, cu = u - pix * UPP;
                const int y = pix / p.win, x = pix - y * p.win;
                const __half* r = s_raw + ((size_t)pix << cg_shift) + cu * 8;
                const uint4 qh = *reinterpret_cast<const uint4*>(r);
                const uint4 ql = *reinterpret_cast<const uint4*>(r + ((size_t)npin << cg_shift));
                const __half2* h = reinterpret_cast<const __half2*>(&qh);
                const __half2* l = reinterpret_cast<const __half2*>(&ql);
                const float2 a0 = __half22float2(h[0]), a1 = __half22float2(h[1]), a2 = __half22float2(h[2]), a3 = __half22float2(h[3]);
                const float2 d0 = __half22float2(l[0]), d1 = __half22float2(l[1]), d2 = __half22float2(l[2]), d3 = __half22float2(l[3]);
                float4* dst = reinterpret_cast<float4*>(s_in + ((size_t)((y + p.pad) * wp + x + p.pad) << cg_shift) + cu * 8);
                dst[0] = make_float4(a0.x + d0.x, a0.y + d0.y, a1.x + d1.x, a1.y + d1.y);
                dst[1] = make_float4(a2.x + d2.x, a2.y + d2.y, a3.x + d3.x, a3.y + d3.y);
            }
        }
        // this thread's channel pair: weights and bias (landed with the pixels)
        const int c = c0 + 2 * cp;
        unsigned long long w[K * K];
        const float* wb = s_w + (size_t)buf * WROWS * CG + 2 * cp;
#pragma unroll
        for (int i = 0; i < K * K; ++i) w[i] = *reinterpret_cast<const unsigned long long*>(wb + i * CG);
        const unsigned long long bias = *reinterpret_cast<const unsigned long long*>(wb + K * K * CG);
        if (!F32IN) {
            __syncthreads();                            // patch ready, raw buffer free
            if (c0 + CG < C) prefetch(c0 + CG, buf ^ 1);
        }
        // ---- depthwise conv + SiLU, un-gated D to the planes, pooled partial sums ----
        float2 pool = make_float2(0.f, 0.f);
        for (int blk = pg; blk < nblk; blk += PG) {
            const int by = blk / xblocks;
            const int oy0 = by * YB, ox0 = (blk - by * xblocks) * XB;
            unsigned long long acc[YB][XB];
#pragma unroll
            for (int y = 0; y < YB; ++y)
#pragma unroll
                for (int j = 0; j < XB; ++j) acc[y][j] = bias;
            const float* base = patch + (((size_t)(oy0 * S) * wp + ox0 * S) << cg_shift) + 2 * cp;
            // every staged value is loaded once and scattered into all the outputs of the block it touches:
            // (NROW x NCOL) 64-bit shared-memory reads for XB*YB*K*K packed FMAs - shared-memory bandwidth is the most
            // utilised resource of this loop, so taller / wider blocks are what make it faster
#pragma unroll
            for (int r = 0; r < NROW; ++r) {
                const float* rowp = base + (size_t)(r * wp) * CG;
#pragma unroll
                for (int x = 0; x < NCOL; ++x) {
                    const unsigned long long v = *reinterpret_cast<const unsigned long long*>(rowp + x * CG);
#pragma unroll
                    for (int y = 0; y < YB; ++y) {
                        const int ky = r - y * S;
                        if (ky < 0 || ky >= K) continue;
#pragma unroll
                        for (int j = 0; j < XB; ++j) {
                            const int kx = x - j * S;
                            if (kx < 0 || kx >= K) continue;
                            acc[y][j] = ffma2(v, w[ky * K + kx], acc[y][j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int y = 0; y < YB; ++y)
#pragma unroll
                for (int j = 0; j < XB; ++j) {
                    const float v0 = silu_fast(lo_f(acc[y][j])), v1 = silu_fast(hi_f(acc[y][j]));
                    pool.x += v0;
                    pool.y += v1;
                    if (!(p.debug & 4)) store_pair(out_hi, p.out.plane, ((size_t)(oy0 + y) * p.wout + ox0 + j) * C + c, v0, v1);
                }
        }
        // pooled partials [pg][CG], double buffered by group parity (read after the next group's barrier)
        reinterpret_cast<float2*>(s_part + buf * (2 * DW_THREADS))[tid] = pool;
    }
    __syncthreads();
    if (tid < CG) {
        const float* part = s_part + ((g - 1) & 1) * (2 * DW_THREADS);
        float s = 0.f;
        for (int q = 0; q < PG; ++q) s += part[q * CG + tid];
        s_pool[C - CG + tid] = s * inv_np;
    }
    __syncthreads();
    if (p.pooled_out)
        for (int i = tid; i < C; i += DW_THREADS) p.pooled_out[(size_t)b * C + i] = s_pool[i];

    // ---- gate: r = silu(W1^T pooled + b1), g = sigmoid(W2^T r + b2) ----
    if (!(p.debug & 2)) {
        const int cpw = (C + DW_WARPS - 1) / DW_WARPS;          // FC1: warps split C, lanes = output j
        const int cbeg = warp * cpw, cend = min(C, cbeg + cpw);
        for (int j0 = 0; j0 < R; j0 += 32) {
            const int j = j0 + lane;
            if (j < R) {
                float acc = 0.f;
                int cc = cbeg;
                for (; cc + 8 <= cend; cc += 8) {
                    float wv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) wv[u] = __ldg(p.w1 + (size_t)(cc + u) * p.ldw1 + j);
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = fmaf(s_pool[cc + u], wv[u], acc);
                }
                for (; cc < cend; ++cc) acc = fmaf(s_pool[cc], __ldg(p.w1 + (size_t)cc * p.ldw1 + j), acc);
                s_part[warp * R + j] = acc;
            }
        }
        __syncthreads();
        for (int j = tid; j < R; j += DW_THREADS) {
            float v = p.b1[j];
#pragma unroll
            for (int wi = 0; wi < DW_WARPS; ++wi) v += s_part[wi * R + j];
            s_r[j] = v * (1.0f / (1.0f + expf(-v)));
        }
        __syncthreads();
        for (int cc = tid; cc < C; cc += DW_THREADS) {          // FC2: thread per channel, rows of W2 contiguous in c
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
            const float gt = 1.0f / (1.0f + expf(-v));
            s_gate[cc] = gt;
            if (p.gate_out) p.gate_out[(size_t)b * C + cc] = gt;
        }
        __syncthreads();          // also orders this CTA's D stores before the loads below
    }

    // ---- in-place rescale of this segment's D (mostly L2-resident): 8 channels per unit, 4 units in flight per thread ----
    if (!(p.debug & 1)) {
        const int cunits = C >> 3;
        const int total = npout * cunits;
        constexpr int UB = 4;
        for (int u0 = tid; u0 < total; u0 += DW_THREADS * UB) {
            uint4 qh[UB], ql[UB];
#pragma unroll
            for (int i = 0; i < UB; ++i) {
                const int u = u0 + i * DW_THREADS;
                if (u < total) {
                    const __half* ph = out_hi + (size_t)u * 8;
                    qh[i] = *reinterpret_cast<const uint4*>(ph);
                    ql[i] = *reinterpret_cast<const uint4*>(ph + p.out.plane);
                }
            }
#pragma unroll
            for (int i = 0; i < UB; ++i) {
                const int u = u0 + i * DW_THREADS;
                if (u >= total) break;
                const int cu = (u % cunits) << 3;
                const float4 g0 = *reinterpret_cast<const float4*>(s_gate + cu);
                const float4 g1 = *reinterpret_cast<const float4*>(s_gate + cu + 4);
                const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const __half2* h = reinterpret_cast<const __half2*>(&qh[i]);
                const __half2* l = reinterpret_cast<const __half2*>(&ql[i]);
                __half2 oh[4], ol[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = __half22float2(h[e]), d = __half22float2(l[e]);
                    const float v0 = (a.x + d.x) * gv[2 * e], v1 = (a.y + d.y) * gv[2 * e + 1];
                    oh[e] = __floats2half2_rn(v0, v1);
                    const float2 bk = __half22float2(oh[e]);
                    ol[e] = __floats2half2_rn(v0 - bk.x, v1 - bk.y);
                }
                __half* ph = out_hi + (size_t)u * 8;
                *reinterpret_cast<uint4*>(ph) = *reinterpret_cast<uint4*>(oh);
                *reinterpret_cast<uint4*>(ph + p.out.plane) = *reinterpret_cast<uint4*>(ol);
            }
        }
    }
}

constexpr size_t DW_SE_SMEM_TWO = (BN_DW_MINB >= 3 ? 74 : 113) * 1024;      // BN_DW_MINB CTAs per SM
constexpr size_t DW_SE_SMEM_MAX = 220 * 1024;

template <int K, int S, int XB, int YB, bool F32IN>
cudaError_t launch_one(const DwSeParams& p, int cg_shift, size_t smem, cudaStream_t stream) {
    if (cg_shift == 6) k_dw_se<K, S, XB, YB, 6, F32IN><<<p.batch, DW_THREADS, smem, stream>>>(p);
    else if (cg_shift == 5) k_dw_se<K, S, XB, YB, 5, F32IN><<<p.batch, DW_THREADS, smem, stream>>>(p);
    else if (cg_shift == 4) k_dw_se<K, S, XB, YB, 4, F32IN><<<p.batch, DW_THREADS, smem, stream>>>(p);
    else k_dw_se<K, S, XB, YB, 3, F32IN><<<p.batch, DW_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int K, int S, int XB, int YB, bool F32IN>
cudaError_t set_attr_one() {
    cudaError_t e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 6, F32IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 5, F32IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 4, F32IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 3, F32IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    return e;
}

// pixel blocks (XB x YB outputs per thread): stride 1 -> 4x3, 2x3, 4x1, 1x3; stride 2 -> 2x3, 2x1, 1x3.  The 4x3 / 2x3
// blocks exist only for the FP32 hand-off input.
template <int K, int S>
cudaError_t set_attr_ks() {
    cudaError_t e = set_attr_one<K, S, 1, 3, true>();
    if (e == cudaSuccess) e = set_attr_one<K, S, 1, 3, false>();
    if (e == cudaSuccess) e = set_attr_one<K, S, S == 1 ? 4 : 2, 1, true>();
    if (e == cudaSuccess) e = set_attr_one<K, S, S == 1 ? 4 : 2, 1, false>();
    if (e == cudaSuccess) e = set_attr_one<K, S, 2, 3, true>();
    if (e == cudaSuccess && S == 1) e = set_attr_one<K, S, 4, 3, true>();
    return e;
}

template <int K, int S>
cudaError_t launch_ks(const DwSeParams& p, int xb, int yb, int cg_shift, size_t smem, cudaStream_t stream) {
    const bool f32 = p.in_f32 != nullptr;
    if (f32 && yb == 3 && xb == 4 && S == 1) return launch_one<K, S, 4, 3, true>(p, cg_shift, smem, stream);
    if (f32 && yb == 3 && xb == 2) return launch_one<K, S, 2, 3, true>(p, cg_shift, smem, stream);
    if (yb == 3 && xb == 1) return f32 ? launch_one<K, S, 1, 3, true>(p, cg_shift, smem, stream) : launch_one<K, S, 1, 3, false>(p, cg_shift, smem, stream);
    if (yb == 1 && xb == (S == 1 ? 4 : 2))
        return f32 ? launch_one<K, S, S == 1 ? 4 : 2, 1, true>(p, cg_shift, smem, stream) : launch_one<K, S, S == 1 ? 4 : 2, 1, false>(p, cg_shift, smem, stream);
    return cudaErrorInvalidValue;
}

}  // namespace

// Picks the pixel block (XB x YB) and the channel group; false when the layer does not fit this kernel.
static bool dw_se_config(const DwSeParams& p, int* xb, int* yb, int* cg_shift, size_t* smem) {
    if ((p.k != 3 && p.k != 5) || (p.stride != 1 && p.stride != 2) || p.pad != p.k / 2 || (p.c & 15) || p.act != KACT_SILU) return false;
    if (p.r < 1 || p.r > 256 || p.batch <= 0) return false;
    const bool f32 = p.in_f32 != nullptr;
    const size_t per_ch = dw_per_channel_floats(p.hin, p.win, p.pad, p.k, f32) * sizeof(float);
    const size_t fixed = dw_fixed_floats(p.c, p.r) * sizeof(float);
    static const int force = [] { const char* ev = getenv("BN_DW_SHAPE"); return ev ? atoi(ev) : 0; }();     // e.g. 41, 43, 23, 13, 21
    const int shapes[5][2] = {{4, 3}, {2, 3}, {4, 1}, {2, 1}, {1, 3}};
    double best_cost = 1e30;
    int bx = 0, by = 0, bsh = -1;
    for (auto& sh2 : shapes) {
        const int x = sh2[0], y = sh2[1];
        if (force && force != x * 10 + y) continue;
        if ((x == 4 && p.stride != 1) || (x == 2 && y == 1 && p.stride != 2)) continue;       // instantiated combinations only
        if ((x * y > 4 || (x == 2 && y == 3)) && !f32) continue;
        if (p.wout % x || p.hout % y) continue;
        const int nblk = (p.hout / y) * (p.wout / x);
        // cost per output: 64-bit shared-memory reads (2 LSU cycles per warp each, one LSU per SM) + packed FMAs (2 cycles on one of 4 pipes)
        const int nrow = (y - 1) * p.stride + p.k, ncol = (x - 1) * p.stride + p.k;
        const double per_out = 2.0 * nrow * ncol / (x * y) + 0.5 * p.k * p.k + 6.0;
        for (int sh = 6; sh >= 3; --sh) {                   // CG = 64, 32, 16, 8
            const int cg = 1 << sh;
            if (p.c % cg) continue;
            const size_t bytes = fixed + per_ch * cg;
            if (bytes > DW_SE_SMEM_MAX) continue;
            const int pgn = DW_THREADS / (cg / 2);
            const double util = (double)nblk / (double)(((nblk + pgn - 1) / pgn) * pgn);
            const double cost = per_out / util * (bytes <= DW_SE_SMEM_TWO ? 1.0 : 1.6) * (1.0 + 0.02 * (6 - sh));   // fewer, larger groups on ties
            if (cost < best_cost - 1e-9) { best_cost = cost; bx = x; by = y; bsh = sh; }
        }
    }
    if (bsh < 0) return false;
    *xb = bx; *yb = by; *cg_shift = bsh;
    *smem = fixed + per_ch * ((size_t)1 << bsh);
    return true;
}

cudaError_t dw_se_init_device() {
    cudaError_t e = set_attr_ks<3, 1>();
    if (e == cudaSuccess) e = set_attr_ks<3, 2>();
    if (e == cudaSuccess) e = set_attr_ks<5, 1>();
    if (e == cudaSuccess) e = set_attr_ks<5, 2>();
    return e;
}

bool dw_se_supported(const DwSeParams& p) {
    int xb, yb, sh;
    size_t smem;
    return dw_se_config(p, &xb, &yb, &sh, &smem);
}

cudaError_t launch_dw_se(const DwSeParams& pin, cudaStream_t stream) {
    DwSeParams p = pin;
    { static int dbg = -1; if (dbg < 0) { const char* ev = getenv("BN_DW_DEBUG"); dbg = ev ? atoi(ev) : 0; } p.debug = dbg; }
    int xb, yb, sh;
    size_t smem;
    if (!dw_se_config(p, &xb, &yb, &sh, &smem)) return cudaErrorInvalidValue;
    if (p.k == 3 && p.stride == 1) return launch_ks<3, 1>(p, xb, yb, sh, smem, stream);
    if (p.k == 3) return launch_ks<3, 2>(p, xb, yb, sh, smem, stream);
    if (p.stride == 1) return launch_ks<5, 1>(p, xb, yb, sh, smem, stream);
    return launch_ks<5, 2>(p, xb, yb, sh, smem, stream);
}

}  // namespace bn
