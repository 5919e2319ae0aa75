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
// from L2, not HBM: DRAM traffic of the whole dw + SE tail is "E read once, D*g written once".
// The previous design (dw kernel, SE kernel) moved E + 3 x D through HBM and paid two launches.
//
// Thread mapping in the conv phase: a thread owns a channel PAIR (float2 shared-memory reads,
// half2 plane stores) and an XB x YB block of output pixels, so every staged input value feeds
// several FMAs.  Sums over pixels are reduced in a fixed order -> deterministic results.
#include "kernels.h"

#include <cstdio>

namespace bn {

namespace {

__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ void store_pair(__half* hi, size_t plane, size_t o, float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 bk = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - bk.x, b - bk.y);
    *reinterpret_cast<__half2*>(hi + o) = h;
    *reinterpret_cast<__half2*>(hi + plane + o) = l;
}

constexpr int DW_THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int K, int S, int XB, int YB, int CG_SHIFT>
__global__ void __launch_bounds__(DW_THREADS, 2) k_dw_se(const DwSeParams p) {
    constexpr int cg_shift = CG_SHIFT;
    extern __shared__ __align__(16) float smem_dw[];
    constexpr int NCOL = (XB - 1) * S + K;
    constexpr int NROW = (YB - 1) * S + K;
    const int C = p.c, R = p.r;
    constexpr int CG = 1 << cg_shift, NP = CG >> 1, PG = DW_THREADS / NP;
    float* s_pool = smem_dw;                            // [C]
    float* s_gate = s_pool + C;                         // [C]
    float* s_r = s_gate + C;                            // [R rounded up to 4]
    float* s_part = s_r + ((R + 3) & ~3);               // [max(8 * R, 2 * DW_THREADS)]
    const int part_n = 8 * R > 2 * DW_THREADS ? 8 * R : 2 * DW_THREADS;
    float* s_in = s_part + ((part_n + 3) & ~3);         // [hp][wp][CG] FP32, zero halo
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int hp = p.hin + 2 * p.pad, wp = p.win + 2 * p.pad;
    const int npin = p.hin * p.win;
    const int npout = p.hout * p.wout;
    __half* s_raw = reinterpret_cast<__half*>(s_in + ((size_t)(hp * wp) << cg_shift));   // [hi|lo][hin*win][CG] fp16, next group in flight
    const int cp = tid & (NP - 1), pg = tid >> (cg_shift - 1);
    const int xblocks = p.wout / XB;
    const int nblk = (p.hout / YB) * xblocks;
    constexpr int upp = CG >> 3;                        // 16-byte units per pixel and plane
    const int nunits = npin * upp;
    const __half* in_hi = p.in.hi + (size_t)b * npin * C;
    __half* out_hi = p.out.hi + (size_t)b * npout * C;
    const float inv_np = 1.0f / (float)npout;
    const uint32_t raw_u32 = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t raw_plane_bytes = (uint32_t)npin << (cg_shift + 1);

    // raw hi/lo planes of one channel group -> shared memory, asynchronously (lands while the previous group computes)
    auto prefetch = [&](int c0) {
        for (int u = tid; u < nunits; u += DW_THREADS) {
            const int pix = u >> (cg_shift - 3), cu = u & (upp - 1);
            const size_t o = (size_t)pix * C + c0 + cu * 8;
            const uint32_t d = raw_u32 + (((uint32_t)pix << cg_shift) + (uint32_t)cu * 8u) * 2u;
            cp_async16(d, in_hi + o);
            cp_async16(d + raw_plane_bytes, in_hi + p.in.plane + o);
        }
    };
    prefetch(0);
    // the halo never changes: zero the whole FP32 patch once
    for (int i = tid; i < ((hp * wp) << cg_shift) >> 2; i += DW_THREADS) reinterpret_cast<float4*>(s_in)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int c0 = 0; c0 < C; c0 += CG) {
        cp_async_wait_all();
        __syncthreads();                                // raw group landed; every thread is past the previous group's conv
        if (c0 > 0 && tid < CG) {                       // pooled mean of the previous group, fixed summation order
            float s = 0.f;
            for (int g = 0; g < PG; ++g) s += s_part[g * CG + tid];
            s_pool[c0 - CG + tid] = s * inv_np;
        }
        // ---- raw fp16 hi + lo -> FP32 interior of the patch ----
        for (int u = tid; u < nunits; u += DW_THREADS) {
            const int pix = u >> (cg_shift - 3), cu = u & (upp - 1);
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
        // this thread's channel pair: weights and bias
        const int c = c0 + 2 * cp;
        float2 w[K * K];
#pragma unroll
        for (int i = 0; i < K * K; ++i) w[i] = __ldg(reinterpret_cast<const float2*>(p.weight + (size_t)i * C + c));
        const float2 bias = __ldg(reinterpret_cast<const float2*>(p.bias + c));
        __syncthreads();                                // patch ready, raw buffer free
        if (c0 + CG < C) prefetch(c0 + CG);
        // ---- depthwise conv + SiLU, un-gated D to the planes, pooled partial sums ----
        float2 pool = make_float2(0.f, 0.f);
        for (int blk = pg; blk < nblk; blk += PG) {
            const int by = blk / xblocks;
            const int oy0 = by * YB, ox0 = (blk - by * xblocks) * XB;
            float2 acc[YB][XB];
#pragma unroll
            for (int y = 0; y < YB; ++y)
#pragma unroll
                for (int j = 0; j < XB; ++j) acc[y][j] = bias;
            const float* base = s_in + (((size_t)(oy0 * S) * wp + ox0 * S) << cg_shift) + 2 * cp;
#pragma unroll
            for (int r = 0; r < NROW; ++r) {
                float2 col[NCOL];
                const float* rowp = base + (size_t)(r * wp) * CG;
#pragma unroll
                for (int x = 0; x < NCOL; ++x) col[x] = *reinterpret_cast<const float2*>(rowp + x * CG);
#pragma unroll
                for (int y = 0; y < YB; ++y) {
                    const int ky = r - y * S;
                    if (ky < 0 || ky >= K) continue;
#pragma unroll
                    for (int j = 0; j < XB; ++j)
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            acc[y][j].x = fmaf(col[j * S + kx].x, w[ky * K + kx].x, acc[y][j].x);
                            acc[y][j].y = fmaf(col[j * S + kx].y, w[ky * K + kx].y, acc[y][j].y);
                        }
                }
            }
#pragma unroll
            for (int y = 0; y < YB; ++y)
#pragma unroll
                for (int j = 0; j < XB; ++j) {
                    const float v0 = silu_fast(acc[y][j].x), v1 = silu_fast(acc[y][j].y);
                    pool.x += v0;
                    pool.y += v1;
                    store_pair(out_hi, p.out.plane, ((size_t)(oy0 + y) * p.wout + ox0 + j) * C + c, v0, v1);
                }
        }
        // s_part [pg][CG]: last read (by tid < CG) was before this iteration's second barrier
        reinterpret_cast<float2*>(s_part)[tid] = pool;
    }
    __syncthreads();
    if (tid < CG) {
        float s = 0.f;
        for (int g = 0; g < PG; ++g) s += s_part[g * CG + tid];
        s_pool[C - CG + tid] = s * inv_np;
    }
    __syncthreads();
    if (p.pooled_out)
        for (int i = tid; i < C; i += DW_THREADS) p.pooled_out[(size_t)b * C + i] = s_pool[i];

    // ---- gate: r = silu(W1^T pooled + b1), g = sigmoid(W2^T r + b2) ----
    {
        const int cpw = (C + 7) / 8;                            // FC1: warps split C, lanes = output j
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
            for (int wi = 0; wi < 8; ++wi) v += s_part[wi * R + j];
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
            const float g = 1.0f / (1.0f + expf(-v));
            s_gate[cc] = g;
            if (p.gate_out) p.gate_out[(size_t)b * C + cc] = g;
        }
        __syncthreads();          // also orders this CTA's D stores before the loads below
    }

    // ---- in-place rescale of this segment's D (L2-resident): 8 channels per unit, 4 units in flight per thread ----
    {
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
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const __half2* h = reinterpret_cast<const __half2*>(&qh[i]);
                const __half2* l = reinterpret_cast<const __half2*>(&ql[i]);
                __half2 oh[4], ol[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = __half22float2(h[e]), d = __half22float2(l[e]);
                    const float v0 = (a.x + d.x) * g[2 * e], v1 = (a.y + d.y) * g[2 * e + 1];
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

size_t dw_se_fixed_floats(int c, int r) {
    const int part_n = 8 * r > 2 * DW_THREADS ? 8 * r : 2 * DW_THREADS;
    return (size_t)2 * c + ((r + 3) & ~3) + ((part_n + 3) & ~3);
}

constexpr size_t DW_SE_SMEM_TWO = 113 * 1024;      // two CTAs per SM
constexpr size_t DW_SE_SMEM_MAX = 220 * 1024;

template <int K, int S, int XB, int YB>
cudaError_t launch_one(const DwSeParams& p, int cg_shift, size_t smem, cudaStream_t stream) {
    if (cg_shift == 6) k_dw_se<K, S, XB, YB, 6><<<p.batch, DW_THREADS, smem, stream>>>(p);
    else if (cg_shift == 5) k_dw_se<K, S, XB, YB, 5><<<p.batch, DW_THREADS, smem, stream>>>(p);
    else k_dw_se<K, S, XB, YB, 4><<<p.batch, DW_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int K, int S, int XB, int YB>
cudaError_t set_attr_one() {
    cudaError_t e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dw_se<K, S, XB, YB, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SE_SMEM_MAX);
    return e;
}

// pixel blocks: 3-row maps take a whole column per thread (1 x 3); wider maps 4 (stride 1) or 2 (stride 2) pixels of a row
template <int K, int S>
cudaError_t set_attr_ks() {
    cudaError_t e = set_attr_one<K, S, 1, 3>();
    if (e == cudaSuccess) e = set_attr_one<K, S, S == 1 ? 4 : 2, 1>();
    return e;
}

template <int K, int S>
cudaError_t launch_ks(const DwSeParams& p, int xb, int yb, int cg_shift, size_t smem, cudaStream_t stream) {
    if (yb == 3) return launch_one<K, S, 1, 3>(p, cg_shift, smem, stream);
    if (xb == (S == 1 ? 4 : 2)) return launch_one<K, S, S == 1 ? 4 : 2, 1>(p, cg_shift, smem, stream);
    return cudaErrorInvalidValue;
}

}  // namespace

// Picks the pixel block (XB x YB) and the channel group; false when the layer does not fit this kernel.
static bool dw_se_config(const DwSeParams& p, int* xb, int* yb, int* cg_shift, size_t* smem) {
    if ((p.k != 3 && p.k != 5) || (p.stride != 1 && p.stride != 2) || p.pad != p.k / 2 || (p.c & 15) || p.act != KACT_SILU) return false;
    if (p.r < 1 || p.r > 256 || p.batch <= 0) return false;
    if (p.hout == 3) { *xb = 1; *yb = 3; }
    else if (p.stride == 1 && (p.wout % 4) == 0) { *xb = 4; *yb = 1; }
    else if (p.stride == 2 && (p.wout % 2) == 0) { *xb = 2; *yb = 1; }
    else return false;
    const int nblk = (p.hout / *yb) * (p.wout / *xb);
    // per channel: FP32 zero-halo patch + raw fp16 hi/lo interior of the next group
    const size_t np = (size_t)(p.hin + 2 * p.pad) * (p.win + 2 * p.pad) + (size_t)p.hin * p.win;
    const size_t fixed = dw_se_fixed_floats(p.c, p.r) * sizeof(float);
    int best = -1;
    double best_util = 0.0;
    for (int sh = 6; sh >= 4; --sh) {                       // CG = 64, 32, 16
        const int cg = 1 << sh;
        if (p.c % cg) continue;
        const size_t bytes = fixed + np * cg * sizeof(float);
        if (bytes > DW_SE_SMEM_TWO && !(sh == 4 && bytes <= DW_SE_SMEM_MAX)) continue;
        const int pgn = DW_THREADS / (cg / 2);
        const double util = (double)nblk / (double)(((nblk + pgn - 1) / pgn) * pgn);
        if (util > best_util + 1e-9) { best_util = util; best = sh; }
    }
    if (best < 0) return false;
    *cg_shift = best;
    *smem = fixed + np * ((size_t)1 << best) * sizeof(float);
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

cudaError_t launch_dw_se(const DwSeParams& p, cudaStream_t stream) {
    int xb, yb, sh;
    size_t smem;
    if (!dw_se_config(p, &xb, &yb, &sh, &smem)) return cudaErrorInvalidValue;
    if (p.k == 3 && p.stride == 1) return launch_ks<3, 1>(p, xb, yb, sh, smem, stream);
    if (p.k == 3) return launch_ks<3, 2>(p, xb, yb, sh, smem, stream);
    if (p.stride == 1) return launch_ks<5, 1>(p, xb, yb, sh, smem, stream);
    return launch_ks<5, 2>(p, xb, yb, sh, smem, stream);
}

}  // namespace bn
