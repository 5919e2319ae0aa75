// sm_100a kernels, generation 1: correct FP32 CUDA-core path for every stage of the hot path
// (front-end, CNN, epilogue).  Tensor-core (tcgen05) replacements live in tc_*.cu and are
// validated against these.
#include "kernels.h"

#include <cfloat>
#include <cmath>

namespace bn {

// ======================================================================================
// Front-end: per-segment min/max normalisation (row A7)
// ======================================================================================
__device__ __forceinline__ float nan_min(float a, float b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ float nan_max(float a, float b) { return (a > b || a != a) ? a : b; }

__global__ void __launch_bounds__(1024) k_minmax_normalize(const float* __restrict__ x, float* __restrict__ y,
                                                           int S, float eps, float half, float two) {
    const float* xs = x + (size_t)blockIdx.x * S;
    float* ys = y + (size_t)blockIdx.x * S;
    const int n4 = S >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 v = x4[i];
        mn = nan_min(nan_min(mn, v.x), nan_min(nan_min(v.y, v.z), v.w));
        mx = nan_max(nan_max(mx, v.x), nan_max(nan_max(v.y, v.z), v.w));
    }
    for (int i = (n4 << 2) + threadIdx.x; i < S; i += blockDim.x) {
        mn = nan_min(mn, xs[i]);
        mx = nan_max(mx, xs[i]);
    }
    __shared__ float s_mn[32], s_mx[32];
    for (int o = 16; o > 0; o >>= 1) {
        mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = blockDim.x >> 5;
        mn = threadIdx.x < nw ? s_mn[threadIdx.x] : INFINITY;
        mx = threadIdx.x < nw ? s_mx[threadIdx.x] : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) {
            mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (threadIdx.x == 0) { s_mn[0] = mn; s_mx[0] = mx; }
    }
    __syncthreads();
    const float lo = s_mn[0];
    // max(x - lo) == fl(max(x) - lo): subtraction of a constant is monotone
    const float den = __fadd_rn(__fsub_rn(s_mx[0], lo), eps);
    float4* y4 = reinterpret_cast<float4*>(ys);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 v = x4[i], r;
        r.x = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.x, lo), den), half), two);
        r.y = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.y, lo), den), half), two);
        r.z = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.z, lo), den), half), two);
        r.w = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.w, lo), den), half), two);
        y4[i] = r;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < S; i += blockDim.x)
        ys[i] = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(xs[i], lo), den), half), two);
}

cudaError_t launch_minmax_normalize(const float* x, float* y, int batch, int sample_count,
                                    float eps, float half, float two, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    k_minmax_normalize<<<batch, 1024, 0, stream>>>(x, y, sample_count, eps, half, two);
    return cudaGetLastError();
}

// ======================================================================================
// Generic FP32 implicit GEMM: C[M][N] = A[M][K] * W[K][ldw], A supplied by a loader functor
// ======================================================================================
constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == KACT_SILU) return v * (1.0f / (1.0f + expf(-v)));
    if (act == KACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    return v;
}

struct ConvLoader {
    const float* in;
    const float* in_scale;
    int hin, win, cin, hout, wout, k, stride, pad;
    struct Row {
        const float* base;
        const float* scale;
        int iy0, ix0;
        bool valid;
    };
    __device__ Row row(int m, int M) const {
        Row r;
        r.valid = m < M;
        int mm = r.valid ? m : 0;
        int hw = hout * wout;
        int b = mm / hw, rem = mm - b * hw;
        int oy = rem / wout, ox = rem - oy * wout;
        r.iy0 = oy * stride - pad;
        r.ix0 = ox * stride - pad;
        r.base = in + (size_t)b * hin * win * cin;
        r.scale = in_scale ? in_scale + (size_t)b * cin : nullptr;
        return r;
    }
    __device__ void load4(const Row& r, int kidx, int K, float v[4]) const {
        v[0] = v[1] = v[2] = v[3] = 0.f;
        if (!r.valid) return;
        if ((cin & 3) == 0) {
            if (kidx >= K) return;
            int tap = kidx / cin, ci = kidx - tap * cin;
            int ky = tap / k, kx = tap - ky * k;
            int iy = r.iy0 + ky, ix = r.ix0 + kx;
            if (iy < 0 || iy >= hin || ix < 0 || ix >= win) return;
            float4 t = *reinterpret_cast<const float4*>(r.base + ((size_t)iy * win + ix) * cin + ci);
            if (r.scale) {
                float4 s = *reinterpret_cast<const float4*>(r.scale + ci);
                t.x *= s.x; t.y *= s.y; t.z *= s.z; t.w *= s.w;
            }
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            return;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int kk = kidx + j;
            if (kk >= K) continue;
            int tap = kk / cin, ci = kk - tap * cin;
            int ky = tap / k, kx = tap - ky * k;
            int iy = r.iy0 + ky, ix = r.ix0 + kx;
            if (iy < 0 || iy >= hin || ix < 0 || ix >= win) continue;
            float t = r.base[((size_t)iy * win + ix) * cin + ci];
            if (r.scale) t *= r.scale[ci];
            v[j] = t;
        }
    }
};

struct ConvEpilogue {
    const float* bias;
    const float* residual;
    float* out;
    int cout, act;
    __device__ void store4(int m, int n, const float a[4], int N) const {
        size_t o = (size_t)m * cout + n;
        if (n + 3 < N && (cout & 3) == 0) {
            float4 b = *reinterpret_cast<const float4*>(bias + n);
            float4 r;
            r.x = apply_act(a[0] + b.x, act); r.y = apply_act(a[1] + b.y, act);
            r.z = apply_act(a[2] + b.z, act); r.w = apply_act(a[3] + b.w, act);
            if (residual) {
                float4 q = *reinterpret_cast<const float4*>(residual + o);
                r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
            }
            *reinterpret_cast<float4*>(out + o) = r;
            return;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n + j >= N) break;
            float v = apply_act(a[j] + bias[n + j], act);
            if (residual) v += residual[o + j];
            out[o + j] = v;
        }
    }
};

struct FrameLoader {
    const float* x;     // normalised audio [B][S]
    int S, hop, n_frames;
    struct Row { const float* base; bool valid; };
    __device__ Row row(int m, int M) const {
        Row r;
        r.valid = m < M;
        int mm = r.valid ? m : 0;
        int b = mm / n_frames, t = mm - b * n_frames;
        r.base = x + (size_t)b * S + (size_t)t * hop;
        return r;
    }
    __device__ void load4(const Row& r, int kidx, int K, float v[4]) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (r.valid && kidx + j < K) ? r.base[kidx + j] : 0.f;
    }
};

struct SpecEpilogueV24 {
    float* spec;    // [B][n_mels][n_frames][n_ch]
    int n_frames, n_mels, n_ch, ch;
    float exponent;
    __device__ void store4(int m, int n, const float a[4], int N) const {
        int b = m / n_frames, t = m - b * n_frames;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n + j >= N) break;
            float p = a[j] * a[j];
            float v = p > 0.f ? powf(p, exponent) : (p == 0.f ? 0.f : p);   // NaN propagates
            spec[(((size_t)b * n_mels + (n + j)) * n_frames + t) * n_ch + ch] = v;
        }
    }
};

template <class Loader, class Epilogue>
__global__ void __launch_bounds__(256) k_igemm_f32(Loader ld, Epilogue ep, const float* __restrict__ W,
                                                   int ldw, int M, int N, int K) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int a_row = t >> 2, a_k = (t & 3) << 2;
    const int b_k = t >> 4, b_n = (t & 15) << 2;
    const int ty = t >> 4, tx = t & 15;
    typename Loader::Row row = ld.row(m0 + a_row, M);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        float av[4];
        ld.load4(row, k0 + a_k, K, av);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + b_k < K && n0 + b_n < ldw) bv = *reinterpret_cast<const float4*>(W + (size_t)(k0 + b_k) * ldw + n0 + b_n);
#pragma unroll
        for (int j = 0; j < 4; ++j) As[a_k + j][a_row] = av[j];
        *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty << 2]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx << 2]);
            const float aa[4] = {a.x, a.y, a.z, a.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + (ty << 2) + i;
        if (m < M && n0 + (tx << 2) < N) ep.store4(m, n0 + (tx << 2), acc[i], N);
    }
}

cudaError_t launch_conv_igemm(const ConvParams& p, cudaStream_t stream) {
    const long long M = (long long)p.batch * p.hout * p.wout;
    if (M <= 0) return cudaSuccess;
    const int K = p.k * p.k * p.cin, N = p.cout;
    ConvLoader ld{p.in, p.in_scale, p.hin, p.win, p.cin, p.hout, p.wout, p.k, p.stride, p.pad};
    ConvEpilogue ep{p.bias, p.residual, p.out, p.cout, p.act};
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    k_igemm_f32<<<grid, 256, 0, stream>>>(ld, ep, p.weight, p.ldw, (int)M, N, K);
    return cudaGetLastError();
}

cudaError_t launch_spectrogram_v24(const float* xnorm, const float* basis, int ldb, float* spec,
                                   int batch, int sample_count, int n_fft, int hop, int n_frames,
                                   int n_mels, int n_ch, int ch, float exponent, cudaStream_t stream) {
    const long long M = (long long)batch * n_frames;
    if (M <= 0) return cudaSuccess;
    FrameLoader ld{xnorm, sample_count, hop, n_frames};
    SpecEpilogueV24 ep{spec, n_frames, n_mels, n_ch, ch, exponent};
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((n_mels + BN - 1) / BN));
    k_igemm_f32<<<grid, 256, 0, stream>>>(ld, ep, basis, ldb, (int)M, n_mels, n_fft);
    return cudaGetLastError();
}

// ======================================================================================
// Depthwise conv + bias + activation, 4 channels per thread
// ======================================================================================
__global__ void __launch_bounds__(256) k_dwconv(DwParams p, long long total4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4n = p.c >> 2;
    int c4 = (int)(i % c4n);
    long long pix = i / c4n;
    int ox = (int)(pix % p.wout);
    long long t2 = pix / p.wout;
    int oy = (int)(t2 % p.hout);
    int b = (int)(t2 / p.hout);
    const int c = c4 << 2;
    float4 acc = *reinterpret_cast<const float4*>(p.bias + c);
    const float* inb = p.in + (size_t)b * p.hin * p.win * p.c;
    for (int ky = 0; ky < p.k; ++ky) {
        int iy = oy * p.stride - p.pad + ky;
        if (iy < 0 || iy >= p.hin) continue;
        for (int kx = 0; kx < p.k; ++kx) {
            int ix = ox * p.stride - p.pad + kx;
            if (ix < 0 || ix >= p.win) continue;
            float4 v = *reinterpret_cast<const float4*>(inb + ((size_t)iy * p.win + ix) * p.c + c);
            float4 w = *reinterpret_cast<const float4*>(p.weight + (size_t)(ky * p.k + kx) * p.c + c);
            acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
            acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
        }
    }
    acc.x = apply_act(acc.x, p.act); acc.y = apply_act(acc.y, p.act);
    acc.z = apply_act(acc.z, p.act); acc.w = apply_act(acc.w, p.act);
    *reinterpret_cast<float4*>(p.out + (size_t)pix * p.c + c) = acc;
}

cudaError_t launch_dwconv(const DwParams& p, cudaStream_t stream) {
    if (p.c & 3) return cudaErrorInvalidValue;
    long long total4 = (long long)p.batch * p.hout * p.wout * (p.c >> 2);
    if (total4 <= 0) return cudaSuccess;
    k_dwconv<<<(unsigned)((total4 + 255) / 256), 256, 0, stream>>>(p, total4);
    return cudaGetLastError();
}

// ======================================================================================
// Global average pool [B][hw][c] -> [B][c]
// ======================================================================================
__global__ void __launch_bounds__(128) k_gap(const float* __restrict__ in, float* __restrict__ out, int hw, int c) {
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const float* p = in + (size_t)blockIdx.y * hw * c + ch;
    float s = 0.f;
    for (int i = 0; i < hw; ++i) s += p[(size_t)i * c];
    out[(size_t)blockIdx.y * c + ch] = s / (float)hw;
}

cudaError_t launch_gap(const float* in, float* out, int batch, int hw, int c, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    dim3 grid((c + 127) / 128, batch);
    k_gap<<<grid, 128, 0, stream>>>(in, out, hw, c);
    return cudaGetLastError();
}

// ======================================================================================
// Epilogue: top-k by IEEE total order -> sigmoid -> min_confidence -> range mask / rerank
// (reference: src/postprocess.rs:40-93, src/rangefilter.rs:333-386)
// ======================================================================================
__device__ __forceinline__ uint32_t total_order_key(float x) {   // f32::total_cmp as unsigned order
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_total_order_key(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// in-place descending bitonic sort of P (power of two) 64-bit keys in shared memory
__device__ void bitonic_sort_desc(unsigned long long* keys, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                int lo = ((i / stride) * (stride << 1)) + (i % stride);
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// block-wide exclusive scan of one flag per thread-slot; returns (offset, total via smem)
__device__ int block_excl_scan(int flag, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned bal = __ballot_sync(0xffffffffu, flag);
    int pre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        int v = lane < nw ? s_warp[lane] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane < nw) s_warp[lane] = incl - v;
        if (lane == 31) s_warp[32] = incl;
    }
    __syncthreads();
    total = s_warp[32];
    int r = s_warp[warp] + pre;
    __syncthreads();
    return r;
}

// Shared tail: entries j = 0..count_in-1 (already in final order unless rerank) given by
// (idx_of[j], conf_of[j]) in shared memory; applies the range mask, compacts, reranks, writes out.
__device__ void filter_and_emit(uint32_t* s_idx, float* s_conf, unsigned long long* s_keys, int count_in,
                                const uint8_t* __restrict__ state, const float* __restrict__ score, int n,
                                int rerank, Pred* out, uint32_t* out_count, int* s_warp) {
    // pass 1: mask + compaction into s_keys as (conf_key << 32 | ~rank) where rank is the
    // compacted position, side arrays rewritten in place (compaction only moves entries down)
    int base = 0;
    for (int j0 = 0; j0 < count_in; j0 += blockDim.x) {
        int j = j0 + threadIdx.x;
        int keep = 0;
        uint32_t idx = 0;
        float conf = 0.f;
        if (j < count_in) {
            idx = s_idx[j];
            conf = s_conf[j];
            keep = 1;
            if (state && idx < (uint32_t)n) {
                uint8_t st = state[idx];
                if (st == 2) keep = 0;
                else if (st == 1 && rerank) conf = conf * score[idx];
            }
        }
        int total;
        int off = block_excl_scan(keep, s_warp, total);   // contains __syncthreads: reads above are done
        if (keep) {
            int r = base + off;
            s_idx[r] = idx;
            s_conf[r] = conf;
        }
        base += total;
        __syncthreads();
    }
    const int cnt = base;
    if (rerank && state && cnt > 1) {
        int P = 1;
        while (P < cnt) P <<= 1;
        for (int j = threadIdx.x; j < P; j += blockDim.x)
            s_keys[j] = j < cnt ? (((unsigned long long)total_order_key(s_conf[j]) << 32) | (0xFFFFFFFFu - (uint32_t)j)) : 0ull;
        bitonic_sort_desc(s_keys, P);
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
            uint32_t rank = 0xFFFFFFFFu - (uint32_t)(s_keys[j] & 0xFFFFFFFFull);
            Pred p;
            p.index = s_idx[rank];
            p.confidence = s_conf[rank];
            out[j] = p;
        }
    } else {
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
            Pred p;
            p.index = s_idx[j];
            p.confidence = s_conf[j];
            out[j] = p;
        }
    }
    if (threadIdx.x == 0) *out_count = (uint32_t)cnt;
}

__global__ void __launch_bounds__(1024) k_topk(TopkParams p, int P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw);
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_keys + P);
    float* s_conf = reinterpret_cast<float*>(s_idx + p.k);
    __shared__ int s_warp[33];
    const int b = blockIdx.x;
    const float* lg = p.logits + (size_t)b * p.n;
    if (p.k == 0) {
        if (threadIdx.x == 0) p.out_count[b] = 0;
        return;
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        s_keys[i] = i < p.n ? (((unsigned long long)total_order_key(lg[i]) << 32) | (0xFFFFFFFFu - (uint32_t)i)) : 0ull;
    bitonic_sort_desc(s_keys, P);
    // sigmoid on the k survivors, min_confidence: survivors stay in logit-descending order,
    // which is confidence-descending because sigmoid is monotone (ties: higher logit first)
    int base = 0;
    for (int j0 = 0; j0 < (int)p.k; j0 += blockDim.x) {
        int j = j0 + threadIdx.x;
        int keep = 0;
        uint32_t idx = 0;
        float conf = 0.f;
        if (j < (int)p.k) {
            unsigned long long kk = s_keys[j];
            idx = 0xFFFFFFFFu - (uint32_t)(kk & 0xFFFFFFFFull);
            float x = from_total_order_key((uint32_t)(kk >> 32));
            conf = 1.0f / (1.0f + expf(-x));
            keep = (!p.has_min_conf) || (conf >= p.min_conf);
        }
        int total;
        int off = block_excl_scan(keep, s_warp, total);
        if (keep) { s_idx[base + off] = idx; s_conf[base + off] = conf; }
        base += total;
    }
    __syncthreads();
    filter_and_emit(s_idx, s_conf, s_keys, base, p.range_state, p.range_score, p.n, p.rerank,
                    p.out + (size_t)b * p.k, p.out_count + b, s_warp);
}

__global__ void k_range_filter(const Pred* in, const uint32_t* in_count, int stride, const uint8_t* state,
                               const float* score, int n, int rerank, Pred* out, uint32_t* out_count, int P);

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

constexpr int kMaxSortSmem = 220 * 1024;

cudaError_t init_kernels_for_device() {
    cudaError_t e = cudaFuncSetAttribute(k_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSortSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_range_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSortSmem);
}

cudaError_t launch_topk(const TopkParams& p, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    const int P = next_pow2(p.n < 2 ? 2 : p.n);
    size_t smem = (size_t)P * 8 + (size_t)p.k * 8;
    if (smem > (size_t)kMaxSortSmem) return cudaErrorInvalidValue;
    int threads = P / 2 < 1024 ? (P / 2 < 32 ? 32 : P / 2) : 1024;
    k_topk<<<p.batch, threads, smem, stream>>>(p, P);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_range_filter(const Pred* in, const uint32_t* in_count, int stride,
                                                      const uint8_t* state, const float* score, int n, int rerank,
                                                      Pred* out, uint32_t* out_count, int P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw);
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_keys + P);
    float* s_conf = reinterpret_cast<float*>(s_idx + stride);
    __shared__ int s_warp[33];
    const int r = blockIdx.x;
    int cnt = (int)in_count[r];
    if (cnt > stride) cnt = stride;
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        Pred q = in[(size_t)r * stride + j];
        s_idx[j] = q.index;
        s_conf[j] = q.confidence;
    }
    __syncthreads();
    filter_and_emit(s_idx, s_conf, s_keys, cnt, state, score, n, rerank, out + (size_t)r * stride, out_count + r, s_warp);
}

cudaError_t launch_range_filter(const Pred* in, const uint32_t* in_count, int rows, int stride,
                                const uint8_t* state, const float* score, int n, int rerank,
                                Pred* out, uint32_t* out_count, cudaStream_t stream) {
    if (rows <= 0 || stride <= 0) return cudaSuccess;
    const int P = next_pow2(stride < 2 ? 2 : stride);
    size_t smem = (size_t)P * 8 + (size_t)stride * 8;
    if (smem > (size_t)kMaxSortSmem) return cudaErrorInvalidValue;
    k_range_filter<<<rows, 256, smem, stream>>>(in, in_count, stride, state, score, n, rerank, out, out_count, P);
    return cudaGetLastError();
}

}  // namespace bn
