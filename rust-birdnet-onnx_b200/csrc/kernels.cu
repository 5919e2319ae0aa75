// sm_100a kernels, generation 1: correct FP32 CUDA-core path for every stage of the hot path
// (front-end, CNN, epilogue).  Tensor-core (tcgen05) replacements live in tc_*.cu and are
// validated against these.
#include "kernels.h"
#include "fast_act.cuh"

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

struct FeOuts { FePlaneOut o[2]; int n; };

__global__ void __launch_bounds__(1024) k_minmax_normalize_fe(const float* __restrict__ x, float* __restrict__ y, FeOuts fo,
                                                              int S, float eps, float half, float two) {
    const float* xs = x + (size_t)blockIdx.x * S;
    float* ys = y ? y + (size_t)blockIdx.x * S : nullptr;
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
    const float den = __fadd_rn(__fsub_rn(s_mx[0], lo), eps);
    if (ys) {
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
    for (int bi = 0; bi < fo.n; ++bi) {
        const FePlaneOut o = fo.o[bi];
        const int units_per_row = o.row_stride >> 3;
        const int n_units = o.rows * units_per_row;
        __half* hi = o.hi + (size_t)blockIdx.x * o.rows * o.row_stride;
        for (int u = threadIdx.x; u < n_units; u += blockDim.x) {
            const int r = u / units_per_row, c0 = (u - r * units_per_row) << 3;
            __half2 h[4], l[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                float v[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = c0 + 2 * e2 + q;
                    const int src = r * o.hop + c;
                    v[q] = (c < o.hop && src < S)
                               ? __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(xs[src], lo), den), half), two) : 0.f;
                }
                h[e2] = __floats2half2_rn(v[0], v[1]);
                float2 bk = __half22float2(h[e2]);
                l[e2] = __floats2half2_rn(v[0] - bk.x, v[1] - bk.y);
            }
            *reinterpret_cast<uint4*>(hi + (size_t)u * 8) = *reinterpret_cast<uint4*>(h);
            *reinterpret_cast<uint4*>(hi + o.plane + (size_t)u * 8) = *reinterpret_cast<uint4*>(l);
        }
    }
}

// ---- two-kernel form: (1) per-segment min / max over SLICES CTAs with order-preserving integer
// atomics (min / max are order independent -> deterministic), (2) normalise + emit, SLICES CTAs
// per segment.  Keeps all 148 SMs busy for B = 256 (a 1-CTA-per-segment kernel runs 1.7 waves).
constexpr int NORM_SLICES = 8;

__device__ __forceinline__ uint32_t f2key(float x) {
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256) k_minmax_init(uint32_t* mm, int batch) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) { mm[2 * i] = 0xFFFFFFFFu; mm[2 * i + 1] = 0u; }
}

__global__ void __launch_bounds__(512) k_minmax_partial(const float* __restrict__ x, uint32_t* __restrict__ mm, int S) {
    const int b = blockIdx.y;
    const float4* x4 = reinterpret_cast<const float4*>(x + (size_t)b * S);
    const int n4 = S >> 2;
    const int per = (n4 + gridDim.x - 1) / gridDim.x;
    const int beg = blockIdx.x * per, end = min(n4, beg + per);
    float mn = INFINITY, mx = -INFINITY;
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
        float4 v = x4[i];
        mn = nan_min(nan_min(mn, v.x), nan_min(nan_min(v.y, v.z), v.w));
        mx = nan_max(nan_max(mx, v.x), nan_max(nan_max(v.y, v.z), v.w));
    }
    if (blockIdx.x == gridDim.x - 1)
        for (int i = (n4 << 2) + threadIdx.x; i < S; i += blockDim.x) {
            const float v = x[(size_t)b * S + i];
            mn = nan_min(mn, v);
            mx = nan_max(mx, v);
        }
    for (int o = 16; o > 0; o >>= 1) {
        mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // nan_min / nan_max propagate NaN, so a NaN anywhere in the slice shows up here
        if (mn != mn || mx != mx) {
            atomicMax(mm + 2 * b + 1, 0xFFC00000u);   // canonical +NaN key: poisons the segment
        } else {
            atomicMin(mm + 2 * b, f2key(mn));
            atomicMax(mm + 2 * b + 1, f2key(mx));
        }
    }
}

__global__ void __launch_bounds__(512) k_normalize_emit(const float* __restrict__ x, float* __restrict__ y, const uint32_t* __restrict__ mm,
                                                        FeOuts fo, int S, float eps, float half, float two) {
    const int b = blockIdx.y;
    const float* xs = x + (size_t)b * S;
    float lo = key2f(mm[2 * b]);
    float hi_v = key2f(mm[2 * b + 1]);
    if (hi_v != hi_v) lo = hi_v;                       // any NaN poisons the whole segment, like torch
    const float den = __fadd_rn(__fsub_rn(hi_v, lo), eps);
    if (y) {
        float* ys = y + (size_t)b * S;
        const int n4 = S >> 2;
        const int per = (n4 + gridDim.x - 1) / gridDim.x;
        const int beg = blockIdx.x * per, end = min(n4, beg + per);
        const float4* x4 = reinterpret_cast<const float4*>(xs);
        float4* y4 = reinterpret_cast<float4*>(ys);
        for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
            float4 v = x4[i], r;
            r.x = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.x, lo), den), half), two);
            r.y = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.y, lo), den), half), two);
            r.z = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.z, lo), den), half), two);
            r.w = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(v.w, lo), den), half), two);
            y4[i] = r;
        }
        if (blockIdx.x == gridDim.x - 1)
            for (int i = (n4 << 2) + threadIdx.x; i < S; i += blockDim.x)
                ys[i] = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(xs[i], lo), den), half), two);
    }
    for (int bi = 0; bi < fo.n; ++bi) {
        const FePlaneOut o = fo.o[bi];
        const int units_per_row = o.row_stride >> 3;
        const int n_units = o.rows * units_per_row;
        const int per = (n_units + gridDim.x - 1) / gridDim.x;
        const int beg = blockIdx.x * per, end = min(n_units, beg + per);
        __half* hi = o.hi + (size_t)b * o.rows * o.row_stride;
        for (int u = beg + threadIdx.x; u < end; u += blockDim.x) {
            const int r = u / units_per_row, c0 = (u - r * units_per_row) << 3;
            __half2 h[4], l[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                float v[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = c0 + 2 * e2 + q;
                    const int src = r * o.hop + c;
                    v[q] = (c < o.hop && src < S)
                               ? __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(xs[src], lo), den), half), two) : 0.f;
                }
                h[e2] = __floats2half2_rn(v[0], v[1]);
                float2 bk = __half22float2(h[e2]);
                l[e2] = __floats2half2_rn(v[0] - bk.x, v[1] - bk.y);
            }
            *reinterpret_cast<uint4*>(hi + (size_t)u * 8) = *reinterpret_cast<uint4*>(h);
            *reinterpret_cast<uint4*>(hi + o.plane + (size_t)u * 8) = *reinterpret_cast<uint4*>(l);
        }
    }
}

cudaError_t launch_minmax_normalize_fe(const float* x, float* y, uint32_t* minmax_scratch, const FePlaneOut* outs, int n_outs,
                                       int batch, int sample_count, float eps, float half, float two, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    if (n_outs > 2) return cudaErrorInvalidValue;
    FeOuts fo{};
    fo.n = n_outs;
    for (int i = 0; i < n_outs; ++i) fo.o[i] = outs[i];
    k_minmax_init<<<(batch + 255) / 256, 256, 0, stream>>>(minmax_scratch, batch);
    dim3 grid(NORM_SLICES, batch);
    k_minmax_partial<<<grid, 512, 0, stream>>>(x, minmax_scratch, sample_count);
    if (y != nullptr || n_outs > 0) k_normalize_emit<<<grid, 512, 0, stream>>>(x, y, minmax_scratch, fo, sample_count, eps, half, two);
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

// ---- planes variants of the loaders / epilogues -------------------------------------------
__device__ __forceinline__ void split_store(__half* hi, size_t plane, size_t o, float v) {
    __half h = __float2half_rn(v);
    hi[o] = h;
    hi[plane + o] = __float2half_rn(v - __half2float(h));
}
__device__ __forceinline__ float planes_load(const __half* hi, size_t plane, size_t o) {
    return __half2float(hi[o]) + __half2float(hi[plane + o]);
}

struct ConvLoaderPlanes {
    const __half* in;
    size_t plane;
    const float* in_scale;
    int hin, win, cin, hout, wout, k, stride, pad;
    struct Row {
        size_t base;
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
        r.base = (size_t)b * hin * win * cin;
        r.scale = in_scale ? in_scale + (size_t)b * cin : nullptr;
        return r;
    }
    __device__ void load4(const Row& r, int kidx, int K, float v[4]) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = 0.f;
            int kk = kidx + j;
            if (!r.valid || kk >= K) continue;
            int tap = kk / cin, ci = kk - tap * cin;
            int ky = tap / k, kx = tap - ky * k;
            int iy = r.iy0 + ky, ix = r.ix0 + kx;
            if (iy < 0 || iy >= hin || ix < 0 || ix >= win) continue;
            float t = planes_load(in, plane, r.base + ((size_t)iy * win + ix) * cin + ci);
            if (r.scale) t *= r.scale[ci];
            v[j] = t;
        }
    }
};

struct ConvEpiloguePlanes {
    const float* bias;
    const __half* res;
    size_t res_plane;
    __half* out;
    size_t out_plane;
    int cout, act;
    __device__ void store4(int m, int n, const float a[4], int N) const {
        size_t o = (size_t)m * cout + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n + j >= N) break;
            float v = apply_act(a[j] + bias[n + j], act);
            if (res) v += planes_load(res, res_plane, o + j);
            split_store(out, out_plane, o + j, v);
        }
    }
};

struct SpecEpilogueV24Planes {
    __half* spec;
    size_t plane;
    int n_frames, n_mels, n_ch, ch;
    float exponent;
    __device__ void store4(int m, int n, const float a[4], int N) const {
        int b = m / n_frames, t = m - b * n_frames;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n + j >= N) break;
            float p = a[j] * a[j];
            float v = p > 0.f ? powf(p, exponent) : (p == 0.f ? 0.f : p);
            split_store(spec, plane, (((size_t)b * n_mels + (n + j)) * n_frames + t) * n_ch + ch, v);
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

cudaError_t launch_conv_igemm_planes(const ConvPlanesParams& p, cudaStream_t stream) {
    const long long M = (long long)p.batch * p.hout * p.wout;
    if (M <= 0) return cudaSuccess;
    const int K = p.k * p.k * p.cin, N = p.cout;
    ConvLoaderPlanes ld{p.in.hi, p.in.plane, p.in_scale, p.hin, p.win, p.cin, p.hout, p.wout, p.k, p.stride, p.pad};
    ConvEpiloguePlanes ep{p.bias, p.residual.hi, p.residual.plane, p.out.hi, p.out.plane, p.cout, p.act};
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    k_igemm_f32<<<grid, 256, 0, stream>>>(ld, ep, p.weight, p.ldw, (int)M, N, K);
    return cudaGetLastError();
}

cudaError_t launch_spectrogram_v24_planes(const float* xnorm, const float* basis, int ldb, PlanesPtr spec,
                                          int batch, int sample_count, int n_fft, int hop, int n_frames,
                                          int n_mels, int n_ch, int ch, float exponent, cudaStream_t stream) {
    const long long M = (long long)batch * n_frames;
    if (M <= 0) return cudaSuccess;
    FrameLoader ld{xnorm, sample_count, hop, n_frames};
    SpecEpilogueV24Planes ep{spec.hi, spec.plane, n_frames, n_mels, n_ch, ch, exponent};
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((n_mels + BN - 1) / BN));
    k_igemm_f32<<<grid, 256, 0, stream>>>(ld, ep, basis, ldb, (int)M, n_mels, n_fft);
    return cudaGetLastError();
}

// ======================================================================================
// Depthwise conv on planes.  One CTA per (segment, group of CG channels): the whole input
// patch of the group is staged in shared memory as FP32 (hi + lo) with 16-byte loads, every
// thread then owns one channel and a strided set of output pixels.  The global average of the
// activated output (the squeeze of squeeze-excite) is reduced in a fixed order -> deterministic.
// ======================================================================================
template <int K, int STRIDE, int CG>
__global__ void __launch_bounds__(256) k_dwconv_planes(DwPlanesParams p) {
    extern __shared__ __align__(16) float s_in[];          // [(hin+2p)][(win+2p)][CG], zero halo
    __shared__ float s_part[256];
    constexpr int PG = 256 / CG;                            // threads per channel
    constexpr int XB = STRIDE == 1 ? 4 : 2;                 // outputs per thread along x
    constexpr int NCOL = (XB - 1) * STRIDE + K;             // input columns one x-block touches
    const int cgroups = p.c / CG;
    const int b = blockIdx.x / cgroups, cg = blockIdx.x - b * cgroups;
    const int c0 = cg * CG;
    const int hp = p.hin + 2 * p.pad, wp = p.win + 2 * p.pad;
    const int npout = p.hout * p.wout;
    const size_t in_base = (size_t)b * p.hin * p.win * p.c + c0;
    constexpr int UPP = CG / 8;                             // 16-byte units per pixel and plane
    for (int u = threadIdx.x; u < hp * wp * UPP; u += 256) {
        const int pix = u / UPP, cu = u - pix * UPP;
        const int y = pix / wp - p.pad, x = pix - (pix / wp) * wp - p.pad;
        float* dst = s_in + (size_t)pix * CG + cu * 8;
        if (y >= 0 && y < p.hin && x >= 0 && x < p.win) {
            const size_t o = in_base + ((size_t)y * p.win + x) * p.c + cu * 8;
            uint4 qh = __ldg(reinterpret_cast<const uint4*>(p.in.hi + o));
            uint4 ql = __ldg(reinterpret_cast<const uint4*>(p.in.hi + p.in.plane + o));
            const __half2* h = reinterpret_cast<const __half2*>(&qh);
            const __half2* l = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 a = __half22float2(h[i]), d = __half22float2(l[i]);
                dst[2 * i] = a.x + d.x;
                dst[2 * i + 1] = a.y + d.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = 0.f;
        }
    }
    const int cl = threadIdx.x % CG, pg = threadIdx.x / CG;
    const int c = c0 + cl;
    float wreg[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) wreg[i] = p.weight[(size_t)i * p.c + c];
    const float bias = p.bias[c];
    __syncthreads();
    float pool = 0.f;
    const int xblocks = p.wout / XB;
    for (int blk = pg; blk < p.hout * xblocks; blk += PG) {
        const int oy = blk / xblocks, ox0 = (blk - oy * xblocks) * XB;
        float acc[XB];
#pragma unroll
        for (int j = 0; j < XB; ++j) acc[j] = bias;
        const float* base = s_in + ((size_t)(oy * STRIDE) * wp + ox0 * STRIDE) * CG + cl;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            float col[NCOL];
#pragma unroll
            for (int x = 0; x < NCOL; ++x) col[x] = base[((size_t)ky * wp + x) * CG];
#pragma unroll
            for (int j = 0; j < XB; ++j)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) acc[j] = fmaf(col[j * STRIDE + kx], wreg[ky * K + kx], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < XB; ++j) {
            const float v = apply_act(acc[j], p.act);
            split_store(p.out.hi, p.out.plane, ((size_t)b * npout + oy * p.wout + ox0 + j) * p.c + c, v);
            pool += v;
        }
    }
    if (p.pooled) {
        s_part[threadIdx.x] = pool;
        __syncthreads();
        if (threadIdx.x < CG) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < PG; ++g) s += s_part[g * CG + threadIdx.x];
            p.pooled[(size_t)b * p.c + c0 + threadIdx.x] = s / (float)npout;
        }
    }
}

template <int K, int STRIDE>
static cudaError_t launch_dw_ks(const DwPlanesParams& p, int cg, size_t smem, cudaStream_t stream) {
    const int grid = p.batch * (p.c / cg);
    if (cg == 64) k_dwconv_planes<K, STRIDE, 64><<<grid, 256, smem, stream>>>(p);
    else if (cg == 32) k_dwconv_planes<K, STRIDE, 32><<<grid, 256, smem, stream>>>(p);
    else k_dwconv_planes<K, STRIDE, 16><<<grid, 256, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int K, int STRIDE>
static void dw_set_attr() {
    cudaFuncSetAttribute(k_dwconv_planes<K, STRIDE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_dwconv_planes<K, STRIDE, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_dwconv_planes<K, STRIDE, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t launch_dwconv_planes(const DwPlanesParams& p, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    if ((p.k != 3 && p.k != 5) || (p.stride != 1 && p.stride != 2) || (p.c & 15) || p.pad != p.k / 2) return cudaErrorInvalidValue;
    const int xb = p.stride == 1 ? 4 : 2;
    if (p.wout % xb) return cudaErrorInvalidValue;
    const size_t np = (size_t)(p.hin + 2 * p.pad) * (p.win + 2 * p.pad);
    // largest channel group whose zero-padded FP32 patch fits in ~100 KB (two CTAs per SM)
    int cg = 64;
    while (cg > 16 && ((p.c % cg) != 0 || np * cg * 4 > 100 * 1024)) cg >>= 1;
    if ((p.c % cg) != 0 || np * cg * 4 > 200 * 1024) return cudaErrorInvalidValue;
    const size_t smem = np * cg * 4;
    if (p.k == 3 && p.stride == 1) return launch_dw_ks<3, 1>(p, cg, smem, stream);
    if (p.k == 3) return launch_dw_ks<3, 2>(p, cg, smem, stream);
    if (p.stride == 1) return launch_dw_ks<5, 1>(p, cg, smem, stream);
    return launch_dw_ks<5, 2>(p, cg, smem, stream);
}

// ======================================================================================
// Squeeze-excite tail, one kernel: r = silu(W1^T p + b1); s = sigmoid(W2^T r + b2); then the
// depthwise output is rescaled IN PLACE (d <- d * s) so the projection conv can stream it with
// cp.async like any other tensor.  grid = (chunks, B); every CTA recomputes the two tiny FCs.
// ======================================================================================
__global__ void __launch_bounds__(256) k_se_scale(SeParams p) {
    extern __shared__ float s_se[];                 // pooled[C] | gate[C] | r[R] | partial[8][R]
    float* s_p = s_se;
    float* s_g = s_se + p.c;
    float* s_r = s_g + p.c;
    float* s_part = s_r + p.r;
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < p.c; i += 256) s_p[i] = p.pooled[(size_t)b * p.c + i];
    __syncthreads();
    // FC1: warps split the C range, lanes = output j (coalesced rows of W1[c][ldw1])
    const int cpw = (p.c + 7) / 8;
    const int cbeg = warp * cpw, cend = min(p.c, cbeg + cpw);
    for (int j0 = 0; j0 < p.r; j0 += 32) {
        const int j = j0 + lane;
        float acc = 0.f;
        if (j < p.r) {
            for (int cc = cbeg; cc < cend; ++cc) acc = fmaf(s_p[cc], p.w1[(size_t)cc * p.ldw1 + j], acc);
            s_part[warp * p.r + j] = acc;
        }
    }
    __syncthreads();
    for (int j = tid; j < p.r; j += 256) {
        float v = p.b1[j];
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_part[w * p.r + j];
        s_r[j] = v * (1.0f / (1.0f + expf(-v)));
    }
    __syncthreads();
    // FC2: thread per channel, W2[j][ldw2] rows are contiguous in c
    for (int cc = tid; cc < p.c; cc += 256) {
        float v = p.b2[cc];
        for (int j = 0; j < p.r; ++j) v = fmaf(s_r[j], p.w2[(size_t)j * p.ldw2 + cc], v);
        const float g = 1.0f / (1.0f + expf(-v));
        s_g[cc] = g;
        if (p.gate_out && blockIdx.x == 0) p.gate_out[(size_t)b * p.c + cc] = g;
    }
    __syncthreads();
    // rescale this CTA's share of the pixels, 8 channels (16 bytes per plane) per thread-step
    const int upp = p.c >> 3;
    const int total_u = p.npix * upp;
    const int per = (total_u + gridDim.x - 1) / gridDim.x;
    const int ubeg = blockIdx.x * per, uend = min(total_u, ubeg + per);
    __half* base = p.d.hi + (size_t)b * p.npix * p.c;
    for (int u = ubeg + tid; u < uend; u += 256) {
        const int cu = (u % upp) << 3;
        __half* ph = base + (size_t)u * 8;
        uint4 qh = *reinterpret_cast<const uint4*>(ph);
        uint4 ql = *reinterpret_cast<const uint4*>(ph + p.d.plane);
        const __half2* h = reinterpret_cast<const __half2*>(&qh);
        const __half2* l = reinterpret_cast<const __half2*>(&ql);
        __half2 oh[4], ol[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 a = __half22float2(h[i]), d = __half22float2(l[i]);
            const float v0 = (a.x + d.x) * s_g[cu + 2 * i], v1 = (a.y + d.y) * s_g[cu + 2 * i + 1];
            oh[i] = __floats2half2_rn(v0, v1);
            float2 bk = __half22float2(oh[i]);
            ol[i] = __floats2half2_rn(v0 - bk.x, v1 - bk.y);
        }
        *reinterpret_cast<uint4*>(ph) = *reinterpret_cast<uint4*>(oh);
        *reinterpret_cast<uint4*>(ph + p.d.plane) = *reinterpret_cast<uint4*>(ol);
    }
}

cudaError_t launch_se_scale(const SeParams& p, int batch, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    if ((p.c & 7) || p.r > 256) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)2 * p.c + p.r + 8 * p.r) * sizeof(float);
    int chunks = (p.npix * (p.c >> 3) + 4095) / 4096;      // ~16 units per thread
    if (chunks < 1) chunks = 1;
    if (chunks > 16) chunks = 16;
    dim3 grid(chunks, batch);
    k_se_scale<<<grid, 256, smem, stream>>>(p);
    return cudaGetLastError();
}

// Squeeze-excite gate: gate[b][c] = sigmoid(W2^T silu(W1^T pooled[b] + b1) + b2); one CTA per segment.
// Pure L2 latency (two dependent FCs over ~150-450 KB of weights), so every load loop keeps 8
// independent loads in flight.
__global__ void __launch_bounds__(256) k_se_gate(const float* __restrict__ pooled, const float* __restrict__ w1,
                                                 const float* __restrict__ b1, const float* __restrict__ w2,
                                                 const float* __restrict__ b2, float* __restrict__ gate,
                                                 int c, int r, int ldw1, int ldw2) {
    extern __shared__ float s_g[];                          // pooled[C] | r[R] | partial[8][R]
    float* s_p = s_g;
    float* s_r = s_p + c;
    float* s_pt = s_r + r;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < c; i += 256) s_p[i] = pooled[(size_t)b * c + i];
    __syncthreads();
    const int cpw = (c + 7) / 8;                            // FC1: warps split C, lanes = output j
    const int cbeg = warp * cpw, cend = min(c, cbeg + cpw);
    for (int j0 = 0; j0 < r; j0 += 32) {
        const int j = j0 + lane;
        if (j < r) {
            float acc = 0.f;
            int cc = cbeg;
            for (; cc + 8 <= cend; cc += 8) {
                float w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) w[u] = __ldg(w1 + (size_t)(cc + u) * ldw1 + j);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = fmaf(s_p[cc + u], w[u], acc);
            }
            for (; cc < cend; ++cc) acc = fmaf(s_p[cc], __ldg(w1 + (size_t)cc * ldw1 + j), acc);
            s_pt[warp * r + j] = acc;
        }
    }
    __syncthreads();
    for (int j = tid; j < r; j += 256) {
        float v = b1[j];
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_pt[w * r + j];
        s_r[j] = v * (1.0f / (1.0f + expf(-v)));
    }
    __syncthreads();
    for (int cc = tid; cc < c; cc += 256) {                 // FC2: thread per channel, rows of W2 contiguous in c
        float v = b2[cc];
        int j = 0;
        for (; j + 8 <= r; j += 8) {
            float w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = __ldg(w2 + (size_t)(j + u) * ldw2 + cc);
#pragma unroll
            for (int u = 0; u < 8; ++u) v = fmaf(s_r[j + u], w[u], v);
        }
        for (; j < r; ++j) v = fmaf(s_r[j], __ldg(w2 + (size_t)j * ldw2 + cc), v);
        gate[(size_t)b * c + cc] = 1.0f / (1.0f + expf(-v));
    }
}

cudaError_t launch_se_gate(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                           float* gate, int batch, int c, int r, int ldw1, int ldw2, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    if (r > 256) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)c + r + 8 * r) * sizeof(float);
    k_se_gate<<<batch, 256, smem, stream>>>(pooled, w1, b1, w2, b2, gate, c, r, ldw1, ldw2);
    return cudaGetLastError();
}

// Pure streaming squeeze-excite rescale: d[b][pix][c] *= gate[b][c], hi/lo planes in place.
__global__ void __launch_bounds__(256) k_se_rescale(__half* __restrict__ d, size_t plane, const float* __restrict__ gate,
                                                    int c, unsigned ups, unsigned long long total) {
    const unsigned upp = (unsigned)c >> 3;                  // 16-byte units per pixel
    const unsigned long long stride = (unsigned long long)gridDim.x * 256ull;
    for (unsigned long long u = (unsigned long long)blockIdx.x * 256ull + threadIdx.x; u < total; u += stride) {
        const unsigned b = (unsigned)(u / ups);
        const unsigned cu = ((unsigned)(u % upp)) << 3;
        __half* ph = d + u * 8ull;
        const uint4 qh = *reinterpret_cast<const uint4*>(ph);
        const uint4 ql = *reinterpret_cast<const uint4*>(ph + plane);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * c + cu));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * c + cu) + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const __half2* h = reinterpret_cast<const __half2*>(&qh);
        const __half2* l = reinterpret_cast<const __half2*>(&ql);
        __half2 oh[4], ol[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 a = __half22float2(h[i]), e = __half22float2(l[i]);
            const float v0 = (a.x + e.x) * g[2 * i], v1 = (a.y + e.y) * g[2 * i + 1];
            oh[i] = __floats2half2_rn(v0, v1);
            const float2 bk = __half22float2(oh[i]);
            ol[i] = __floats2half2_rn(v0 - bk.x, v1 - bk.y);
        }
        *reinterpret_cast<uint4*>(ph) = *reinterpret_cast<uint4*>(oh);
        *reinterpret_cast<uint4*>(ph + plane) = *reinterpret_cast<uint4*>(ol);
    }
}

cudaError_t launch_se_rescale(PlanesPtr d, const float* gate, int batch, int npix, int c, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    if (c & 7) return cudaErrorInvalidValue;
    const unsigned ups = (unsigned)npix * (unsigned)(c >> 3);
    const unsigned long long total = (unsigned long long)batch * ups;
    unsigned long long blocks = (total + 1023) / 1024;       // ~4 units per thread
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    if (blocks < 1) blocks = 1;
    k_se_rescale<<<(unsigned)blocks, 256, 0, stream>>>(d.hi, d.plane, gate, c, ups, total);
    return cudaGetLastError();
}

// ======================================================================================
// CLI ingest on the device: 16-bit PCM -> FP32 / 32768 (read_wav,
// src/bin/birdnet-analyze.rs:21, 684-687) and chunk_audio (707-743): segment b starts at sample
// first_pos + b * step, samples past the end of the recording are zeros.  One thread = 4 samples.
// ======================================================================================
__global__ void __launch_bounds__(256) k_pcm16_to_segments(const int16_t* __restrict__ pcm, unsigned long long base,
                                                           unsigned long long n_total, unsigned long long first_pos,
                                                           unsigned long long step, float* __restrict__ out, int S) {
    const int b = blockIdx.y;
    const int i0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= S) return;
    const unsigned long long pos = first_pos + (unsigned long long)b * step + (unsigned long long)i0;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long q = pos + j;
        v[j] = (i0 + j < S && q < n_total) ? (float)__ldg(pcm + (q - base)) * (1.0f / 32768.0f) : 0.f;   // exact: power-of-two scale
    }
    float* dst = out + (size_t)b * S + i0;
    if (i0 + 3 < S) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else for (int j = 0; j < 4 && i0 + j < S; ++j) dst[j] = v[j];
}

cudaError_t launch_pcm16_to_segments(const int16_t* pcm, uint64_t base, uint64_t n_total, uint64_t first_pos, uint64_t step,
                                     float* out, int batch, int sample_count, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    if (sample_count & 3) return cudaErrorInvalidValue;          // float4 row alignment
    dim3 grid((unsigned)((sample_count / 4 + 255) / 256), (unsigned)batch);
    k_pcm16_to_segments<<<grid, 256, 0, stream>>>(pcm, base, n_total, first_pos, step, out, sample_count);
    return cudaGetLastError();
}

// ======================================================================================
// Stem: direct k x k conv for tiny Cin (the 2-channel spectrogram), planes in / planes out.
// One thread per output pixel computes all Cout (<= 32) channels; weights live in smem.
// ======================================================================================
// CTA = one strip of STRIP output pixels of one output row; the 3 input rows it needs are staged in smem
// as FP32 with every value DUPLICATED (x, x), which is the multiplicand layout of the packed FFMA2
// (fma.rn.f32x2): thread = (pixel slot, channel quad), its 9*cin x 4 weights live in registers as two
// f32x2 pairs per tap, so a tap costs one LDS and two (cin = 1) or four (cin = 2) FFMA2.
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float silu_fast_k(float v) { return silu_approx(v); }

template <int CIN, int STRIP>
__global__ void __launch_bounds__(256, 2) k_stem_planes(ConvPlanesParams p) {
    constexpr int K = 3, STRIDE = 2;
    constexpr int WIN = (STRIP - 1) * STRIDE + K;               // input columns per strip
    __shared__ __align__(16) float s_x[K][WIN * CIN * 2];
    const int quads = p.cout >> 2;                               // 8 for cout = 32
    const int slots = 256 / quads;                               // pixel slots per pass
    const int strips = (p.wout + STRIP - 1) / STRIP;
    const int q = threadIdx.x % quads, slot = threadIdx.x / quads;
    unsigned long long w[K * K * CIN][2];
#pragma unroll
    for (int t = 0; t < K * K * CIN; ++t) {
        const float4 wv = *reinterpret_cast<const float4*>(p.weight + (size_t)t * p.ldw + q * 4);
        w[t][0] = pack_f2(wv.x, wv.y);
        w[t][1] = pack_f2(wv.z, wv.w);
    }
    const float4 bv = *reinterpret_cast<const float4*>(p.bias + q * 4);
    const unsigned long long b01 = pack_f2(bv.x, bv.y), b23 = pack_f2(bv.z, bv.w);
    const bool silu = p.act == KACT_SILU;
    const int total = p.batch * p.hout * strips;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {     // persistent: weights stay in registers
    const int strip = work % strips;
    const int oy = (work / strips) % p.hout;
    const int b = work / (strips * p.hout);
    const int ox0 = strip * STRIP;
    const int ix_base = ox0 * STRIDE - p.pad, iy_base = oy * STRIDE - p.pad;
    const size_t in_base = (size_t)b * p.hin * p.win * CIN;
    __syncthreads();                                                   // previous strip fully consumed
    // stage K rows x WIN pixels: one 4-byte (2-channel) or 2-byte (1-channel) load per pixel and
    // plane, all loads of a thread issued before any shared-memory store so their latencies overlap
    static_assert(CIN == 1 || CIN == 2, "pixel = one 16/32-bit word per plane");
    constexpr int NPIX = K * WIN;
    constexpr int NLD = (NPIX + 255) / 256;
    uint32_t vh[NLD], vl[NLD];
    const uint32_t* ph32 = reinterpret_cast<const uint32_t*>(p.in.hi);
    const uint32_t* pl32 = reinterpret_cast<const uint32_t*>(p.in.hi + p.in.plane);
    const unsigned short* ph16 = reinterpret_cast<const unsigned short*>(p.in.hi);
    const unsigned short* pl16 = reinterpret_cast<const unsigned short*>(p.in.hi + p.in.plane);
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
        const int i = threadIdx.x + j * 256;
        const int ky = i / WIN, x = i - ky * WIN;
        const int iy = iy_base + ky, ix = ix_base + x;
        vh[j] = 0u; vl[j] = 0u;
        if (i < NPIX && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win) {
            if (CIN == 2) {
                const size_t o = (in_base >> 1) + (size_t)iy * p.win + ix;
                vh[j] = __ldg(ph32 + o);
                vl[j] = __ldg(pl32 + o);
            } else {
                const size_t o = in_base + (size_t)iy * p.win + ix;
                vh[j] = __ldg(ph16 + o);
                vl[j] = __ldg(pl16 + o);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < NPIX) {
            const int ky = i / WIN, x = i - ky * WIN;
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&vh[j]));
            const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&vl[j]));
            if (CIN == 2) {
                const float x0 = a.x + d.x, x1 = a.y + d.y;
                *reinterpret_cast<float4*>(&s_x[ky][x * 4]) = make_float4(x0, x0, x1, x1);
            } else {
                const float x0 = a.x + d.x;
                *reinterpret_cast<float2*>(&s_x[ky][x * 2]) = make_float2(x0, x0);
            }
        }
    }
    __syncthreads();
    for (int px = slot; px < STRIP && ox0 + px < p.wout; px += slots) {
        unsigned long long a01 = b01, a23 = b23;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int t = (ky * K + kx) * CIN;
                if (CIN == 2) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&s_x[ky][(px * STRIDE + kx) * 4]);
                    a01 = ffma2(v.x, w[t][0], a01);
                    a23 = ffma2(v.x, w[t][1], a23);
                    a01 = ffma2(v.y, w[t + 1][0], a01);
                    a23 = ffma2(v.y, w[t + 1][1], a23);
                } else {
                    const unsigned long long v = *reinterpret_cast<const unsigned long long*>(&s_x[ky][(px * STRIDE + kx) * 2]);
                    a01 = ffma2(v, w[t][0], a01);
                    a23 = ffma2(v, w[t][1], a23);
                }
            }
        float acc[4] = {__uint_as_float((uint32_t)a01), __uint_as_float((uint32_t)(a01 >> 32)),
                        __uint_as_float((uint32_t)a23), __uint_as_float((uint32_t)(a23 >> 32))};
        if (silu) {
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[n] = silu_fast_k(acc[n]);
        } else {
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[n] = apply_act(acc[n], p.act);
        }
        const __half2 h0 = __floats2half2_rn(acc[0], acc[1]), h1 = __floats2half2_rn(acc[2], acc[3]);
        const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(acc[0] - b0.x, acc[1] - b0.y), l1 = __floats2half2_rn(acc[2] - b1.x, acc[3] - b1.y);
        const size_t o = (((size_t)b * p.hout + oy) * p.wout + ox0 + px) * p.cout + q * 4;
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
        lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(p.out.hi + o) = hv;
        *reinterpret_cast<uint2*>(p.out.hi + p.out.plane + o) = lv;
    }
    }
}

cudaError_t launch_stem_planes(const ConvPlanesParams& p, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    if (p.k != 3 || (p.cin != 2 && p.cin != 1) || p.stride != 2 || (p.cout & 3) || p.cout > 64 || (256 % (p.cout >> 2)) || p.in_scale || p.residual.hi)
        return cudaErrorInvalidValue;
    const int strip = p.wout > 128 ? 256 : (p.wout > 64 ? 128 : 64);
    const int strips = (p.wout + strip - 1) / strip;
    const long long total = (long long)p.batch * p.hout * strips;
    const int grid = (int)(total < 148 * 2 ? total : 148 * 2);
    if (p.cin == 2) {
        if (strip == 256) k_stem_planes<2, 256><<<grid, 256, 0, stream>>>(p);
        else if (strip == 128) k_stem_planes<2, 128><<<grid, 256, 0, stream>>>(p);
        else k_stem_planes<2, 64><<<grid, 256, 0, stream>>>(p);
    } else {
        if (strip == 256) k_stem_planes<1, 256><<<grid, 256, 0, stream>>>(p);
        else if (strip == 128) k_stem_planes<1, 128><<<grid, 256, 0, stream>>>(p);
        else k_stem_planes<1, 64><<<grid, 256, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

__global__ void __launch_bounds__(128) k_gap_planes(const __half* __restrict__ in, size_t plane, float* __restrict__ out, int hw, int c) {
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const size_t base = (size_t)blockIdx.y * hw * c + ch;
    float s = 0.f;
    for (int i = 0; i < hw; ++i) s += planes_load(in, plane, base + (size_t)i * c);
    out[(size_t)blockIdx.y * c + ch] = s / (float)hw;
}

cudaError_t launch_gap_planes(PlanesPtr in, float* out, int batch, int hw, int c, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    dim3 grid((c + 127) / 128, batch);
    k_gap_planes<<<grid, 128, 0, stream>>>(in.hi, in.plane, out, hw, c);
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
__device__ __forceinline__ uint32_t total_order_key_bits(uint32_t u) {   // f32::total_cmp as unsigned order
    // opaque to the optimiser: seen as float bits, `u | sign` is rewritten into FADD -|x|, which canonicalises a NaN
    // (payload and ordering lost; caught by the +NaN known-answer test of postprocess.rs:233-242)
    asm("" : "+r"(u));
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint32_t total_order_key(float x) { return total_order_key_bits(__float_as_uint(x)); }
__device__ __forceinline__ bool bits_nonfinite(uint32_t u) { return (u & 0x7f800000u) == 0x7f800000u; }
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

// ws == nullptr: key / index / confidence arrays in dynamic shared memory (next_pow2(n)*8 + k*8 bytes must fit);
// otherwise they live in the global workspace ws, (P*8 + k*8) bytes per row (any n, any k: postprocess.rs:50 has no limit)
__global__ void __launch_bounds__(1024) k_topk(TopkParams p, int P, unsigned char* ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* base_mem = ws ? ws + (size_t)blockIdx.x * ((size_t)P * 8 + (size_t)p.k * 8) : smem_raw;
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(base_mem);
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_keys + P);
    float* s_conf = reinterpret_cast<float*>(s_idx + p.k);
    __shared__ int s_warp[33];
    const int b = blockIdx.x;
    const float* lg = p.logits + (size_t)b * p.n;
    if (p.k == 0) {
        if (threadIdx.x == 0) p.out_count[b] = 0;
        return;
    }
    int bad = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const uint32_t u = i < p.n ? __ldg(reinterpret_cast<const uint32_t*>(lg) + i) : 0u;     // bits, never a float value
        bad |= bits_nonfinite(u);
        s_keys[i] = i < p.n ? (((unsigned long long)total_order_key_bits(u) << 32) | (0xFFFFFFFFu - (uint32_t)i)) : 0ull;
    }
    if (p.nonfinite) {
        bad = __syncthreads_or(bad);
        if (bad && threadIdx.x == 0) atomicAdd(p.nonfinite, 1u);
    }
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
                               const float* score, int n, int rerank, Pred* out, uint32_t* out_count, int P, unsigned char* ws);

// Small-k path (k <= 32): k rounds of block-wide arg-max over register-resident keys instead of a
// full sort.  Keys are (total-order(logit) << 32 | ~index), so every key is unique and the order
// is the documented one (higher logit first, lower index first on ties).
constexpr int TOPK_SMALL_EPT = 32;
__global__ void __launch_bounds__(1024) k_topk_small(TopkParams p) {
    __shared__ unsigned long long s_wmax[32];
    __shared__ unsigned long long s_best;
    __shared__ unsigned long long s_keys[64];
    __shared__ uint32_t s_idx[32];
    __shared__ float s_conf[32];
    __shared__ int s_warp[33];
    __shared__ int s_cnt;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const float* lg = p.logits + (size_t)b * p.n;
    unsigned long long keys[TOPK_SMALL_EPT];
    int bad = 0;
#pragma unroll
    for (int j = 0; j < TOPK_SMALL_EPT; ++j) {
        const int i = tid + j * blockDim.x;
        const uint32_t u = i < p.n ? __ldg(reinterpret_cast<const uint32_t*>(lg) + i) : 0u;     // bits, never a float value
        bad |= bits_nonfinite(u);
        keys[j] = i < p.n ? (((unsigned long long)total_order_key_bits(u) << 32) | (0xFFFFFFFFu - (uint32_t)i)) : 0ull;
    }
    if (p.nonfinite) {            // a segment whose logits are not all finite: bn_ctx_nonfinite_segments()
        bad = __syncthreads_or(bad);
        if (bad && tid == 0) atomicAdd(p.nonfinite, 1u);
    }
    for (uint32_t r = 0; r < p.k; ++r) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int j = 0; j < TOPK_SMALL_EPT; ++j) m = keys[j] > m ? keys[j] : m;
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
            m = t > m ? t : m;
        }
        if (lane == 0) s_wmax[warp] = m;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = lane < nw ? s_wmax[lane] : 0ull;
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long u = __shfl_xor_sync(0xffffffffu, t, o);
                t = u > t ? u : t;
            }
            if (lane == 0) { s_best = t; s_keys[r] = t; }
        }
        __syncthreads();
        const unsigned long long best = s_best;
#pragma unroll
        for (int j = 0; j < TOPK_SMALL_EPT; ++j)
            if (keys[j] == best) keys[j] = 0ull;
    }
    // sigmoid + min_confidence on the k survivors (one warp; order is already confidence-descending)
    if (warp == 0) {
        uint32_t idx = 0;
        float conf = 0.f;
        int keep = 0;
        if (lane < (int)p.k) {
            const unsigned long long kk = s_keys[lane];
            idx = 0xFFFFFFFFu - (uint32_t)(kk & 0xFFFFFFFFull);
            const float x = from_total_order_key((uint32_t)(kk >> 32));
            conf = 1.0f / (1.0f + expf(-x));
            keep = (!p.has_min_conf) || (conf >= p.min_conf);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = __popc(bal & ((1u << lane) - 1u));
            s_idx[pos] = idx;
            s_conf[pos] = conf;
        }
        if (lane == 0) s_cnt = __popc(bal);
    }
    __syncthreads();
    filter_and_emit(s_idx, s_conf, s_keys, s_cnt, p.range_state, p.range_score, p.n, p.rerank,
                    p.out + (size_t)b * p.k, p.out_count + b, s_warp);
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

constexpr int kMaxSortSmem = 220 * 1024;

cudaError_t init_kernels_for_device() {
    dw_set_attr<3, 1>(); dw_set_attr<3, 2>(); dw_set_attr<5, 1>(); dw_set_attr<5, 2>();
    cudaError_t e = cudaFuncSetAttribute(k_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSortSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_range_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSortSmem);
}

cudaError_t launch_topk(const TopkParams& p, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    if (p.k >= 1 && p.k <= 32 && p.n <= TOPK_SMALL_EPT * 1024) {
        int threads = ((p.n + TOPK_SMALL_EPT - 1) / TOPK_SMALL_EPT + 31) / 32 * 32;
        if (threads < 64) threads = 64;
        k_topk_small<<<p.batch, threads, 0, stream>>>(p);
        return cudaGetLastError();
    }
    const int P = next_pow2(p.n < 2 ? 2 : p.n);
    size_t smem = (size_t)P * 8 + (size_t)p.k * 8;
    int threads = P / 2 < 1024 ? (P / 2 < 32 ? 32 : P / 2) : 1024;
    if (smem > (size_t)kMaxSortSmem) {
        // too large for shared memory (n > 16 K with k > 32, or n > 32 K): same kernel over a stream-ordered global
        // workspace.  Slower (the sort runs out of L2), but the reference's top_k_predictions has no size limit.
        unsigned char* ws = nullptr;
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), smem * (size_t)p.batch, stream);
        if (e != cudaSuccess) return e;
        k_topk<<<p.batch, threads, 0, stream>>>(p, P, ws);
        e = cudaGetLastError();
        const cudaError_t e2 = cudaFreeAsync(ws, stream);
        return e != cudaSuccess ? e : e2;
    }
    k_topk<<<p.batch, threads, smem, stream>>>(p, P, nullptr);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_range_filter(const Pred* in, const uint32_t* in_count, int stride,
                                                      const uint8_t* state, const float* score, int n, int rerank,
                                                      Pred* out, uint32_t* out_count, int P, unsigned char* ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* base_mem = ws ? ws + (size_t)blockIdx.x * ((size_t)P * 8 + (size_t)stride * 8) : smem_raw;
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(base_mem);
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
    if (smem > (size_t)kMaxSortSmem) {             // lists longer than ~14 K entries: global workspace (see launch_topk)
        unsigned char* ws = nullptr;
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), smem * (size_t)rows, stream);
        if (e != cudaSuccess) return e;
        k_range_filter<<<rows, 256, 0, stream>>>(in, in_count, stride, state, score, n, rerank, out, out_count, P, ws);
        e = cudaGetLastError();
        const cudaError_t e2 = cudaFreeAsync(ws, stream);
        return e != cudaSuccess ? e : e2;
    }
    k_range_filter<<<rows, 256, smem, stream>>>(in, in_count, stride, state, score, n, rerank, out, out_count, P, nullptr);
    return cudaGetLastError();
}

}  // namespace bn
