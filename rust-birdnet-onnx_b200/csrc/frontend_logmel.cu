// Log-mel front-end of the 32 kHz graphs (BirdNET v3.0, Perch v2):
//
//   frames (Hann window, no centre padding, zero pad_end) -> |STFT| -> mel matrix -> log_scale * ln(. + log_floor)
//
// The magnitude sits between the DFT and the mel projection, so unlike the v2.4 front-end this is
// not one linear map and does not belong on the tensor cores: a 1024-point real FFT is ~25 kFLOP
// where the DFT-as-GEMM would be 2 MFLOP.  One CTA transforms FRAMES_PER_PASS frames at a time in
// shared memory:
//   * real FFT of length N as a complex Stockham FFT of length M = N/2 over z[n] = x[2n] + i x[2n+1]
//     (mixed radix 8/4/2/5/3, twiddles from a float64-built table staged in smem; frame buffers padded by one
//     element per 32 against the scatter's bank conflicts),
//   * split:  X[k] = E[k] + W_N^k O[k],  E = (Z[k] + conj Z[M-k])/2,  O = -i (Z[k] - conj Z[M-k])/2,
//   * |X[k]| for k = 0..N/2, then the mel projection as a banded sum (each triangular filter is a
//     contiguous run of bins), log, hi/lo fp16 split, 256-byte coalesced row stores.
// HBM: audio read once (overlapping frames hit L1/L2), spectrogram written once.
#include "kernels.h"

#include <cmath>
#include <vector>

namespace bn {

namespace {

constexpr int LM_THREADS = 256;
constexpr int LM_FRAMES = 4;        // frames transformed concurrently by one CTA
constexpr int LM_PASSES = 2;        // passes per CTA (twiddle table staged once)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// element i of a frame buffer lives at i + i/32: one pad float2 per 32 breaks the power-of-two strides of the Stockham
// scatter (stride R * 8 bytes in the first stage: 16-way bank conflicts unpadded, 2-way padded)
__device__ __forceinline__ int lm_skew(int i) { return i + (i >> 5); }
__host__ __device__ constexpr int lm_padded(int m) { return m + (m >> 5) + 1; }

constexpr int LM_TPF = LM_THREADS / LM_FRAMES;       // threads that work on one frame (64): frame = tid / 64, no division by
                                                     // run-time lengths anywhere in the loops below

// One Stockham stage of radix R over the CTA's frames (transforms of length M, one per 64 threads).
//   x[j + r*M/R] * W^(r * (j % Ns))  -> R-point DFT ->  y[(j / Ns) * Ns * R + j % Ns + r * Ns]
// tws = this stage's own twiddles [R-1][Ns] (W_(Ns*R)^(r*k), contiguous in k: conflict-free 64-bit reads; the strided
// reads of the full table were 8- to 16-way bank conflicts); tw = the full table exp(-2*pi*i*m / (2M)) for the R = 3 / 5
// butterflies.  POW2: Ns is a power of two (always true unless two odd radices follow each other).
template <int R, bool POW2>
__device__ __forceinline__ void stockham_stage(const float2* __restrict__ x, float2* __restrict__ y, const float2* __restrict__ tws,
                                               const float2* __restrict__ tw, int M, int Ns, int ns_shift, int F) {
    const int nb = M / R;
    const int f = threadIdx.x / LM_TPF;
    if (f >= F) return;
    const int MP = lm_padded(M);
    const float2* xf = x + f * MP;
    float2* yf = y + f * MP;
    for (int j = threadIdx.x % LM_TPF; j < nb; j += LM_TPF) {
        const int k = POW2 ? (j & (Ns - 1)) : (j % Ns);
        const int jq = POW2 ? (j >> ns_shift) : (j / Ns);
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = xf[lm_skew(j + r * nb)];
            if (r > 0 && Ns > 1) v[r] = cmul(v[r], tws[(r - 1) * Ns + k]);
        }
        float2 o[R];
        if (R == 2) {
            o[0] = make_float2(v[0].x + v[1].x, v[0].y + v[1].y);
            o[1] = make_float2(v[0].x - v[1].x, v[0].y - v[1].y);
        } else if (R == 8) {
            // 8-point DFT: two 4-point DFTs (even / odd inputs) + W8 twiddles; h = 1/sqrt(2)
            const float h = 0.70710678118654752f;
            auto add = [](float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); };
            auto sub = [](float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); };
            auto mi = [](float2 a) { return make_float2(a.y, -a.x); };                    // -i * a
            const float2 a0 = add(v[0], v[4]), a1 = sub(v[0], v[4]), a2 = add(v[2], v[6]), a3 = mi(sub(v[2], v[6]));
            const float2 a4 = add(v[1], v[5]), a5 = sub(v[1], v[5]), a6 = add(v[3], v[7]), a7 = mi(sub(v[3], v[7]));
            const float2 e0 = add(a0, a2), e2 = sub(a0, a2), e1 = add(a1, a3), e3 = sub(a1, a3);
            const float2 q0 = add(a4, a6), q2 = sub(a4, a6), q1 = add(a5, a7), q3 = sub(a5, a7);
            const float2 t1 = make_float2((q1.x + q1.y) * h, (q1.y - q1.x) * h);          // q1 * (1 - i) / sqrt 2
            const float2 t2 = mi(q2);
            const float2 t3 = make_float2((q3.y - q3.x) * h, -(q3.x + q3.y) * h);         // q3 * (-1 - i) / sqrt 2
            o[0] = add(e0, q0); o[4] = sub(e0, q0);
            o[1] = add(e1, t1); o[5] = sub(e1, t1);
            o[2] = add(e2, t2); o[6] = sub(e2, t2);
            o[3] = add(e3, t3); o[7] = sub(e3, t3);
        } else if (R == 4) {
            const float2 a = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
            const float2 b = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
            const float2 c = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
            const float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
            o[0] = make_float2(a.x + c.x, a.y + c.y);
            o[2] = make_float2(a.x - c.x, a.y - c.y);
            o[1] = make_float2(b.x + d.y, b.y - d.x);      // b - i d
            o[3] = make_float2(b.x - d.y, b.y + d.x);      // b + i d
        } else {
            // generic small DFT (R = 3, 5): o[q] = sum_p v[p] * W_R^(p q)
            const int rstep = (2 * M) / R;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                float2 s = v[0];
#pragma unroll
                for (int p = 1; p < R; ++p) {
                    const float2 t = cmul(v[p], tw[((p * q) % R) * rstep]);
                    s.x += t.x; s.y += t.y;
                }
                o[q] = s;
            }
        }
        const int d0 = jq * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) yf[lm_skew(d0 + r * Ns)] = o[r];
    }
}

template <int R>
__device__ __forceinline__ void stockham_dispatch(const float2* x, float2* y, const float2* tws, const float2* tw, int M, int Ns, int F) {
    if ((Ns & (Ns - 1)) == 0) stockham_stage<R, true>(x, y, tws, tw, M, Ns, 31 - __clz(Ns), F);
    else stockham_stage<R, false>(x, y, tws, tw, M, Ns, 0, F);
}

__global__ void __launch_bounds__(LM_THREADS) k_logmel(LogmelParams p) {
    extern __shared__ __align__(16) unsigned char lm_smem[];
    const int N = p.n_fft, M = N >> 1;
    float2* tw = reinterpret_cast<float2*>(lm_smem);             // [N]   full table
    float2* tws = tw + N;                                        // [M]   per-stage tables, stage with stride Ns at offset Ns - 1
    const int MP = lm_padded(M);
    float2* buf0 = tws + M;                                      // [LM_FRAMES][MP], element i at lm_skew(i)
    float2* buf1 = buf0 + LM_FRAMES * MP;                        // [LM_FRAMES][MP]
    for (int i = threadIdx.x; i < N; i += LM_THREADS) tw[i] = p.twiddle[i];
    {
        int Ns = 1;
        for (int st = 0; st < p.n_stages; ++st) {
            const int R = p.radix[st];
            const int tstep = (2 * M) / (Ns * R);
            // sum over earlier stages of Ns' * (R' - 1) telescopes to Ns - 1
            for (int i = threadIdx.x; i < Ns * (R - 1); i += LM_THREADS) {
                const int r = i / Ns + 1, k = i - (r - 1) * Ns;
                tws[Ns - 1 + i] = p.twiddle[r * k * tstep];
            }
            Ns *= R;
        }
    }
    const int b = blockIdx.y;
    const float* xs = p.audio + (size_t)b * p.sample_count;
    const int bins = M + 1;
    const int f = threadIdx.x / LM_TPF, l = threadIdx.x % LM_TPF;        // this thread's frame of the pass, lane inside it
    for (int pass = 0; pass < LM_PASSES; ++pass) {
        const int t0 = (blockIdx.x * LM_PASSES + pass) * LM_FRAMES;
        if (t0 >= p.n_frames) break;                             // uniform across the CTA
        const int F = min(LM_FRAMES, p.n_frames - t0);
        __syncthreads();                                         // twiddles staged / previous pass consumed
        // windowed frames, even samples -> re, odd samples -> im; zero beyond the segment (pad_end)
        if (f < F) {
            const int fs = (t0 + f) * p.hop;
            const bool vec = ((fs & 1) == 0) && fs + N <= p.sample_count;        // 8-byte aligned pairs, all inside the segment
            for (int n = l; n < M; n += LM_TPF) {
                const int s0 = fs + 2 * n;
                const float2 wv = __ldg(reinterpret_cast<const float2*>(p.window) + n);
                float2 v;
                if (vec) {
                    const float2 xv = __ldg(reinterpret_cast<const float2*>(xs + s0));
                    v = make_float2(xv.x * wv.x, xv.y * wv.y);
                } else {
                    v.x = s0 < p.sample_count ? __ldg(xs + s0) * wv.x : 0.f;
                    v.y = s0 + 1 < p.sample_count ? __ldg(xs + s0 + 1) * wv.y : 0.f;
                }
                buf0[f * MP + lm_skew(n)] = v;
            }
        }
        __syncthreads();
        float2* src = buf0;
        float2* dst = buf1;
        int Ns = 1;
        for (int st = 0; st < p.n_stages; ++st) {
            const int R = p.radix[st];
            const float2* ts = tws + (Ns - 1);
            if (R == 8) stockham_dispatch<8>(src, dst, ts, tw, M, Ns, F);
            else if (R == 4) stockham_dispatch<4>(src, dst, ts, tw, M, Ns, F);
            else if (R == 2) stockham_dispatch<2>(src, dst, ts, tw, M, Ns, F);
            else if (R == 5) stockham_dispatch<5>(src, dst, ts, tw, M, Ns, F);
            else stockham_dispatch<3>(src, dst, ts, tw, M, Ns, F);
            Ns *= R;
            __syncthreads();
            float2* t = src; src = dst; dst = t;
        }
        // |X[k]|, k = 0..M, into the free buffer as float mag[f][bins]
        float* mag = reinterpret_cast<float*>(dst);              // F * (M + 1) floats <= F * MP float2
        if (f < F) {
            const float2* zf = src + f * MP;
            for (int k = l; k < bins; k += LM_TPF) {
                const float2 z = zf[lm_skew(k == M ? 0 : k)];
                const float2 zc = zf[lm_skew(k == 0 || k == M ? 0 : M - k)];
                const float er = 0.5f * (z.x + zc.x), ei = 0.5f * (z.y - zc.y);     // E[k]
                const float orr = 0.5f * (z.y + zc.y), oi = -0.5f * (z.x - zc.x);   // O[k] = -i (Z - conj Zc) / 2
                const float2 wk = tw[k];
                const float xr = er + (orr * wk.x - oi * wk.y);
                const float xi = ei + (orr * wk.y + oi * wk.x);
                mag[f * bins + k] = sqrtf(xr * xr + xi * xi);
            }
        }
        __syncthreads();
        // Mel projection, one thread per filter (a banded dot product).  Sharing a filter between eight lanes (strided
        // partial sums + shuffles) was measured SLOWER (1.18 -> 1.91 ms): the phase is bound by the latency of the three
        // per-filter metadata loads, which the shared form repeats eight times as often per lane, not by the MAC count.
        if (f < F) {
            for (int m = l; m < p.n_mels; m += LM_TPF) {
                const int lo = __ldg(p.mel_lo + m), cnt = __ldg(p.mel_cnt + m);
                const float* wt = p.mel_w + __ldg(p.mel_off + m);
                const float* mg = mag + f * bins + lo;
                float s = 0.f;
                for (int i = 0; i < cnt; ++i) s = fmaf(mg[i], __ldg(wt + i), s);
                const float y = p.log_scale * logf(s + p.log_floor);
                const size_t o = ((size_t)b * p.n_frames + t0 + f) * p.n_mels + m;
                const __half h = __float2half_rn(y);
                p.out.hi[o] = h;
                p.out.hi[p.out.plane + o] = __float2half_rn(y - __half2float(h));
                if (p.out_f32) p.out_f32[o] = y;
            }
        }
    }
}

}  // namespace

size_t logmel_smem_bytes(int n_fft) { return (size_t)(n_fft + n_fft / 2 + 2 * LM_FRAMES * lm_padded(n_fft / 2)) * sizeof(float2); }

bool logmel_factorize(int n_fft, int radix[8], int* n_stages) {
    if (n_fft < 8 || (n_fft & 1)) return false;
    int m = n_fft / 2, n = 0;
    const int cand[5] = {8, 4, 2, 5, 3};          // largest power-of-two radix first: 512 = 8*8*8, 320 = 8*8*5
    for (int ci = 0; ci < 5; ++ci)
        while (m % cand[ci] == 0) {
            if (n >= 8) return false;
            radix[n++] = cand[ci];
            m /= cand[ci];
        }
    *n_stages = n;
    return m == 1;
}

void logmel_twiddles(int n_fft, std::vector<float>& out) {
    out.resize((size_t)2 * n_fft);
    for (int i = 0; i < n_fft; ++i) {
        const double a = -2.0 * M_PI * (double)i / (double)n_fft;
        out[2 * i] = (float)cos(a);
        out[2 * i + 1] = (float)sin(a);
    }
}

cudaError_t logmel_init_device() {
    return cudaFuncSetAttribute(k_logmel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

cudaError_t launch_logmel(const LogmelParams& p, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    const size_t smem = logmel_smem_bytes(p.n_fft);
    if (smem > 96 * 1024 || p.n_stages < 1 || p.n_stages > 8) return cudaErrorInvalidValue;
    const int per_cta = LM_FRAMES * LM_PASSES;
    dim3 grid((unsigned)((p.n_frames + per_cta - 1) / per_cta), (unsigned)p.batch);
    k_logmel<<<grid, LM_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace bn
