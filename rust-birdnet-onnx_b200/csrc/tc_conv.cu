// Implicit-GEMM convolution / dense layer on the 5th-gen tensor cores (tcgen05 + TMEM).
//
//   D[M = B*Hout*Wout][N = Cout] = A[M][K = k*k*Cin] * W[K][N]      (+bias, activation, +residual)
//
// Precision policy "FP32-equivalent": operands are fp16 hi + fp16 lo pairs (x = hi + lo to ~22
// bits) and every K step issues three kind::f16 MMAs (hi*hi + hi*lo + lo*hi) into one FP32 TMEM
// accumulator; the dropped lo*lo term is ~2^-22 relative.  See DESIGN.md "precision policy".
//
// Persistent, warp-specialised CTA (416 threads, one or two per SM), static round-robin tile schedule:
//   warps 0-3  A producers: im2col gather of the 128-row tile straight from the hi/lo planes with
//              16-byte cp.async (zero-fill for padding) into SWIZZLE_128B K-major smem; thread 0
//              also issues one 1-D bulk copy (TMA engine) per K chunk for the pre-swizzled weights
//   warp  4    owns TMEM; lane 0 issues tcgen05.mma and commits stage/accumulator barriers
//   warps 5-12 epilogue: tcgen05.ld -> bias / SiLU / sigmoid / residual -> split -> global planes
// Two TMEM accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
// smem ring: STAGES x { A_hi 16 KB | A_lo 16 KB | W_hi NT*128 B | W_lo NT*128 B }.
#include "tc_common.cuh"
#include "tc_conv.h"
#include "fast_act.cuh"

#include <cstdio>
#include <cstdlib>

namespace bn {
using namespace tc;

#ifndef BN_EPI_FMA_RCP
#define BN_EPI_FMA_RCP 0          // 1: SiLU reciprocal by Newton steps on the FMA pipe (one MUFU op instead of two); measured 1-4 % SLOWER: these phases are issue-bound, not MUFU-bound
#endif
constexpr int TM = 128;           // UMMA M (one TMEM lane per output row)
constexpr int KC = 64;            // K elements per smem stage = one 128-byte swizzle row of fp16
constexpr int MAX_STAGES = 6;
constexpr int MAX_K_CHUNKS = 64;     // K <= 4096
constexpr int MAX_SMEM_BIAS = 256;      // layers with more output channels read the bias through L1
constexpr int EPI_STAGE_BYTES = 4096;    // per epilogue warp: 32 rows x 32 FP32 columns
constexpr int A_TILE_BYTES = TM * 128;
constexpr int HALO_W = TM + 2;               // patch columns (one 128-pixel run of an image row + halo)
constexpr int HALO_PIX = 3 * HALO_W;           // patch pixels: 3 input rows
constexpr int MAX_HALO_SLOTS = 3;
constexpr int MAX_EPI_WARPS = 16;    // one CTA per SM: 8 or 16 epilogue warps; two CTAs per SM: 4 each (warps 5.. of the CTA)
constexpr int NTHREADS = (5 + MAX_EPI_WARPS) * 32;
constexpr int NTHREADS_PW8 = (9 + MAX_EPI_WARPS) * 32;      // 8 producer warps (cp.async gather modes, one CTA per SM)
static int g_epi_warps_one = 16;     // BN_EPI_WARPS=8|16 (development knob)

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == KACT_SILU) return silu_approx(v);
    if (act == KACT_SIGMOID) return sigmoid_approx(v);
    return v;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& q, float v[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// ---- development aid: where each role of CTA 0 spends its cycles --------------------------------
//  [0] producer: wait for a free ring slot   [1] producer: total loop
//  [2] MMA: wait acc_empty   [3] MMA: wait stage full   [4] MMA: total loop
//  [5] epilogue warp 5: wait acc_full   [6] epilogue warp 5: total loop   [7] tiles of CTA 0
__device__ unsigned long long g_tc_prof[128 * 16];
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long& acc, bool on) {
    if (!on) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}

// TWO: built for two co-resident CTAs per SM (288 threads, <= 113 registers) instead of one (up to 672 threads)
// PW: producer warps (4, or 8 for the cp.async gather modes: their per-chunk issue chain - table lookup, mask test,
// 16 copies - is what bounds the deep-K and 3x3 layers, and it parallelises over rows).  Warp PW issues the MMAs,
// warps PW+1.. are the epilogue.
template <int MODE, bool TWO = false, int PW = 4>
__global__ void __launch_bounds__(TWO ? 288 : (PW == 8 ? NTHREADS_PW8 : NTHREADS), TWO ? 2 : 1)
k_tc_conv(const __grid_constant__ TcConvParams p) {
    static_assert(PW == 4 || (PW == 8 && (MODE == TC_IN_PLANES || MODE == TC_IN_PLANES_SCALED || MODE == TC_IN_F32) && !TWO), "8 producer warps: gather modes only");
    constexpr int RPT = 32 / PW;                // rows per producer thread
    constexpr int RSTEP = 4 * PW;               // row distance between a thread's rows
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
    __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[4];
    __shared__ __align__(8) uint64_t bar_acc_empty[4];
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(8) uint64_t bar_w;         // HALO: resident weights landed
    __shared__ __align__(8) uint64_t halo_db[18];
    __shared__ uint32_t halo_a16[18];
    __shared__ uint32_t tap_tab[MAX_K_CHUNKS * 8];   // (element offset << 5) | tap bit
    __shared__ __align__(16) float s_bias[MAX_SMEM_BIAS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = p.nt, STAGES = p.stages;
    const int nthreads = (int)blockDim.x;
    const int epi_warps = (nthreads >> 5) - (PW + 1);
    // narrow tiles (NT <= 32) with 8 epilogue warps: the per-tile epilogue is a latency chain, so the two
    // warp sets take alternate tiles (accumulator a <-> set a) instead of splitting the few columns
    const bool tile_split = epi_warps >= 8 && NT <= 32;
    // accumulators in TMEM: two, or four when the tiles are narrow and every one of the four epilogue warp sets can own
    // one (a set then serves every fourth tile: the MMA thread no longer waits for a free accumulator)
    const uint32_t acc_shift = (tile_split && epi_warps >= 16 && p.tmem_cols >= 8 * NT) ? 2u : 1u;
    const uint32_t n_acc = 1u << acc_shift;
    const bool prof = p.prof != nullptr && blockIdx.x == 0;
    unsigned long long pw0 = 0, pw1 = 0;
    const long long prof_t0 = prof ? clock64() : 0;
    const uint32_t w_bytes = 2u * (uint32_t)NT * 128u;
    const bool w_res = MODE == TC_IN_TMA && p.w_resident != 0;        // weights of this CTA's n tile resident, ring = A tiles only
    const uint32_t stage_bytes = 2u * A_TILE_BYTES + (w_res ? 0u : w_bytes);
    const int total_tiles = p.m_tiles * p.n_tiles;
    // aligned by pointer arithmetic on the shared array (a uintptr_t round trip loses the address space: generic LD / ST)
    uint8_t* tiles0 = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    uint8_t* tiles = tiles0 + (w_res ? (size_t)p.k_chunks * w_bytes : 0);     // ring base (resident weights, if any, sit in front of it)

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            // cp.async modes: 128 producer threads + thread 0's expect_tx arrive; TMA: thread 0 only
            mbar_init(&bar_full[s], MODE == TC_IN_TMA ? 1 : (MODE == TC_IN_HALO ? 128 : PW * 32 + 1));
            mbar_init(&bar_empty[s], 1);       // tcgen05.commit
        }
        mbar_init(&bar_w, 1);
        for (uint32_t a = 0; a < n_acc; ++a) {
            mbar_init(&bar_acc_full[a], 1);    // tcgen05.commit
            mbar_init(&bar_acc_empty[a], tile_split ? (uint32_t)epi_warps >> acc_shift : (uint32_t)epi_warps);   // one arrive per epilogue warp serving it
        }
        fence_barrier_init();
    }
    if (warp == PW) tmem_alloc(&tmem_holder, p.tmem_cols);
    // K unit -> (tap element offset, tap bit) table, shared by all tiles of this CTA
    for (int u = tid; u < p.k_chunks * 8; u += nthreads) {
        const int k0 = u * 8;
        uint32_t e = 31u;                          // tap 31 is never set in a row mask -> zero fill
        if (k0 < p.K) {
            const int tap = k0 / p.tab_cin, ci = k0 - tap * p.tab_cin;
            const int ky = tap / p.k, kx = tap - ky * p.k;
            e = ((uint32_t)((ky * p.win + kx) * p.pix_stride + ci) << 5) | (uint32_t)tap;
        }
        tap_tab[u] = e;
    }
    const bool s_bias_ok = p.bias != nullptr && p.cout <= MAX_SMEM_BIAS;
    if (s_bias_ok)
        for (int i = tid; i < p.cout; i += nthreads) s_bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    // HALO geometry: [resident W: k_chunks x w_bytes][slot 0][slot 1][slot 2]; a slot holds the hi and
    // the lo patch, each [cin/8][HALO_PIX] 16-byte cells (channel-group major, pixels at 16-byte pitch)
    const uint32_t halo_plane_bytes = (uint32_t)(p.cin >> 3) * HALO_PIX * 16u;
    const uint32_t halo_slot_bytes = (2u * halo_plane_bytes + 1023u) & ~1023u;
    uint8_t* halo_slots = tiles + (size_t)p.k_chunks * w_bytes;
    // the epilogue staging tiles start after the ring
    const size_t ring_bytes = MODE == TC_IN_HALO ? (size_t)p.k_chunks * w_bytes + (size_t)STAGES * halo_slot_bytes
                                                 : (size_t)STAGES * stage_bytes;      // measured from `tiles`
    if (MODE == TC_IN_HALO && warp < PW) {
        // ================================ halo-patch producers ================================
        const int n_tile = blockIdx.x % p.n_tiles;               // fixed per CTA (grid % n_tiles == 0)
        if (tid == 0) {
            asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}"
                         ::"r"(smem_u32(&bar_w)), "r"((uint32_t)p.k_chunks * w_bytes) : "memory");
            for (int kc = 0; kc < p.k_chunks; ++kc) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) + ((size_t)n_tile * p.k_chunks + kc) * w_bytes;
                bulk_copy_g2s(tiles + (size_t)kc * w_bytes, src, w_bytes, &bar_w);
            }
        }
        // The cell -> (smem offset, global offset relative to the tile's first pixel, border flags) map is
        // the same for every tile: computed once, kept in registers (cin <= 32 -> at most 13 cells / thread).
        const int J = p.cin >> 3;
        const int cells = J * HALO_PIX;                          // 16-byte cells per plane
        constexpr int HC = 13;
        uint32_t soff[HC];
        int32_t goff[HC];
        uint32_t flg[HC];                                        // 1: row 0, 2: row 2, 4: col 0, 8: last col, 16: no cell
#pragma unroll
        for (int i = 0; i < HC; ++i) {
            const int e = tid + 128 * i;
            const int pix = e / J, j = e - pix * J;              // consecutive lanes: the J cells of a pixel, then the next pixel
            const int r = pix / HALO_W, x = pix - r * HALO_W;
            soff[i] = (uint32_t)(j * HALO_PIX + pix) * 16u;
            goff[i] = ((r - 1) * p.win + (x - 1)) * p.cin + j * 8;
            flg[i] = (r == 0 ? 1u : 0u) | (r == 2 ? 2u : 0u) | (x == 0 ? 4u : 0u) | (x == HALO_W - 1 ? 8u : 0u) | (e >= cells ? 16u : 0u);
        }
        const __half* in_lo = p.in_hi + p.in_plane;
        const int hw = p.hout * p.wout;
        uint32_t s = 0, ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int m0 = (t / p.n_tiles) * TM;
            const int b = m0 / hw, rem = m0 - b * hw;
            const int oy = rem / p.wout, ox0 = rem - oy * p.wout;
            const uint32_t bad = (oy == 0 ? 1u : 0u) | (oy == p.hin - 1 ? 2u : 0u) | (ox0 == 0 ? 4u : 0u) |
                                 (ox0 + TM == p.win ? 8u : 0u) | 16u;
            const int tile_base = b * p.seg_stride + (oy * p.win + ox0) * p.cin;
            mbar_wait_t(&bar_empty[s], ph ^ 1u, pw0, prof);
            const uint32_t d0 = smem_u32(halo_slots + (size_t)s * halo_slot_bytes);
#pragma unroll
            for (int i = 0; i < HC; ++i) {
                const bool ok = (flg[i] & bad) == 0u;
                const int eo = ok ? tile_base + goff[i] : 0;     // masked cells never form an address
                const uint32_t nb = ok ? 16u : 0u;
                if ((flg[i] & 16u) == 0u) {
                    cp_async16(d0 + soff[i], p.in_hi + eo, nb);
                    cp_async16(d0 + soff[i] + halo_plane_bytes, in_lo + eo, nb);
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_full[s])) : "memory");
            if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
        }
        cp_async_wait_all();
        if (prof && tid == 0) { p.prof[0] = pw0; p.prof[1] = (unsigned long long)(clock64() - prof_t0); }
    } else if (MODE == TC_IN_TMA && warp < PW) {
        // ================================ TMA producer (one thread) ================================
        if (warp == 0 && elect_one()) {
            tma_prefetch_desc(&p.tmap);
            if (w_res) {
                const int n_res = blockIdx.x % p.n_tiles;                 // fixed per CTA (grid % n_tiles == 0)
                asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}"
                             ::"r"(smem_u32(&bar_w)), "r"((uint32_t)p.k_chunks * w_bytes) : "memory");
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) + ((size_t)n_res * p.k_chunks + kc) * w_bytes;
                    bulk_copy_g2s(tiles0 + (size_t)kc * w_bytes, src, w_bytes, &bar_w);
                }
            }
            const uint32_t sub_bytes = (uint32_t)TM * (uint32_t)p.kb * 2u;       // one [128][kb] fp16 box
            const int subs_per_chunk = KC / p.kb;
            uint32_t s = 0, ph = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int mt = t / p.n_tiles, n_tile = t - mt * p.n_tiles;
                const int b = mt / p.tiles_per_seg, lt = mt - b * p.tiles_per_seg;
                int x0, y0, seg;
                if (p.flat) {
                    x0 = p.tiles_per_seg == p.m_tiles ? mt * TM : lt * TM;     // flat matrix row / frame index
                    y0 = 0;
                    seg = p.tiles_per_seg == p.m_tiles ? 0 : b;
                } else {
                    const int lp0 = lt * TM;
                    const int oy0 = lp0 / p.wout, ox0 = lp0 - oy0 * p.wout;
                    x0 = ox0 * p.stride - p.pad;
                    y0 = oy0 * p.stride - p.pad;
                    seg = b;
                }
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&bar_empty[s], ph ^ 1u);
                    uint8_t* st = tiles + (size_t)s * stage_bytes;
                    int nsub = (p.K - kc * KC + p.kb - 1) / p.kb;
                    if (nsub > subs_per_chunk) nsub = subs_per_chunk;
                    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}"
                                 ::"r"(smem_u32(&bar_full[s])), "r"((w_res ? 0u : w_bytes) + 2u * (uint32_t)nsub * sub_bytes) : "memory");
                    for (int sb = 0; sb < nsub; ++sb) {
                        const int k0 = kc * KC + sb * p.kb;
                        const int tap = k0 / p.tab_cin, ci = k0 - tap * p.tab_cin;
                        const int ky = tap / p.k, kx = tap - ky * p.k;
                        tma_load_5d(st + (size_t)sb * sub_bytes, &p.tmap, ci, x0 + kx, y0 + ky, seg, 0, &bar_full[s]);
                        tma_load_5d(st + A_TILE_BYTES + (size_t)sb * sub_bytes, &p.tmap, ci, x0 + kx, y0 + ky, seg, 1, &bar_full[s]);
                    }
                    if (!w_res) {
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) + ((size_t)n_tile * p.k_chunks + kc) * w_bytes;
                        bulk_copy_g2s(st + 2 * A_TILE_BYTES, src, w_bytes, &bar_full[s]);
                    }
                    if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp < PW) {
        // ================================ A / W producers ================================
        // Per thread: one 16-byte K unit (8 channels) of RPT rows (rbase + RSTEP*i).  Everything that
        // does not depend on the K chunk is hoisted to tile setup: per row a 32-bit element offset
        // of tap (0,0) and a bit mask of the taps that fall inside the image; per (chunk, unit) the
        // tap's element offset and bit index come from the smem table built above.
        const int unit = tid & 7;
        const int rbase = tid >> 3;
        const int hw = p.hout * p.wout;
        const uint32_t dst0 = sw128_offset((uint32_t)rbase, (uint32_t)unit);   // row i: + i * RSTEP * 128 (RSTEP is a multiple of 8 rows)
        const __half* in_lo = p.in_hi + p.in_plane;
        uint32_t s = 0, ph = 0;                // ring slot / phase of the chunk being issued
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int m0 = (t / p.n_tiles) * TM;
            const int n_tile = t - (t / p.n_tiles) * p.n_tiles;
            int32_t pix_off[RPT];              // element offset of input pixel (iy0, ix0), channel 0
            uint32_t tmask[RPT];               // bit (ky*k + kx) set <=> tap inside the image
            uint32_t seg_idx[RPT];
            {
                int m = m0 + rbase;
                int b = m / hw, rem = m - b * hw;
                int oy = rem / p.wout, ox = rem - oy * p.wout;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    if (m + RSTEP * i < p.M) {
                        const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
                        // taps kx in [xl, xh) and ky in [yl, yh) fall inside the image
                        const int xl = max(0, -ix0), xh = min(p.k, p.win - ix0);
                        const int yl = max(0, -iy0), yh = min(p.k, p.hin - iy0);
                        const uint32_t xm = xh > xl ? ((1u << xh) - 1u) & ~((1u << xl) - 1u) : 0u;
                        uint32_t yrep = 0;             // sum of 1 << (ky*k) over valid ky (k <= 5)
#pragma unroll
                        for (int ky = 0; ky < 5; ++ky)
                            if (ky >= yl && ky < yh) yrep |= 1u << (ky * p.k);
                        tmask[i] = xm * yrep;          // xm < 2^k: the shifted copies do not overlap
                        pix_off[i] = b * p.seg_stride + (iy0 * p.win + ix0) * p.pix_stride;
                        seg_idx[i] = (uint32_t)b;
                    } else {
                        tmask[i] = 0; pix_off[i] = 0; seg_idx[i] = 0;
                    }
                    ox += RSTEP;
                    while (ox >= p.wout) { ox -= p.wout; ++oy; }
                    while (oy >= p.hout) { oy -= p.hout; ++b; }
                }
            }
            for (int kc = 0; kc < p.k_chunks; ++kc) {
                mbar_wait_t(&bar_empty[s], ph ^ 1u, pw0, prof);
                uint8_t* st = tiles + (size_t)s * stage_bytes;
                if (warp == 0 && elect_one()) {
                    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}"
                                 ::"r"(smem_u32(&bar_full[s])), "r"(w_bytes) : "memory");
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) +
                                         ((size_t)n_tile * p.k_chunks + kc) * w_bytes;
                    bulk_copy_g2s(st + 2 * A_TILE_BYTES, src, w_bytes, &bar_full[s]);
                }
                int2 te;                                  // x: element offset of (tap, ci); y: tap bit (31 = beyond K)
                { const uint32_t tt = tap_tab[kc * 8 + unit]; te.x = (int)(tt >> 5); te.y = (int)(tt & 31u); }
                if (MODE == TC_IN_PLANES) {
                    const uint32_t d = smem_u32(st) + dst0;
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const bool ok = (tmask[i] >> te.y) & 1u;
                        const int eo = ok ? pix_off[i] + te.x : 0;      // clamp: masked taps never form an address
                        const uint32_t nb = ok ? 16u : 0u;
                        cp_async16(d + (uint32_t)i * (RSTEP * 128u), p.in_hi + eo, nb);
                        cp_async16(d + (uint32_t)i * (RSTEP * 128u) + A_TILE_BYTES, in_lo + eo, nb);
                    }
                    // the hardware arrives on the stage barrier once this thread's copies have landed:
                    // no wait in the producer, up to STAGES chunks in flight
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_full[s])) : "memory");
                } else {
                    float vals[RPT][8];
                    const int ci = te.x - ((int)(te.y == 31 ? 0 : te.y) / p.k * p.win + (int)(te.y == 31 ? 0 : te.y) % p.k) * p.pix_stride;
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const bool ok = (tmask[i] >> te.y) & 1u;
#pragma unroll
                        for (int e = 0; e < 8; ++e) vals[i][e] = 0.f;
                        if (ok) {
                            const size_t eo = (size_t)(pix_off[i] + te.x);
                            if (MODE == TC_IN_F32) {
                                const float4* src = reinterpret_cast<const float4*>(p.in_f32 + eo);
                                float4 a = __ldg(src), b = __ldg(src + 1);
                                vals[i][0] = a.x; vals[i][1] = a.y; vals[i][2] = a.z; vals[i][3] = a.w;
                                vals[i][4] = b.x; vals[i][5] = b.y; vals[i][6] = b.z; vals[i][7] = b.w;
                            } else {
                                uint4 qh = __ldg(reinterpret_cast<const uint4*>(p.in_hi + eo));
                                uint4 ql = __ldg(reinterpret_cast<const uint4*>(p.in_hi + p.in_plane + eo));
                                float fh[8], fl[8];
                                unpack8(qh, fh);
                                unpack8(ql, fl);
                                const float4* sp = reinterpret_cast<const float4*>(p.in_scale + (size_t)seg_idx[i] * p.cin + ci);
                                float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
                                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                                for (int e = 0; e < 8; ++e) vals[i][e] = (fh[e] + fl[e]) * sc[e];
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        uint4 hi, lo;
                        split8(vals[i], hi, lo);
                        *reinterpret_cast<uint4*>(st + dst0 + i * (RSTEP * 128)) = hi;
                        *reinterpret_cast<uint4*>(st + dst0 + i * (RSTEP * 128) + A_TILE_BYTES) = lo;
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&bar_full[s]);
                }
                if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
            }
        }
        if (MODE == TC_IN_PLANES) cp_async_wait_all();   // nothing may be in flight when the CTA exits
        if (prof && tid == 0) { p.prof[0] = pw0; p.prof[1] = (unsigned long long)(clock64() - prof_t0); }
    } else if (warp == PW) {
        // ================================ MMA issuer ================================
        // Accumulator a = TMEM columns [a*2NT, a*2NT + 2NT): [main = A_hi*W_hi | corr = A_hi*W_lo + A_lo*W_hi].
        // The tensor core truncates when it adds into the FP32 accumulator; keeping the 2^-11-sized
        // correction terms out of the big accumulator cuts the number of truncating adds 3x and the
        // epilogue adds main + corr once, round-to-nearest.  W_hi and W_lo tiles are adjacent in smem,
        // so A_hi * [W_hi | W_lo] is ONE MMA of N = 2NT.
        if (MODE == TC_IN_HALO) {
            // per K step (tap, 16 channels): A start-address delta and the B descriptor; tile-invariant
            const int ksteps = p.cin >> 4;
            const int nsteps = 9 * ksteps;                       // <= 18 (cin <= 32)
            const uint32_t lbo = HALO_PIX * 16u;                 // next 8-channel group of the same pixels
            for (int i = lane; i < nsteps; i += 32) {
                const int tap = i / ksteps, ks = i - tap * ksteps;
                const int ky = tap / 3, kx = tap - ky * 3;
                const int k = tap * p.cin + ks * 16;
                halo_a16[i] = ((uint32_t)(2 * ks) * lbo + (uint32_t)(ky * HALO_W + kx) * 16u) >> 4;   // row m <-> patch pixel m + ky*130 + kx
                if (p.debug_flags & 1) halo_a16[i] &= ~7u;
                halo_db[i] = umma_desc_sw128(smem_u32(tiles) + (uint32_t)(k >> 6) * w_bytes) + (uint64_t)(kDescKStep * ((k & 63) >> 4));
            }
            __syncwarp();
          if (elect_one()) {     // elect.sync: ptxas then knows a single lane runs the loop and moves descriptors with plain R2UR
            const uint32_t idesc2 = umma_idesc_f16(TM, 2 * NT), idesc1 = umma_idesc_f16(TM, NT);
            const uint64_t lo_delta = (uint64_t)(halo_plane_bytes >> 4);
            mbar_wait(&bar_w, 0);
            uint32_t s = 0, ph = 0, it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const uint32_t a = it & (n_acc - 1u);
                mbar_wait_t(&bar_acc_empty[a], ((it >> acc_shift) & 1u) ^ 1u, pw0, prof);
                mbar_wait_t(&bar_full[s], ph, pw1, prof);
                fence_proxy_async_smem();
                tc_fence_after();
                const uint32_t acc = tmem_base + a * 2u * (uint32_t)NT;
                const uint64_t da0 = umma_desc_nosw(smem_u32(halo_slots + (size_t)s * halo_slot_bytes), lbo, 128u);
#pragma unroll 6
                for (int i = 0; i < nsteps; ++i) {
                    const uint64_t da_hi = da0 + (uint64_t)halo_a16[i];
                    const uint64_t db = halo_db[i];
                    umma_f16(acc, da_hi, db, idesc2, i > 0 ? 1u : 0u);
                    umma_f16(acc + (uint32_t)NT, da_hi + lo_delta, db, idesc1, 1u);
                }
                umma_commit(&bar_empty[s]);
                umma_commit(&bar_acc_full[a]);
                if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
            }
            if (prof) { p.prof[2] = pw0; p.prof[3] = pw1; p.prof[4] = (unsigned long long)(clock64() - prof_t0); p.prof[7] = it; }
          }
        } else if (MODE != TC_IN_HALO && elect_one()) {
            const uint32_t idesc2 = umma_idesc_f16(TM, 2 * NT), idesc1 = umma_idesc_f16(TM, NT);
            const uint32_t a_kb = MODE == TC_IN_TMA ? (uint32_t)p.kb : 64u;
            // everything that does not depend on the tile is formed once: stage s / chunk kc / K step j only ADD to the
            // start-address field of a base descriptor (16-byte units; the whole ring lies below 256 KB)
            const uint32_t stage16 = stage_bytes >> 4, w16 = w_bytes >> 4;
            const uint64_t dA0 = MODE == TC_IN_TMA ? umma_desc_kmajor(smem_u32(tiles), a_kb * 2u) : umma_desc_sw128(smem_u32(tiles));
            const uint64_t dW0 = w_res ? umma_desc_sw128(smem_u32(tiles0)) : umma_desc_sw128(smem_u32(tiles) + 2 * A_TILE_BYTES);
            uint32_t dlt[4];                               // A start-address delta of K step j inside a chunk
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == TC_IN_TMA) {
                    // A: K block of kb channels per sub-tile (kb = 64 -> one SWIZZLE_128B tile per chunk)
                    const uint32_t kk = (uint32_t)j * 16u;
                    const uint32_t sub = kk / a_kb, in_sub = kk - sub * a_kb;
                    dlt[j] = ((sub * (uint32_t)TM * a_kb * 2u) >> 4) + (in_sub >> 3);
                } else {
                    dlt[j] = kDescKStep * (uint32_t)j;     // one SWIZZLE_128B [128][64] tile per plane
                }
            }
            if (w_res) mbar_wait(&bar_w, 0);
            uint32_t s = 0, ph = 0, it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const uint32_t a = it & (n_acc - 1u);
                mbar_wait_t(&bar_acc_empty[a], ((it >> acc_shift) & 1u) ^ 1u, pw0, prof);
                tc_fence_after();
                const uint32_t acc = tmem_base + a * 2u * (uint32_t)NT;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait_t(&bar_full[s], ph, pw1, prof);
                    fence_proxy_async_smem();          // cp.async / st.shared (generic proxy) -> tcgen05 reads (async proxy)
                    tc_fence_after();
                    const uint64_t da_s = dA0 + (uint64_t)(s * stage16);
                    const uint64_t db_hi = w_res ? dW0 + (uint64_t)((uint32_t)kc * w16) : dW0 + (uint64_t)(s * stage16);   // [W_hi | W_lo]: 2NT rows
                    int ksteps = (p.K - kc * KC + 15) >> 4;
                    if (ksteps > 4) ksteps = 4;
#pragma unroll 4
                    for (int j = 0; j < ksteps; ++j) {
                        const uint64_t da_hi = da_s + (uint64_t)dlt[j];
                        const uint64_t db = db_hi + (uint64_t)(kDescKStep * j);
                        umma_f16(acc, da_hi, db, idesc2, (kc | j) != 0 ? 1u : 0u);
                        if (MODE != TC_IN_TMA || !(p.debug_flags & 4))
                            umma_f16(acc + (uint32_t)NT, da_hi + (uint64_t)(A_TILE_BYTES >> 4), db, idesc1, 1u);
                    }
                    umma_commit(&bar_empty[s]);        // frees the smem stage when these MMAs retire
                    if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(&bar_acc_full[a]);         // accumulator complete -> epilogue
            }
            if (prof) { p.prof[2] = pw0; p.prof[3] = pw1; p.prof[4] = (unsigned long long)(clock64() - prof_t0); p.prof[7] = it; }
        }
        __syncwarp();
    } else {
        // ================================ epilogue ================================
        // Phase 1 (lane = output row = TMEM lane): main + corr + bias, activation, FP32 into this warp's
        // staging tile [32 rows][32 columns] (16-byte chunks XOR-swizzled by row & 7).
        // Phase 2 (lanes along the channel axis): + residual, hi/lo split, 16-byte stores that cover
        // whole 32/64-byte row pieces (full sectors) instead of one 16-byte piece in each of 32 lines.
        const int ew = warp - (PW + 1);
        const bool epi_leader = elect_one();           // the lane that issues this warp's TMA stores (and waits for them)
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        // tile_split: half of the warp sets serve accumulator 0, the other half accumulator 1
        const int nset_all = epi_warps >> 2;
        const int sets_per_acc = tile_split ? nset_all >> acc_shift : nset_all;
        const int wset = tile_split ? (ew >> 2) / sets_per_acc : 0;
        const int cset = tile_split ? (ew >> 2) % sets_per_acc : (ew >> 2);
        const int nsets = sets_per_acc;
        float* stg = reinterpret_cast<float*>(tiles + ring_bytes + (size_t)ew * EPI_STAGE_BYTES);
        uint32_t it = 0;
        const bool vec = (p.cout & 15) == 0;           // every 16-column group is full and 16-byte aligned
        const bool staged = vec && (p.out_f32 == nullptr || p.out_f32_rows) && p.spec_nframes == 0;   // measured: per-lane direct FP32 row stores are 10-35% slower than the staged, coalesced ones
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int mt = p.n_tiles == 1 ? t : t / p.n_tiles;
            const int n_tile = p.n_tiles == 1 ? 0 : t - mt * p.n_tiles;
            const int sb = mt / p.tiles_per_seg, lt = mt - sb * p.tiles_per_seg;
            const int lp0 = lt * TM + q * 32;                  // first row of this warp inside the segment
            const int lp = lp0 + lane;
            const uint32_t a = it & (n_acc - 1u);
            if (tile_split && (int)a != wset) continue;
            mbar_wait_t(&bar_acc_full[a], (it >> acc_shift) & 1u, pw0, prof && warp == PW + 1);
            tc_fence_after();
            const int row = sb * p.pix_per_seg + lp;
            const bool row_ok = lp < p.pix_per_seg && row < p.M;
            const uint32_t t_lane = tmem_base + a * 2u * (uint32_t)NT + ((uint32_t)(q * 32) << 16);
            if (p.debug_flags & 2) {
            } else if (staged) {
                const bool tma_out = p.out_tma != 0 && p.tiles_per_seg == p.m_tiles;
                const bool f32_out = p.out_f32 != nullptr;
                for (int c0 = cset * 32; c0 < NT; c0 += 32 * nsets) {
                    const int gw = NT - c0 < 32 ? 16 : 32;         // NT is a multiple of 16
                    if (tma_out && gw != 32) {                      // the old path reuses the tile a TMA store may still be reading
                        if (epi_leader) bulk_wait_group_read0();
                        __syncwarp();
                    }
                    for (int h = 0; h < gw; h += 16) {
                        uint32_t rm[16], rc[16];
                        __syncwarp();                              // tcgen05.ld is .sync.aligned: reconverge first
                        tmem_ld16_nowait(t_lane + (uint32_t)(c0 + h), rm);
                        tmem_ld16_nowait(t_lane + (uint32_t)(NT + c0 + h), rc);
                        tmem_ld_wait();
                        const int n = n_tile * NT + c0 + h;
                        float v[16];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 bv = s_bias_ok ? *reinterpret_cast<const float4*>(s_bias + n + 4 * j4)
                                                        : __ldg(reinterpret_cast<const float4*>(p.bias + n) + j4);
                            v[4 * j4 + 0] = __uint_as_float(rm[4 * j4 + 0]) + __uint_as_float(rc[4 * j4 + 0]) + bv.x;
                            v[4 * j4 + 1] = __uint_as_float(rm[4 * j4 + 1]) + __uint_as_float(rc[4 * j4 + 1]) + bv.y;
                            v[4 * j4 + 2] = __uint_as_float(rm[4 * j4 + 2]) + __uint_as_float(rc[4 * j4 + 2]) + bv.z;
                            v[4 * j4 + 3] = __uint_as_float(rm[4 * j4 + 3]) + __uint_as_float(rc[4 * j4 + 3]) + bv.w;
                        }
                        if (p.act == KACT_SILU) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = BN_EPI_FMA_RCP ? silu_fma_rcp(v[j]) : silu_approx(v[j]);
                        } else if (p.act == KACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = sigmoid_approx(v[j]);
                        }
                        if (tma_out && gw == 32 && p.res_hi != nullptr && row_ok) {
                            // residual: this lane's row, 16 channels = one 32-byte sector per plane
                            const size_t o = (size_t)row * p.cout + n;
                            const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + o);
                            const uint4* rl = reinterpret_cast<const uint4*>(p.res_hi + p.res_plane + o);
#pragma unroll
                            for (int j2 = 0; j2 < 2; ++j2) {
                                const uint4 qh = __ldg(rh + j2), ql = __ldg(rl + j2);
                                float fh[8], fl[8];
                                unpack8(qh, fh);
                                unpack8(ql, fl);
#pragma unroll
                                for (int e8 = 0; e8 < 8; ++e8) v[8 * j2 + e8] += fh[e8] + fl[e8];
                            }
                        }
                        if (tma_out && gw == 32 && h == 0) {
                            // the TMA engine may still be reading this warp's tile (previous block): wait right before the
                            // first shared-memory store, after this block's TMEM loads and math are already done
                            if (epi_leader) bulk_wait_group_read0();
                            __syncwarp();
                        }
                        if (tma_out && gw == 32 && !f32_out) {
                            // planes: split here (lane = row), fp16 tiles [hi|lo][32 rows][64 B], SWIZZLE_64B chunk order
                            uint4 hq[2], lq[2];
                            split8(v, hq[0], lq[0]);
                            split8(v + 8, hq[1], lq[1]);
                            uint8_t* t8 = reinterpret_cast<uint8_t*>(stg);
                            const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
                            for (int j2 = 0; j2 < 2; ++j2) {
                                const uint32_t cc = (uint32_t)((h >> 3) + j2) ^ sw;
                                *reinterpret_cast<uint4*>(t8 + lane * 64 + (cc << 4)) = hq[j2];
                                *reinterpret_cast<uint4*>(t8 + 2048 + lane * 64 + (cc << 4)) = lq[j2];
                            }
                        } else {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const int cc = (h >> 2) + j4;          // 16-byte chunk of the 32-column row
                            *reinterpret_cast<float4*>(stg + lane * 32 + ((cc ^ (lane & 7)) << 2)) =
                                make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                        }
                        }
                    }
                    if (tma_out && gw == 32) {
                        fence_proxy_async_smem();                  // generic-proxy tile writes -> visible to the TMA engine
                        __syncwarp();
                        if (epi_leader) {
                            const int row0 = sb * p.pix_per_seg + lp0;
                            if (f32_out) tma_store_2d(&p.omap, stg, n_tile * NT + c0, row0);
                            else tma_store_3d(&p.omap, stg, n_tile * NT + c0, row0, 0);
                            bulk_commit_group();
                        }
                        continue;                                  // rows past M are clipped by the tensor map
                    }
                    __syncwarp();
                    const int osh = gw == 32 ? 2 : 1;              // log2(8-column octets per row)
                    const int opr = 1 << osh;
                    const int rpi = 32 >> osh;                     // rows per iteration
                    const int oc = lane & (opr - 1);
#pragma unroll 2
                    for (int i = 0; i < opr; ++i) {
                        const int r = i * rpi + (lane >> osh);
                        const float4 x0 = *reinterpret_cast<const float4*>(stg + r * 32 + (((2 * oc) ^ (r & 7)) << 2));
                        const float4 x1 = *reinterpret_cast<const float4*>(stg + r * 32 + (((2 * oc + 1) ^ (r & 7)) << 2));
                        const int lpr = lp0 + r;
                        const int rr = sb * p.pix_per_seg + lpr;
                        if (lpr < p.pix_per_seg && rr < p.M) {
                            float w8[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                            const size_t o = (size_t)rr * p.cout + n_tile * NT + c0 + oc * 8;
                            if (p.res_hi) {
                                const uint4 rh = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o));
                                const uint4 rl = __ldg(reinterpret_cast<const uint4*>(p.res_hi + p.res_plane + o));
                                float fh[8], fl[8];
                                unpack8(rh, fh);
                                unpack8(rl, fl);
#pragma unroll
                                for (int e = 0; e < 8; ++e) w8[e] += fh[e] + fl[e];
                            }
                            if (p.out_f32) {
                                *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(w8[0], w8[1], w8[2], w8[3]);
                                *reinterpret_cast<float4*>(p.out_f32 + o + 4) = make_float4(w8[4], w8[5], w8[6], w8[7]);
                            } else {
                                uint4 hi, lo;
                                split8(w8, hi, lo);
                                *reinterpret_cast<uint4*>(p.out_hi + o) = hi;
                                *reinterpret_cast<uint4*>(p.out_hi + p.out_plane + o) = lo;
                            }
                        }
                    }
                    __syncwarp();                                  // staging tile free for the next group
                }
            } else {
                size_t spec_base = 0;
                if (p.spec_nframes > 0 && row_ok) {
                    const int b = row / p.spec_nframes, tt = row - b * p.spec_nframes;
                    spec_base = ((size_t)b * p.cout * p.spec_nframes + tt) * p.spec_nch + p.spec_ch;
                }
                for (int c0 = cset * 16; c0 < NT; c0 += 16 * nsets) {
                    uint32_t rm[16], rc[16];
                    float v[16];
                    __syncwarp();
                    tmem_ld16_nowait(t_lane + (uint32_t)c0, rm);
                    tmem_ld16_nowait(t_lane + (uint32_t)(NT + c0), rc);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
                    const int n = n_tile * NT + c0;
                    if (!row_ok || n >= p.cout) continue;
                    if (p.spec_nframes > 0) {
                        // spectrogram: square, power-compress, scatter (lanes = consecutive frames -> coalesced)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (n + j >= p.cout) break;
                            const float pw = v[j] * v[j];
                            const float r = pw > 0.f ? exp2f(p.spec_exponent * __log2f(pw)) : pw;
                            const size_t o = spec_base + (size_t)(n + j) * p.spec_nframes * p.spec_nch;
                            const __half hh = __float2half_rn(r);
                            p.out_hi[o] = hh;
                            p.out_hi[p.out_plane + o] = __float2half_rn(r - __half2float(hh));
                        }
                        continue;
                    }
                    const size_t o = (size_t)row * p.cout + n;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n + j >= p.cout) break;
                        float r = act_fn(v[j] + __ldg(p.bias + n + j), p.act);
                        if (p.res_hi) r += __half2float(p.res_hi[o + j]) + __half2float(p.res_hi[p.res_plane + o + j]);
                        if (p.out_f32) {
                            p.out_f32[o + j] = r;
                        } else {
                            __half hh = __float2half_rn(r);
                            p.out_hi[o + j] = hh;
                            p.out_hi[p.out_plane + o + j] = __float2half_rn(r - __half2float(hh));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_empty[a]);
        }
        if (p.out_tma != 0 && epi_leader) bulk_wait_group0();     // shared memory must outlive the engine's reads
        if (prof && warp == PW + 1 && lane == 0) { p.prof[5] = pw0; p.prof[6] = (unsigned long long)(clock64() - prof_t0); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PW) tmem_dealloc(tmem_base, p.tmem_cols);
}

constexpr size_t SMEM_TWO_PER_SM = 108 * 1024;     // dynamic bytes that still let two CTAs share an SM
constexpr size_t SMEM_ONE_PER_SM = 220 * 1024;

size_t tc_conv_smem_bytes(int nt, int stages, int epi_warps) {
    return (size_t)stages * (2 * A_TILE_BYTES + 2 * (size_t)nt * 128) + 1024 + (size_t)epi_warps * EPI_STAGE_BYTES;
}

unsigned long long* tc_conv_prof_slot(int slot) {
    if (slot < 0 || slot >= 128) return nullptr;
    void* base = nullptr;
    if (cudaGetSymbolAddress(&base, g_tc_prof) != cudaSuccess) return nullptr;
    return reinterpret_cast<unsigned long long*>(base) + (size_t)slot * 16;
}
cudaError_t tc_conv_prof_read(unsigned long long* out, int slots) {
    return cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(unsigned long long) * 16 * (size_t)(slots < 128 ? slots : 128));
}

static int resolve_epi_warps(int epi_warps) { return epi_warps == 8 || epi_warps == 16 ? epi_warps : g_epi_warps_one; }

size_t tc_conv_halo_smem_bytes(int cin, int nt, int k_chunks, int slots, int epi_warps) {
    const size_t plane = (size_t)(cin / 8) * HALO_PIX * 16;
    const size_t slot = (2 * plane + 1023) & ~(size_t)1023;
    return (size_t)k_chunks * 2 * nt * 128 + (size_t)slots * slot + 1024 + (size_t)resolve_epi_warps(epi_warps) * EPI_STAGE_BYTES;
}

int tc_conv_halo_slots(int k, int stride, int pad, int cin, int wout, int win, int nt, int k_chunks, int epi_warps) {
    if (k != 3 || stride != 1 || pad != 1 || (cin != 16 && cin != 32) || wout != win || (wout % TM) != 0) return 0;
    for (int s = MAX_HALO_SLOTS; s >= 2; --s)
        if (tc_conv_halo_smem_bytes(cin, nt, k_chunks, s, epi_warps) <= SMEM_ONE_PER_SM) return s;
    return 0;
}

// two CTAs per SM (4 epilogue warps each, <= 256 TMEM columns each) when at least two stages fit,
// else one CTA per SM with 8 epilogue warps and the deepest ring that fits
static bool two_per_sm(int nt, int stages) { return 4 * nt <= 256 && tc_conv_smem_bytes(nt, stages, 4) <= SMEM_TWO_PER_SM; }

int tc_conv_pick_stages(int nt, int k_chunks, int epi_warps) {
    (void)k_chunks;                     // the ring runs across tiles, so depth is useful even for K <= 64
    if (two_per_sm(nt, 2)) {
        int s = MAX_STAGES;
        while (s > 2 && !two_per_sm(nt, s)) --s;
        return s;
    }
    int s = MAX_STAGES;
    while (s > 2 && tc_conv_smem_bytes(nt, s, resolve_epi_warps(epi_warps)) > SMEM_ONE_PER_SM) --s;
    return s;
}

cudaError_t tc_conv_init_device() {
    { const char* ev = getenv("BN_EPI_WARPS"); if (ev && atoi(ev) == 8) g_epi_warps_one = 8; }
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv<TC_IN_PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ONE_PER_SM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tc_conv<TC_IN_PLANES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TWO_PER_SM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tc_conv<TC_IN_PLANES, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ONE_PER_SM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tc_conv<TC_IN_PLANES_SCALED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ONE_PER_SM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tc_conv<TC_IN_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ONE_PER_SM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tc_conv<TC_IN_HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ONE_PER_SM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_tc_conv<TC_IN_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ONE_PER_SM);
}

cudaError_t launch_tc_conv(const TcConvParams& pin, int num_sms, cudaStream_t stream) {
    if (pin.M <= 0) return cudaSuccess;
    TcConvParams p = pin;
    { static int dbg = -1; if (dbg < 0) { const char* ev = getenv("BN_TC_DEBUG"); dbg = ev ? atoi(ev) : 0; } p.debug_flags = dbg; }
    if (p.in_mode != TC_IN_TMA || p.tiles_per_seg <= 0) {      // cp.async modes: one flat "segment"
        if (p.in_mode != TC_IN_TMA) { p.tiles_per_seg = p.m_tiles; p.pix_per_seg = p.M; }
    }
    if (p.in_mode == TC_IN_TMA && (p.kb != 16 && p.kb != 32 && p.kb != 64)) return cudaErrorInvalidValue;
    if (p.k_chunks > MAX_K_CHUNKS || p.k * p.k > 31) return cudaErrorInvalidValue;
    if ((p.cin & 7) || p.nt < 16 || p.nt > 128 || (p.nt & 15) || p.stages < 2 || p.stages > MAX_STAGES ||
        p.tmem_cols < 4 * p.nt || p.tmem_cols > 512)
        return cudaErrorInvalidValue;
    const int total = p.m_tiles * p.n_tiles;
    if (p.in_mode == TC_IN_HALO) {
        // p.stages = patch slots (tc_conv_halo_slots); every CTA keeps one n tile's weights resident
        if (p.stages < 2 || p.stages > MAX_HALO_SLOTS || p.k != 3 || p.stride != 1 || p.pad != 1 || (p.cin != 16 && p.cin != 32) ||
            (p.wout % TM) != 0 || p.win != p.wout || p.K != 9 * p.cin)
            return cudaErrorInvalidValue;
        const size_t hs = tc_conv_halo_smem_bytes(p.cin, p.nt, p.k_chunks, p.stages, p.epi_warps);
        if (hs > SMEM_ONE_PER_SM) return cudaErrorInvalidValue;
        int g = total < num_sms ? total : num_sms;
        g -= g % p.n_tiles;
        if (g <= 0) return cudaErrorInvalidValue;
        k_tc_conv<TC_IN_HALO><<<g, (unsigned)(5 + resolve_epi_warps(p.epi_warps)) * 32u, hs, stream>>>(p);
        return cudaGetLastError();
    }
    // two co-resident CTAs per SM when shared memory, TMEM (512 columns) and registers allow it
    const int per_sm = (p.in_mode == TC_IN_PLANES && two_per_sm(p.nt, p.stages)) ? 2 : 1;
    const int epi_warps = per_sm == 2 ? 4 : resolve_epi_warps(p.epi_warps);
    size_t smem = tc_conv_smem_bytes(p.nt, p.stages, epi_warps);
    if (p.spec_nframes > 0 && per_sm == 1 && p.in_mode == TC_IN_PLANES) {
        // the spectrogram epilogue stores straight from registers: no staging tiles, so the whole budget goes to the
        // operand ring (this GEMM streams ~2 MB per tile through it and is bound by the bytes in flight per SM)
        const size_t stage_b = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)p.nt * 128;
        int st = (int)((SMEM_ONE_PER_SM - 1024) / stage_b);
        if (st > MAX_STAGES) st = MAX_STAGES;
        if (st > p.stages) { p.stages = st; smem = 1024 + (size_t)st * stage_b; }
    }
    if (smem > SMEM_ONE_PER_SM) return cudaErrorInvalidValue;
    const int slots = num_sms * per_sm;
    dim3 grid((unsigned)(total < slots ? total : slots));
    const unsigned nthr = (unsigned)(5 + epi_warps) * 32u;
    if (p.in_mode == TC_IN_TMA) {
        // weights resident when this CTA's whole [K x NT] tile plus a >= 3 deep A ring fit: halves the L2 -> SM traffic of
        // shallow-K layers (expand 1x1 convs).  Measured: -8% with the epilogue switched off, ~1% with it on - these
        // layers are bound by the epilogue warps (tools/tc_role_profile.py), not by operand traffic or the tensor pipe.
        static const bool no_wres = [] { const char* ev = getenv("BN_DISABLE_WRES"); return ev && ev[0] == '1'; }();
        const size_t wres = (size_t)p.k_chunks * 2 * p.nt * 128;
        const size_t fixed = 1024 + (size_t)epi_warps * EPI_STAGE_BYTES + wres;
        int g = (int)grid.x - (int)grid.x % p.n_tiles;
        if (!no_wres && fixed + 3 * (size_t)(2 * A_TILE_BYTES) <= SMEM_ONE_PER_SM && g >= p.n_tiles && p.m_tiles >= 2 * (g / p.n_tiles)) {
            int st = (int)((SMEM_ONE_PER_SM - fixed) / (2 * A_TILE_BYTES));
            if (st > MAX_STAGES) st = MAX_STAGES;
            p.w_resident = 1;
            p.stages = st;
            const size_t smem_res = fixed + (size_t)st * 2 * A_TILE_BYTES;
            k_tc_conv<TC_IN_TMA><<<g, nthr, smem_res, stream>>>(p);
            return cudaGetLastError();
        }
    }
    switch (p.in_mode) {
        case TC_IN_PLANES: {
            static const bool pw8 = [] { const char* ev = getenv("BN_TC_PW"); return !(ev && atoi(ev) == 4); }();
            if (per_sm == 2) k_tc_conv<TC_IN_PLANES, true><<<grid, nthr, smem, stream>>>(p);
            else if (pw8) k_tc_conv<TC_IN_PLANES, false, 8><<<grid, nthr + 128u, smem, stream>>>(p);
            else k_tc_conv<TC_IN_PLANES><<<grid, nthr, smem, stream>>>(p);
            break;
        }
        case TC_IN_PLANES_SCALED: k_tc_conv<TC_IN_PLANES_SCALED><<<grid, nthr, smem, stream>>>(p); break;
        case TC_IN_TMA: k_tc_conv<TC_IN_TMA><<<grid, nthr, smem, stream>>>(p); break;
        default: k_tc_conv<TC_IN_F32><<<grid, nthr, smem, stream>>>(p); break;
    }
    return cudaGetLastError();
}

bool tc_encode_tmap(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                    const uint32_t box[5], const uint32_t elem_strides[5], int kb) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr);
        if (ge != cudaSuccess || !sym) {
            if (getenv("BN_DEBUG")) fprintf(stderr, "[bn] cuTensorMapEncodeTiled entry point unavailable: %s (qr=%d)\n", cudaGetErrorString(ge), (int)qr);
            return false;
        }
        fn = reinterpret_cast<EncodeFn>(sym);
    }
    const CUtensorMapSwizzle sw = kb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < 5; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides[i]; }
    for (int i = 0; i < 4; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("BN_DEBUG"))
        fprintf(stderr, "[bn] cuTensorMapEncodeTiled failed: CUresult %d (dims %llu %llu %llu %llu %llu, strides %llu %llu %llu %llu)\n", (int)r,
                (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2], (unsigned long long)gd[3], (unsigned long long)gd[4],
                (unsigned long long)gs[0], (unsigned long long)gs[1], (unsigned long long)gs[2], (unsigned long long)gs[3]);
    return r == CUDA_SUCCESS;
}

bool tc_encode_out_tmap(CUtensorMap* out, void* base, uint64_t rows, uint64_t cout, bool f32, uint64_t plane_elems) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) != cudaSuccess || !sym) return false;
        fn = reinterpret_cast<EncodeFn>(sym);
    }
    if (cout < 32 || rows == 0) return false;
    CUresult r;
    if (f32) {
        const cuuint64_t gd[2] = {cout, rows};
        const cuuint64_t gs[1] = {cout * 4};
        const cuuint32_t bx[2] = {32, 32}, es[2] = {1, 1};
        r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gd[3] = {cout, rows, 2};
        const cuuint64_t gs[2] = {cout * 2, plane_elems * 2};
        const cuuint32_t bx[3] = {32, 32, 2}, es[3] = {1, 1, 1};
        r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS && getenv("BN_DEBUG")) fprintf(stderr, "[bn] output tensor map refused: CUresult %d (rows %llu cout %llu f32 %d)\n", (int)r,
                                                         (unsigned long long)rows, (unsigned long long)cout, (int)f32);
    return r == CUDA_SUCCESS;
}

// Host: split + swizzle the [K][ldw] FP32 weight matrix into the per-(n_tile, k_chunk) smem images.
void tc_pack_weights(const float* w, int K, int cout, int ldw, int nt, std::vector<uint16_t>& out,
                     int* n_tiles_out, int* k_chunks_out) {
    const int n_tiles = (cout + nt - 1) / nt;
    const int k_chunks = (K + KC - 1) / KC;
    const size_t img = (size_t)nt * 64;                        // halfs per plane image
    out.assign((size_t)n_tiles * k_chunks * 2 * img, 0);
    for (int t = 0; t < n_tiles; ++t)
        for (int kc = 0; kc < k_chunks; ++kc) {
            uint16_t* hi = out.data() + ((size_t)t * k_chunks + kc) * 2 * img;
            uint16_t* lo = hi + img;
            for (int nn = 0; nn < nt; ++nn) {
                const int co = t * nt + nn;
                if (co >= cout) continue;
                for (int u = 0; u < 8; ++u)
                    for (int e = 0; e < 8; ++e) {
                        const int k = kc * KC + u * 8 + e;
                        if (k >= K) continue;
                        const float x = w[(size_t)k * ldw + co];
                        const __half h = __float2half_rn(x);
                        const __half l = __float2half_rn(x - __half2float(h));
                        const size_t idx = (sw128_offset((uint32_t)nn, (uint32_t)u) >> 1) + e;
                        hi[idx] = *reinterpret_cast<const uint16_t*>(&h);
                        lo[idx] = *reinterpret_cast<const uint16_t*>(&l);
                    }
            }
        }
    *n_tiles_out = n_tiles;
    *k_chunks_out = k_chunks;
}

}  // namespace bn
