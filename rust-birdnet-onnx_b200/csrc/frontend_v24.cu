// BirdNET v2.4-style audio front-end, ONE kernel for both mel-spectrogram branches (the reference has this
// arithmetic inside the model file it feeds to ONNX Runtime, src/classifier.rs:851-853 — there is no reference source for it).
//
//   spec[b][mel][t][branch] = pow( ( sum_n  xn[b][t*hop + n] * basis[n][mel] )^2 , exponent )
//   xn = ((x - min) / (max - min + eps) - half) * two            (per segment; min / max come from k_minmax_partial)
//
// What the kernel moves: the raw FP32 audio (once from HBM: the second branch reads the same range out of L2) and the
// spectrogram out.  No normalised copy and no frame matrix ever exists in HBM.  One persistent CTA per SM, 20 warps:
//
//   producer warps   read the raw samples of one 128-frame tile, normalise them (the stand-alone normaliser's formula; the
//                    division is a reciprocal + one residual correction), split to fp16 hi / lo and store them as a
//                    SAMPLE PATCH in shared memory: patch row p = the `hop` samples starting at (t0 + p) * hop, kept as
//                    16-byte cells [cell][row].  Frame t0 + i is then rows i, i+1, ... read left to right (block-Toeplitz):
//                    the K range of hop-block j is the same cells shifted j rows down, which for a no-swizzle K-major UMMA
//                    operand is a start-address shift of j * 16 bytes.  135 rows x 35 cells x 2 planes = 152 KB instead of
//                    128 x 2240 x 2 x 2 = 1.1 MB of frame matrix.
//   loader lane      streams the packed basis [K step][W_hi | W_lo] through a ring of 1-D bulk copies (12 KB per slot).
//   control lane     issues, per 16-sample K step, hi * [W_hi | W_lo] (N = 2 * mels: main | correction) and lo * W_hi
//                    (N = mels, into the correction half) into one of two TMEM accumulator sets; table-driven, 32-bit
//                    shared addresses formed once: it has ~144 cycles per K step.
//   epilogue warps   TMEM -> square -> power-compress -> hi / lo planes of the [mel][frame][branch] image the stem reads,
//                    overlapped with the MMAs of the next tile (second accumulator set).  The two branches of the same 128
//                    frames run back to back: the first one's results wait in spare TMEM columns (tcgen05.st), the second
//                    one's epilogue writes both channels of a pixel as one 32-bit word per plane (full sectors).
//
// The K loop runs cell-column pair by cell-column pair (all hop-blocks of a pair before the next pair), so a column pair of
// the patch is dead as soon as its MMAs have completed: the producers refill it for the NEXT tile while the MMAs of this
// tile are still running on the later columns.  The patch is its own double buffer.
// Measurements, what bounds the kernel and what was tried: DESIGN.md section 6 (`k_spec_v24` in detail).
#include "frontend_v24.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace bn {
using namespace tc;

namespace {

constexpr int SV_EPI_WARPS = 8;                     // two per TMEM lane quarter (alternate 16-column chunks)
constexpr int SV_PROD_WARPS = 10;
constexpr int SV_WARPS = 2 + SV_EPI_WARPS + SV_PROD_WARPS;      // control, loader, epilogue, producers (9 or a third loader warp: no gain)
constexpr int SV_THREADS = SV_WARPS * 32;
constexpr int SV_MAX_GROUPS = 24;                   // cell-column pairs of the patch
constexpr int SV_MAX_STAGES = 12;
constexpr uint32_t SV_FIRST = 1u << 24, SV_LAST = 1u << 25;   // K-step table flags: first / last step of a column pair
constexpr uint32_t SV_TRACE_SLOTS = 670;             // 3 words each: fits the 127 x 16-word scratch in front of the profile slot
constexpr uint32_t SV_SMEM_MAX = 224 * 1024;        // + the static barriers stays under the 227 KB block limit

__device__ __forceinline__ float sv_key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// the single-lane roles address their barriers by 32-bit shared address (formed once, outside the loops)
__device__ __forceinline__ void sv_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (++spins > (1u << 28)) __trap();
    } while (!ok);
}
template <bool PROF>
__device__ __forceinline__ void sv_wait1(uint32_t bar, uint32_t parity, unsigned long long& acc) {
    if (!PROF) { sv_wait_u32(bar, parity); return; }
    const long long t0 = clock64();
    sv_wait_u32(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}
// registers -> TMEM: lane i of the warp -> TMEM lane base+i, 16 consecutive 32-bit columns (the mirror of tmem_ld16)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// out of line on purpose: the producers call it only for degenerate denominators and must not pay for it otherwise
__device__ __noinline__ float sv_ieee_div(float a, float b) { return __fdiv_rn(a, b); }
// keeps an address the compiler would otherwise re-derive (S2UR SR_CgaCtaId + ULEA per use) in a register
__device__ __forceinline__ uint32_t sv_opaque(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void sv_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// cluster forms: the commit arrives on the barrier at the same offset in every CTA of the mask, the bulk copy lands in
// every CTA of the mask (same offsets) and completes bytes on each one's barrier
__device__ __forceinline__ void sv_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void sv_bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void sv_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t sv_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void sv_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sv_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <bool PROF>
__device__ __forceinline__ void sv_wait(uint64_t* bar, uint32_t parity, unsigned long long& acc) {
    if (!PROF) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}

extern __shared__ __align__(16) uint8_t sv_smem_raw[];

// tile n of this CTA -> (branch slot, segment, first frame).
// Paired order (two branches whose results fit the TMEM stash): the CTA runs both branches of the same 128 frames back
// to back - slot 0, then slot 1 - so the second read of the audio range hits L2 and the epilogue can write both channels
// of a spectrogram pixel as one 32-bit word (full 32-byte sectors instead of every other 2 bytes).
// Unpaired order: slot 0 (the longer K loop) is dealt round-robin from CTA 0 up, slot 1 from the last CTA down, so a CTA
// that got one more long tile gets one fewer short tile.
struct SvTile { int slot, seg, t0; };
__device__ __forceinline__ bool sv_paired(const SpecV24Params& p) { return p.n_br == 2 && 5 * p.n_pad <= 512 && p.n_ch == 2; }
__device__ __forceinline__ bool sv_tile(const SpecV24Params& p, int n, SvTile& t) {
    const int grid = (int)gridDim.x, bx = (int)blockIdx.x;
    const int per = p.batch * p.tiles_per_seg;
    int idx;
    if (sv_paired(p)) {
        t.slot = n & 1;
        idx = bx + (n >> 1) * grid;
        if (idx >= per) return false;
    } else {
        const int mine0 = bx < per ? (per - 1 - bx) / grid + 1 : 0;
        if (n < mine0) {
            t.slot = 0;
            idx = bx + n * grid;
        } else {
            if (p.n_br < 2) return false;
            t.slot = 1;
            idx = (grid - 1 - bx) + (n - mine0) * grid;
            if (idx >= per) return false;
        }
    }
    t.seg = idx / p.tiles_per_seg;
    t.t0 = (idx - t.seg * p.tiles_per_seg) * 128;
    return true;
}

// K steps of cell-column pair m: all blocks that still have weighted cells there, + the zero-weight step that makes the
// total even... (the packed basis has the same order: spec_v24_plan)
__device__ __forceinline__ int sv_group_steps(const SpecBranchDev& br, int m, int groups) {
    return (m < br.split ? br.blocks : br.blocks - 1) + (m == groups - 1 ? br.pad : 0);
}

// Order of the column pairs: the same for every CTA (a per-CTA rotation was tried to spread the basis reads over L2 — no
// effect on time, and it would make a segment's bits depend on its position in the batch).
__device__ __forceinline__ int sv_rot(const SpecV24Params&, int) { return 0; }

// CL (opt-in, BN_FE_CLUSTER=1): launched as clusters of two CTAs that share the basis stream - each loader fetches every
// other ring slot and multicasts it into both CTAs' rings (half the bulk-copy bytes per SM, half the L2 -> SM traffic); a
// slot is free again when BOTH control lanes' MMAs on it have completed (their commits arrive on both CTAs' barriers).
// Both CTAs of a pair always have the same number of tiles (tiles per segment is even), so they walk identical slot
// sequences.  Measured: bit-identical results and the same 0.31 ms as the plain launch - the control lane's wait for basis
// slots is not a bandwidth limit of the bulk-copy engine, so the plain launch stays the default.
template <bool PROF, bool CL>
__global__ void __launch_bounds__(SV_THREADS, 1) k_spec_v24(const SpecV24Params p) {
    __shared__ __align__(8) uint64_t g_full[SV_MAX_GROUPS];     // patch column pair written (producer warp -> control)
    __shared__ __align__(8) uint64_t g_empty[SV_MAX_GROUPS];    // its MMAs completed (tcgen05.commit -> producers)
    __shared__ __align__(8) uint64_t w_full[SV_MAX_STAGES];     // basis ring slot (two K steps) landed (bulk copy -> control)
    __shared__ __align__(8) uint64_t w_empty[SV_MAX_STAGES];    // its MMAs completed (tcgen05.commit -> loader)
    __shared__ __align__(8) uint64_t acc_full[2];               // all MMAs of a tile completed (-> epilogue)
    __shared__ __align__(8) uint64_t acc_empty[2];              // accumulator set read out (epilogue warps -> control)
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(8) uint32_t s_tab[2][SV_MAX_KSTEPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sm = sv_smem_raw + ((128u - (smem_u32(sv_smem_raw) & 127u)) & 127u);
    const uint32_t plane = p.patch_plane;                       // bytes of one patch plane (hi; lo follows)
    uint8_t* patch = sm;
    uint8_t* wring = sm + 2u * plane;
    const uint32_t N = (uint32_t)p.n_pad;                       // mel columns, multiple of 16
    const uint32_t kstep_bytes = 64u * N;                       // [2 cells][2N rows][16 B]
    const uint32_t NS = (uint32_t)p.n_stages;

    if (tid == 0) {
        for (int i = 0; i < SV_MAX_GROUPS; ++i) { mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 1); }
        for (int i = 0; i < SV_MAX_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], CL ? 2 : 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], SV_EPI_WARPS); }
        fence_barrier_init();
    }
    // K-step table of this CTA (its own rotation of the column pairs): patch offset in 16-byte cells | pair << 16 | flags
    if (tid < p.n_br) {
        const SpecBranchDev& br = p.br[tid];
        const int groups = (br.kcells + 1) >> 1;
        const int rot = sv_rot(p, groups);
        int ks = 0;
        for (int mi = 0; mi < groups; ++mi) {
            const int m = mi + rot < groups ? mi + rot : mi + rot - groups;
            const int jn = sv_group_steps(br, m, groups);
            for (int j = 0; j < jn; ++j, ++ks)
                s_tab[tid][ks] = ((uint32_t)m * 2u * (uint32_t)p.row_pitch + (uint32_t)j) | ((uint32_t)m << 16) |
                                 (j == 0 ? SV_FIRST : 0u) | (j == jn - 1 ? SV_LAST : 0u);
        }
    }
    // the whole patch starts as zeros: rows and cells no producer writes are still read (with zero weights, or for frames
    // past the end that are never stored), so they must hold finite numbers
    for (uint32_t i = tid; i < 2u * plane / 16u; i += SV_THREADS) reinterpret_cast<uint4*>(patch)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(&tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL) sv_cluster_sync();                                  // the peer's barriers exist before anything arrives on them
    const uint32_t tmem_base = tmem_holder;
    // development aid (PROF): cycles of CTA 0.  control [0] wait accumulator free [1] wait patch columns [2] wait basis step
    // [3] total [4] tiles | producer warp 0: [5] wait columns free [6] total | epilogue warp 0: [7] wait MMAs [8] total |
    // loader: [9] wait ring slot [10] total
    const bool prof = PROF && p.prof != nullptr && blockIdx.x == 0;
    // ring trace (BN_FE_DEBUG bit 1024, CTA 0): per basis slot the cycle the loader issued its copy, the cycle the control
    // lane saw it landed and the cycle it committed the slot's MMAs -> the 127 x 16 words below the counters (tools/fe_ring_trace.py)
    unsigned long long* trace = (prof && (p.debug & 1024)) ? p.prof - 127 * 16 : nullptr;
    unsigned long long pc0 = 0, pc1 = 0, pc2 = 0;
    const long long t_begin = PROF ? clock64() : 0;

    if (warp == 0) {
        // ================================ control: every tcgen05.mma ================================
        // One lane runs this loop and nothing else; it has to stay under the ~144 cycles the two MMAs of a K step take, so
        // everything is plain 32-bit shared addresses and counters: one wait + one commit per ring slot (two K steps), one
        // wait + one commit per column pair.
        if (elect_one()) {
            const uint32_t idesc2 = umma_idesc_f16(128, (int)(2u * N)), idesc1 = umma_idesc_f16(128, (int)N);
            // descriptors without the start-address field; the field is (shared address >> 4), all addresses < 256 KB
            const uint64_t da_hi0 = umma_desc_nosw(smem_u32(patch), (uint32_t)p.row_pitch * 16u, 128u);
            const uint64_t da_lo0 = da_hi0 + (uint64_t)(plane >> 4);
            const uint64_t db0 = umma_desc_nosw(smem_u32(wring), 2u * N * 16u, 128u);
            const uint32_t step16 = kstep_bytes >> 4;
            const uint32_t wf0 = sv_opaque(smem_u32(&w_full[0])), we0 = sv_opaque(smem_u32(&w_empty[0]));
            const uint32_t gf0 = sv_opaque(smem_u32(&g_full[0])), ge0 = sv_opaque(smem_u32(&g_empty[0]));
            const uint32_t af0 = sv_opaque(smem_u32(&acc_full[0])), ae0 = sv_opaque(smem_u32(&acc_empty[0]));
            uint32_t st = 0, wph = 0, seq = 0;                  // basis ring: slot, parity; running slot number (trace)
            SvTile t;
            for (uint32_t it = 0; sv_tile(p, (int)it, t); ++it) {
                const uint32_t as = it & 1u;
                sv_wait1<PROF>(ae0 + 8u * as, ((it >> 1) & 1u) ^ 1u, pc0);
                tc_fence_after();
                const uint32_t acc = tmem_base + as * 2u * N;
                const uint2* tab = reinterpret_cast<const uint2*>(s_tab[t.slot]);
                const int nslots = p.br[t.slot].n_ksteps >> 1;
                const uint32_t gpar = it & 1u;
                uint32_t accum = 0;
                for (int s2 = 0; s2 < nslots; ++s2) {
                    const uint2 e = tab[s2];                    // the two K steps of this ring slot
                    sv_wait1<PROF>(wf0 + 8u * st, wph, pc2);
                    if (PROF && trace && seq < SV_TRACE_SLOTS) trace[seq * 3 + 1] = (unsigned long long)(clock64() - t_begin);
                    const uint64_t db = db0 + (uint64_t)(2u * st * step16);
                    if (e.x & SV_FIRST) sv_wait1<PROF>(gf0 + ((e.x >> 13) & 0x7F8u), gpar, pc1);
                    tc_fence_after();
                    umma_f16(acc, da_hi0 + (uint64_t)(e.x & 0xFFFFu), db, idesc2, accum);
                    umma_f16(acc + N, da_lo0 + (uint64_t)(e.x & 0xFFFFu), db, idesc1, 1u);
                    if (e.x & SV_LAST) sv_commit(ge0 + ((e.x >> 13) & 0x7F8u));
                    if (e.y & SV_FIRST) { sv_wait1<PROF>(gf0 + ((e.y >> 13) & 0x7F8u), gpar, pc1); tc_fence_after(); }
                    umma_f16(acc, da_hi0 + (uint64_t)(e.y & 0xFFFFu), db + step16, idesc2, 1u);
                    umma_f16(acc + N, da_lo0 + (uint64_t)(e.y & 0xFFFFu), db + step16, idesc1, 1u);
                    if (e.y & SV_LAST) sv_commit(ge0 + ((e.y >> 13) & 0x7F8u));
                    if (CL) sv_commit_mc(we0 + 8u * st, (uint16_t)3); else sv_commit(we0 + 8u * st);
                    if (PROF && trace && seq < SV_TRACE_SLOTS) trace[seq * 3 + 2] = (unsigned long long)(clock64() - t_begin);
                    ++seq;
                    accum = 1u;
                    if (++st == NS) { st = 0; wph ^= 1u; }
                }
                sv_commit(af0 + 8u * as);
                if (prof) p.prof[4] = it + 1;
            }
            if (prof) { p.prof[0] = pc0; p.prof[1] = pc1; p.prof[2] = pc2; p.prof[3] = (unsigned long long)(clock64() - t_begin); }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ loader: basis ring, two K steps per slot ================================
        // One 12 KB bulk copy per slot; the packed basis is already in issue order.  The control lane waits for this stream
        // more than for anything else (~40 % of the kernel: a slot arrives every ~600 cycles, the MMAs would take one every
        // 288).  Measured and without effect: a sixth ring slot, sharing the stream with a cluster peer by multicast (half
        // the bytes per SM), and moving the second K step of each slot to 16-byte cp.async (from this warp's 32 lanes, or
        // from a dedicated third loader warp; completion through cp.async.mbarrier.arrive.noinc) - so it is neither ring
        // depth nor the byte rate of the bulk-copy engine.  The ring trace (tools/fe_ring_trace.py) shows a copy landing a
        // median 3.1 k cycles after its issue with five in flight = ~19 B/clk into the SM; giving every CTA its own copy of
        // the basis (4 or 16 replicas) changes nothing either, so it is not contention on the shared L2 lines.
        if (elect_one()) {
            const uint32_t wf0 = sv_opaque(smem_u32(&w_full[0])), we0 = sv_opaque(smem_u32(&w_empty[0])), ring0 = sv_opaque(smem_u32(wring));
            const uint32_t slot_bytes = 2u * kstep_bytes;
            const uint32_t rank = CL ? sv_cluster_rank() : 0u;
            uint32_t st = 0, wph = 1, seq = 0;
            SvTile t;
            for (uint32_t it = 0; sv_tile(p, (int)it, t); ++it) {
                const SpecBranchDev& br = p.br[t.slot];
                const uint8_t* src = reinterpret_cast<const uint8_t*>(br.wpack);
                for (int s2 = 0; s2 < (br.n_ksteps >> 1); ++s2, src += slot_bytes, ++seq) {
                    sv_wait1<PROF>(we0 + 8u * st, wph, pc0);
                    if (PROF && trace && seq < SV_TRACE_SLOTS) trace[seq * 3 + 0] = (unsigned long long)(clock64() - t_begin);
                    sv_expect_tx(wf0 + 8u * st, slot_bytes);
                    if (!CL) sv_bulk_g2s(ring0 + st * slot_bytes, src, slot_bytes, wf0 + 8u * st);
                    else if ((seq & 1u) == rank) sv_bulk_g2s_mc(ring0 + st * slot_bytes, src, slot_bytes, wf0 + 8u * st, (uint16_t)3);
                    if (++st == NS) { st = 0; wph ^= 1u; }
                }
            }
            if (prof) { p.prof[9] = pc0; p.prof[10] = (unsigned long long)(clock64() - t_begin); }
        }
        __syncwarp();
    } else if (warp < 2 + SV_EPI_WARPS) {
        // ================================ epilogue ================================
        const uint32_t q = (uint32_t)warp & 3u;                 // TMEM lane quarter this warp may read
        const uint32_t par = (uint32_t)(warp - 2) >> 2;         // which of the alternating 16-column chunks
        const bool paired = sv_paired(p);
        const uint32_t stash = tmem_base + ((q * 32u) << 16) + 4u * N;      // N spare TMEM columns behind the two accumulator sets
        const bool first_is_ch0 = p.br[0].ch == 0;
        SvTile t;
        for (uint32_t it = 0; sv_tile(p, (int)it, t); ++it) {
            const SpecBranchDev& br = p.br[t.slot];
            const uint32_t as = it & 1u;
            sv_wait<PROF>(&acc_full[as], (it >> 1) & 1u, pc0);
            tc_fence_after();
            const int tt = t.t0 + (int)q * 32 + lane;
            const bool row_ok = tt < p.n_frames;
            const uint32_t t_lane = tmem_base + ((q * 32u) << 16) + as * 2u * N;
            const size_t mel_stride = (size_t)p.n_frames * p.n_ch;
            const size_t pix = ((size_t)t.seg * p.n_mels * p.n_frames + (size_t)tt) * p.n_ch;
            const float ex = br.exponent;
            for (uint32_t c0 = par * 16u; c0 < N; c0 += 16u * (SV_EPI_WARPS / 4)) {
                uint32_t rm[16], rc[16], pk[16];
                tmem_ld16_nowait(t_lane + c0, rm);
                tmem_ld16_nowait(t_lane + N + c0, rc);
                if (paired && t.slot == 1) tmem_ld16_nowait(stash + c0, pk);      // the other branch's hi | lo << 16 of the same pixels
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
                    const float pw = v * v;
                    const float r = pw > 0.f ? exp2f(ex * __log2f(pw)) : pw;
                    const __half hh = __float2half_rn(r);
                    const __half ll = __float2half_rn(r - __half2float(hh));
                    const uint32_t mine = (uint32_t)__half_as_ushort(hh) | ((uint32_t)__half_as_ushort(ll) << 16);
                    if (!paired) {
                        if (row_ok && (int)c0 + j < p.n_mels) {
                            __half* o = p.out_hi + pix + br.ch + (size_t)(c0 + j) * mel_stride;
                            o[0] = hh;
                            o[p.out_plane] = ll;
                        }
                    } else if (t.slot == 0) {
                        pk[j] = mine;
                    } else if (row_ok && (int)c0 + j < p.n_mels) {
                        const uint32_t c_first = pk[j], c_second = mine;        // slot 0's branch, slot 1's branch
                        const uint32_t a0 = first_is_ch0 ? c_first : c_second, a1 = first_is_ch0 ? c_second : c_first;
                        uint32_t* o = reinterpret_cast<uint32_t*>(p.out_hi + pix + (size_t)(c0 + j) * mel_stride);
                        o[0] = (a0 & 0xFFFFu) | (a1 << 16);                                 // hi plane: channel 0 | channel 1
                        o[p.out_plane >> 1] = (a0 >> 16) | (a1 & 0xFFFF0000u);              // lo plane
                    }
                }
                if (paired && t.slot == 0) tmem_st16(stash + c0, pk);
            }
            if (paired && t.slot == 0) tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (prof && warp == 2 && lane == 0) { p.prof[7] = pc0; p.prof[8] = (unsigned long long)(clock64() - t_begin); }
    } else {
        // ================================ producers: raw audio -> normalised hi / lo sample patch ================================
        const int pw = warp - 2 - SV_EPI_WARPS;
        const uint32_t RP = (uint32_t)p.row_pitch;
        SvTile t;
        for (uint32_t it = 0; sv_tile(p, (int)it, t); ++it) {
            const SpecBranchDev& br = p.br[t.slot];
            const float* xs = p.audio + (size_t)t.seg * p.S;
            float lo = sv_key2f(__ldg(p.minmax + 2 * t.seg));
            const float hi_v = sv_key2f(__ldg(p.minmax + 2 * t.seg + 1));
            if (hi_v != hi_v) lo = hi_v;                        // any NaN poisons the whole segment, like torch
            const float den = __fadd_rn(__fsub_rn(hi_v, lo), p.eps);
            const float rden = __frcp_rn(den);
            const bool den_ok = den > 1e-30f && den < 1e30f;    // otherwise (NaN / Inf / denormal range) take the IEEE division
            const int groups = (br.kcells + 1) >> 1;
            const bool even_hop = (br.hop & 1) == 0;
            // the samples of a tile are one contiguous range: pull the NEXT tile's range into L2 while this one is converted
            // (the audio is larger than L2, so after the min/max pass most of it is back in HBM)
            if (pw == 0 && lane == 0) {
                SvTile nt;
                if (sv_tile(p, (int)it + 1, nt)) {
                    const SpecBranchDev& nb = p.br[nt.slot];
                    const long long s_beg = (long long)nt.t0 * nb.hop;
                    long long s_end = s_beg + (long long)nb.rows * nb.hop + 16;
                    if (s_end > p.S) s_end = p.S;
                    const float* g = p.audio + (size_t)nt.seg * p.S + s_beg;
                    const uintptr_t a0 = reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)15;
                    const uint32_t bytes = (uint32_t)(((reinterpret_cast<uintptr_t>(g) + (size_t)(s_end - s_beg) * 4 - a0) + 15) & ~(size_t)15);
                    if (s_end > s_beg && !(p.debug & 256))
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(bytes) : "memory");
                }
            }
            const int rot = sv_rot(p, groups);
            for (int mi = pw; mi < groups; mi += SV_PROD_WARPS) {
                const int m = mi + rot < groups ? mi + rot : mi + rot - groups;
                sv_wait<PROF>(&g_empty[m], (it & 1u) ^ 1u, pc0);
                uint8_t* col = patch + (size_t)(2 * m) * RP * 16u;          // cells 2m (and 2m + 1, RP cells further)
                const bool two = 2 * m + 1 < br.kcells;                     // odd cell count: the last pair has one weighted cell
                // normalised value of raw sample x (samples past the segment end are exactly zero)
                auto norm = [&](float x, bool valid) {
                    // (x - lo) / den as reciprocal + one residual correction: the IEEE quotient in all but rare half-ulp
                    // cases at 3 instructions instead of ~18 (the producers were issue-bound); the value is split to 22
                    // bits right after.  Degenerate denominators (NaN / Inf / denormal range) take the IEEE division.
                    const float a = __fsub_rn(x, lo);
                    const float q0 = __fmul_rn(a, rden);
                    float q = __fmaf_rn(__fmaf_rn(-q0, den, a), rden, q0);
                    if (!den_ok) q = sv_ieee_div(a, den);
                    return valid ? __fmul_rn(__fsub_rn(q, p.half), p.two) : 0.f;
                };
                const int full = br.rows & ~63;                             // rows the two-row passes cover
                if (!(p.debug & 1)) {
                for (int r0 = lane; r0 < full; r0 += 64) {
                    // two rows per pass (both cells of the pair: 16 samples each), every load issued before the first use
                    float v[2][16];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int s0 = (t.t0 + r0 + 32 * u) * br.hop + m * 16;
                        if (even_hop && s0 + 16 <= p.S) {
                            const float2* g = reinterpret_cast<const float2*>(xs + s0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) { const float2 f = __ldg(g + i); v[u][2 * i] = f.x; v[u][2 * i + 1] = f.y; }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[u][i] = s0 + i < p.S ? __ldg(xs + s0 + i) : lo;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int r = r0 + 32 * u;
                        const int nvalid = p.S - ((t.t0 + r) * br.hop + m * 16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[u][i] = norm(v[u][i], i < nvalid);
                        uint4 h, l;
                        split8(v[u], h, l);
                        *reinterpret_cast<uint4*>(col + (size_t)r * 16u) = h;
                        *reinterpret_cast<uint4*>(col + plane + (size_t)r * 16u) = l;
                        if (two) {
                            split8(v[u] + 8, h, l);
                        } else {
                            h = make_uint4(0u, 0u, 0u, 0u);
                            l = h;
                        }
                        *reinterpret_cast<uint4*>(col + (size_t)(RP + r) * 16u) = h;
                        *reinterpret_cast<uint4*>(col + plane + (size_t)(RP + r) * 16u) = l;
                    }
                }
                // the few rows past the last full pass (7 of 135): one (row, cell) item of 8 samples per lane
                for (int idx = lane; idx < 2 * (br.rows - full); idx += 32) {
                    const int r = full + (idx >> 1), kc = idx & 1;
                    const int s0 = (t.t0 + r) * br.hop + m * 16 + kc * 8;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = s0 + i < p.S ? __ldg(xs + s0 + i) : lo;
                    const int nvalid = p.S - s0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = norm(v[i], i < nvalid);
                    uint4 h, l;
                    split8(v, h, l);
                    if (kc == 1 && !two) { h = make_uint4(0u, 0u, 0u, 0u); l = h; }
                    *reinterpret_cast<uint4*>(col + (size_t)(kc * RP + r) * 16u) = h;
                    *reinterpret_cast<uint4*>(col + plane + (size_t)(kc * RP + r) * 16u) = l;
                }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&g_full[m]);
            }
        }
        if (prof && pw == 0 && lane == 0) { p.prof[5] = pc0; p.prof[6] = (unsigned long long)(clock64() - t_begin); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    if (CL) sv_cluster_sync();                                  // nobody leaves while the peer may still signal or copy into it
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------------
static inline uint16_t sv_f2h(float f) {
    return __half_raw(__float2half_rn(f)).x;
}
static inline float sv_h2f(uint16_t u) {
    __half_raw r;
    r.x = u;
    return __half2float(__half(r));
}

bool spec_v24_plan(int n_fft, int hop, int n_mels, SpecBranchHost& out) {
    out = SpecBranchHost{};
    if (hop <= 0 || n_fft < hop || n_mels <= 0 || n_mels > 128) return false;
    out.hop = hop;
    out.kcells = (hop + 7) / 8;
    out.blocks = (n_fft + hop - 1) / hop;
    out.rows = 128 + out.blocks - 1;
    out.n_pad = (n_mels + 15) / 16 * 16;
    const int last_w = n_fft - (out.blocks - 1) * hop;
    const int last_cells = (last_w + 7) / 8;
    const int groups = (out.kcells + 1) / 2;
    if (groups > SV_MAX_GROUPS) return false;
    out.split = std::min(groups, (last_cells + 1) / 2);
    out.table.clear();
    for (int m = 0; m < groups; ++m) {
        const int jn = m < out.split ? out.blocks : out.blocks - 1;
        for (int j = 0; j < jn; ++j) out.table.push_back(SpecKStep{m, j, false});
    }
    out.pad = (int)(out.table.size() & 1);
    if (out.pad) {                              // an even count keeps the ring arithmetic simple: one zero-weight step at the
        const int m = groups - 1;               // end of the last group, one more block down (rows it reads exist and are finite)
        out.table.push_back(SpecKStep{m, (m < out.split ? out.blocks : out.blocks - 1), true});
        out.rows_read = std::max(out.rows, 128 + out.table.back().j);
    } else {
        out.rows_read = out.rows;
    }
    return (int)out.table.size() <= SV_MAX_KSTEPS;
}

void spec_v24_pack(const SpecBranchHost& b, const float* basis, int ldb, int n_fft, int n_mels, std::vector<uint16_t>& wpack) {
    const int N = b.n_pad;
    wpack.assign((size_t)b.table.size() * 2 * 2 * N * 8, 0);
    for (size_t ks = 0; ks < b.table.size(); ++ks) {
        const SpecKStep& k = b.table[ks];
        if (k.zero) continue;
        uint16_t* dst = wpack.data() + ks * (size_t)(2 * 2 * N * 8);
        for (int kc = 0; kc < 2; ++kc)
            for (int kk = 0; kk < 8; ++kk) {
                const int col = (2 * k.m + kc) * 8 + kk;
                const int n = k.j * b.hop + col;
                if (col >= b.hop || n >= n_fft) continue;
                for (int mel = 0; mel < n_mels; ++mel) {
                    const float w = basis[(size_t)n * ldb + mel];
                    const uint16_t h = sv_f2h(w);
                    dst[((size_t)kc * 2 * N + mel) * 8 + kk] = h;
                    dst[((size_t)kc * 2 * N + N + mel) * 8 + kk] = sv_f2h(w - sv_h2f(h));
                }
            }
    }
}

bool spec_v24_layout(const SpecBranchHost* br, int n_br, int& row_pitch, uint32_t& patch_plane, int& n_stages, uint32_t& smem_bytes) {
    int rows = 0, cols = 0, n_pad = 0;
    for (int i = 1; i < n_br; ++i)
        if ((br[i].kcells + 1) / 2 != (br[0].kcells + 1) / 2) return false;     // the group barriers flip once per tile
    for (int i = 0; i < n_br; ++i) {
        rows = std::max(rows, br[i].rows_read);
        cols = std::max(cols, (br[i].kcells + 1) / 2 * 2);
        if (n_pad && n_pad != br[i].n_pad) return false;
        n_pad = br[i].n_pad;
    }
    row_pitch = rows;                                       // any pitch works (cells are 16 bytes); exact rows leave room for one more ring slot
    patch_plane = (uint32_t)cols * (uint32_t)row_pitch * 16u;
    const uint32_t stage = 128u * (uint32_t)n_pad;           // a ring slot = two K steps
    const uint32_t fixed = 2u * patch_plane + 128u;
    if (fixed + 2u * stage > SV_SMEM_MAX) return false;
    n_stages = (int)std::min<uint32_t>((SV_SMEM_MAX - fixed) / stage, SV_MAX_STAGES);
    smem_bytes = fixed + (uint32_t)n_stages * stage;
    return 2 * 2 * n_pad <= 512 && n_stages >= 2;
}

static int g_sv_cluster = -1;       // -1 unknown, 0 plain launch, 1 clusters of two

cudaError_t spec_v24_init_device() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_spec_v24<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SV_SMEM_MAX)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_spec_v24<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SV_SMEM_MAX)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_spec_v24<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SV_SMEM_MAX)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_spec_v24<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SV_SMEM_MAX)) != cudaSuccess) return e;
    return cudaSuccess;
}

// clusters of two need every pair co-resident: ask the occupancy calculator once (BN_FE_CLUSTER=1 opts in)
static bool sv_use_cluster(int grid, uint32_t smem_bytes) {
    if (g_sv_cluster < 0) {
        const char* ev = getenv("BN_FE_CLUSTER");
        g_sv_cluster = 0;
        if (ev && ev[0] == '1') {                           // opt-in: measured equal to the plain launch
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(SV_THREADS); cfg.dynamicSmemBytes = smem_bytes;
            cudaLaunchAttribute at{};
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_spec_v24<false, true>, &cfg) == cudaSuccess && 2 * n >= grid) g_sv_cluster = 1;
            else (void)cudaGetLastError();
        }
    }
    return g_sv_cluster == 1;
}

cudaError_t launch_spec_v24(const SpecV24Params& p, uint32_t smem_bytes, int num_sms, cudaStream_t stream) {
    if (p.batch <= 0) return cudaSuccess;
    const int per = p.batch * p.tiles_per_seg;
    const int grid = std::min(num_sms, per);
    // pairs must walk identical slot sequences: an even grid and an even tile count per branch give both CTAs of a pair
    // the same number of tiles (per - 1 - bx is odd for both, grid is even)
    const bool cl = (grid % 2 == 0) && (per % 2 == 0) && (grid == num_sms) && sv_use_cluster(grid, smem_bytes);
    if (!cl) {
        if (p.prof) k_spec_v24<true, false><<<grid, SV_THREADS, smem_bytes, stream>>>(p);
        else k_spec_v24<false, false><<<grid, SV_THREADS, smem_bytes, stream>>>(p);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(SV_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    return p.prof ? cudaLaunchKernelEx(&cfg, k_spec_v24<true, true>, p) : cudaLaunchKernelEx(&cfg, k_spec_v24<false, true>, p);
}

}  // namespace bn

// Host-only self check of the K-step schedule and the packed basis (development symbol, not part of include/birdnet_b200.h;
// tests/test_frontend_v24_host.py calls it on CPU): for a deterministic pseudo-random signal and basis, the sum over the
// schedule - patch cells addressed exactly as the kernel's descriptors address them (cell column, row shift, LBO = one
// column) times the packed hi + lo weights - against the direct dot product, in double.  Returns the worst difference;
// out[0..2] = K steps, ring slots, shared-memory bytes of the plan.  -1 = shape not covered by the kernel.
extern "C" double bn_debug_spec_v24_selfcheck(int n_fft, int hop, int n_mels, int sample_count, int* out) {
    using namespace bn;
    SpecBranchHost hb;
    if (!spec_v24_plan(n_fft, hop, n_mels, hb)) return -1.0;
    int rp = 0, ns = 0;
    uint32_t plane = 0, smem = 0;
    if (!spec_v24_layout(&hb, 1, rp, plane, ns, smem)) return -1.0;
    if (out) { out[0] = (int)hb.table.size(); out[1] = ns; out[2] = (int)smem; }
    const int ldb = (n_mels + 3) / 4 * 4, N = hb.n_pad, S = sample_count;
    const int T = 1 + (S - n_fft) / hop;
    std::vector<float> x((size_t)S), basis((size_t)n_fft * ldb);
    uint32_t lcg = 12345u;
    auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; return (float)((lcg >> 8) & 0xffff) / 65536.f - 0.5f; };
    for (auto& v : x) v = rnd();
    for (auto& v : basis) v = rnd() * 0.25f;
    std::vector<uint16_t> pack;
    spec_v24_pack(hb, basis.data(), ldb, n_fft, n_mels, pack);
    // the device's own enumeration (column pair outer, blocks inner, zero step at the end of the last pair)
    std::vector<uint32_t> cell0;
    const int groups = (hb.kcells + 1) / 2;
    for (int m = 0; m < groups; ++m) {
        const int jn = (m < hb.split ? hb.blocks : hb.blocks - 1) + (m == groups - 1 ? hb.pad : 0);
        for (int j = 0; j < jn; ++j) cell0.push_back((uint32_t)(2 * m) * (uint32_t)rp + (uint32_t)j);
    }
    if (cell0.size() != hb.table.size()) return 1e30;
    const int cols = (hb.kcells + 1) / 2 * 2;
    double worst = 0.0;
    const int tiles = (T + 127) / 128;
    const int t0s[3] = {0, 128 * (tiles / 2), 128 * (tiles - 1)};
    for (int ti = 0; ti < 3; ++ti) {
        const int t0 = t0s[ti];
        std::vector<float> patch((size_t)cols * rp * 8, 0.f);
        for (int c = 0; c < hb.kcells; ++c)
            for (int r = 0; r < hb.rows; ++r)
                for (int i = 0; i < 8; ++i) {
                    const long long sidx = (long long)(t0 + r) * hop + c * 8 + i;
                    patch[((size_t)c * rp + r) * 8 + i] = sidx < S ? x[(size_t)sidx] : 0.f;
                }
        const int rows_i[5] = {0, 1, 63, 126, 127};
        const int mels[3] = {0, n_mels / 2, n_mels - 1};
        for (int ri = 0; ri < 5; ++ri) {
            const int i = rows_i[ri], t = t0 + i;
            if (t >= T) continue;
            for (int mi = 0; mi < 3; ++mi) {
                const int mel = mels[mi];
                double acc = 0.0;
                for (size_t ks = 0; ks < cell0.size(); ++ks) {
                    const uint16_t* w = pack.data() + ks * (size_t)(2 * 2 * N * 8);
                    for (int kc = 0; kc < 2; ++kc)
                        for (int kk = 0; kk < 8; ++kk) {
                            const float a = patch[((size_t)cell0[ks] + (size_t)kc * rp + i) * 8 + kk];
                            const float wv = sv_h2f(w[((size_t)kc * 2 * N + mel) * 8 + kk]) + sv_h2f(w[((size_t)kc * 2 * N + N + mel) * 8 + kk]);
                            acc += (double)a * (double)wv;
                        }
                }
                double ref = 0.0;
                for (int n = 0; n < n_fft; ++n) {
                    const float wq = basis[(size_t)n * ldb + mel];
                    const uint16_t h = sv_f2h(wq);
                    ref += (double)x[(size_t)t * hop + n] * ((double)sv_h2f(h) + (double)sv_h2f(sv_f2h(wq - sv_h2f(h))));
                }
                worst = std::max(worst, fabs(acc - ref));
            }
        }
    }
    return worst;
}

