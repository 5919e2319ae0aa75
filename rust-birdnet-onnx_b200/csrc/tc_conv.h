// tcgen05 implicit-GEMM conv / dense layer (see tc_conv.cu).
//
// Activation storage "planes": a spatial tensor [B][H][W][C] is kept in HBM as two fp16 planes
// of identical NHWC shape, hi = fp16(x) and lo = fp16(x - hi) (x = hi + lo to ~22 bits, the
// same 4 bytes per element as FP32).  The split is done ONCE by the producing kernel's epilogue,
// so the tensor-core consumers can move operands with cp.async and no ALU work.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "kernels.h"

namespace bn {

enum TcInMode : int { TC_IN_PLANES = 0, TC_IN_PLANES_SCALED = 1, TC_IN_F32 = 2, TC_IN_TMA = 3, TC_IN_HALO = 4 };

struct TcConvParams {
    // TC_IN_TMA: 5-D tensor map (channel, x, y, segment, plane) over the input hi/lo planes; the A
    // tile of a K block is ONE cp.async.bulk.tensor per plane (zero fill outside the image = padding)
    alignas(64) CUtensorMap tmap;
    // output tensor map (out_tma != 0): the epilogue warps hand their staged [32 rows][32 columns] tiles to the TMA
    // engine instead of re-reading them and storing 16-byte pieces.  planes: 3-D (channel, row, plane) fp16, box
    // {32,32,2}, SWIZZLE_64B; FP32 rows: 2-D (channel, row) f32, box {32,32}, SWIZZLE_128B.
    alignas(64) CUtensorMap omap;
    int out_tma;
    int kb;                  // channels per TMA box: 16 / 32 / 64 <-> SWIZZLE_32B / 64B / 128B
    int box_w, box_h;        // output pixels per tile row x rows per tile (box_w * box_h == 128), rect mode
    int flat;                // 1: A is a flat [M][K] matrix (1x1 stride-1 conv, front-end frames)
    // tile -> rows: m tile mt covers rows [lt*128, lt*128+128) of segment b = mt / tiles_per_seg
    // (cp.async modes: tiles_per_seg = m_tiles, pix_per_seg = M, i.e. one flat "segment")
    int tiles_per_seg, pix_per_seg;
    // input: planes (lo = in_hi + in_plane) or FP32
    const __half* in_hi;
    size_t in_plane;
    const float* in_f32;
    const float* in_scale;   // [B][cin] squeeze-excite gate (TC_IN_PLANES_SCALED)
    // residual (planes) added after the activation, or nullptr
    const __half* res_hi;
    size_t res_plane;
    const float* bias;       // [cout]
    // output: planes, or FP32 when out_f32 != nullptr, or (spec_nframes > 0) the v2.4 spectrogram
    // epilogue: v -> pow(v*v, spec_exponent) stored at [b][n][t][spec_ch] of an [B][cout][nframes][nch] plane pair
    __half* out_hi;
    size_t out_plane;
    float* out_f32;
    int out_f32_rows;        // 1: out_f32 is a spatial [M][cout] FP32 tensor written by the staged (coalesced) epilogue
    int spec_nframes, spec_nch, spec_ch;
    float spec_exponent;
    const void* wpack;       // [n_tiles][k_chunks][hi|lo][nt x 64] fp16, SW128 K-major smem images
    int in_mode;
    int batch, hin, win, cin, hout, wout, cout, k, stride, pad, act;
    int K, M, k_chunks, m_tiles, n_tiles;
    int pix_stride;          // elements between consecutive input pixels (= cin for a conv)
    int seg_stride;          // elements between consecutive segments of the input
    int tab_cin;             // channel count used to split K into (tap, channel) (= cin for a conv)
    unsigned long long* prof;   // optional [16] cycle counters written by CTA 0 (tools/tc_role_profile.py), or nullptr
    int nt;                  // UMMA N of this layer (multiple of 16, <= 256)
    int stages;              // smem ring depth
    int tmem_cols;           // power of two >= max(32, 2 * nt): two accumulators
    int epi_warps;           // one CTA per SM: 8 or 16 epilogue warps (0 = build default)
    int debug_flags;         // development experiments (BN_TC_DEBUG), 0 in production
    int w_resident;          // TC_IN_TMA: this CTA's weight tile (all K chunks) stays in smem, the ring carries A only; set by launch_tc_conv
};

cudaError_t tc_conv_init_device();
// development aid: per-launch role wait-cycle counters (slot < 128, 16 counters each); see tc_conv.cu
unsigned long long* tc_conv_prof_slot(int slot);
cudaError_t tc_conv_prof_read(unsigned long long* out, int slots);
cudaError_t launch_tc_conv(const TcConvParams& p, int num_sms, cudaStream_t stream);
size_t tc_conv_smem_bytes(int nt, int stages, int epi_warps);
int tc_conv_pick_stages(int nt, int k_chunks, int epi_warps = 0);
// TC_IN_HALO (3x3 stride-1 pad-1 conv, wout % 128 == 0, cin % 16 == 0): weights resident in smem,
// each input pixel staged ONCE per tile (3 rows x 130 pixels) and the nine taps read as row-shifted
// views of that patch.  Returns the number of patch slots (2 or 3) that fit, 0 if the mode does not apply.
int tc_conv_halo_slots(int k, int stride, int pad, int cin, int wout, int win, int nt, int k_chunks, int epi_warps = 0);
size_t tc_conv_halo_smem_bytes(int cin, int nt, int k_chunks, int slots, int epi_warps = 0);
// Encodes the 5-D (channel, x, y, segment, plane) fp16 tensor map; returns false if the driver refuses.
bool tc_encode_tmap(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                    const uint32_t box[5], const uint32_t elem_strides[5], int kb);
// Encodes the output tensor map described at TcConvParams::omap; returns false if the driver refuses.
bool tc_encode_out_tmap(CUtensorMap* out, void* base, uint64_t rows, uint64_t cout, bool f32, uint64_t plane_elems);
void tc_pack_weights(const float* w, int K, int cout, int ldw, int nt, std::vector<uint16_t>& out,
                     int* n_tiles_out, int* k_chunks_out);

}  // namespace bn
