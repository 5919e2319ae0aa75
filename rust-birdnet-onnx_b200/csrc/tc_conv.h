// tcgen05 implicit-GEMM conv / dense layer (see tc_conv.cu).
//
// Activation storage "planes": a spatial tensor [B][H][W][C] is kept in HBM as two fp16 planes
// of identical NHWC shape, hi = fp16(x) and lo = fp16(x - hi) (x = hi + lo to ~22 bits, the
// same 4 bytes per element as FP32).  The split is done ONCE by the producing kernel's epilogue,
// so the tensor-core consumers can move operands with cp.async and no ALU work.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "kernels.h"

namespace bn {

enum TcInMode : int { TC_IN_PLANES = 0, TC_IN_PLANES_SCALED = 1, TC_IN_F32 = 2 };

struct TcConvParams {
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
    int spec_nframes, spec_nch, spec_ch;
    float spec_exponent;
    const void* wpack;       // [n_tiles][k_chunks][hi|lo][nt x 64] fp16, SW128 K-major smem images
    int in_mode;
    int batch, hin, win, cin, hout, wout, cout, k, stride, pad, act;
    int K, M, k_chunks, m_tiles, n_tiles;
    int pix_stride;          // elements between consecutive input pixels (= cin for a conv)
    int seg_stride;          // elements between consecutive segments of the input
    int tab_cin;             // channel count used to split K into (tap, channel) (= cin for a conv)
    int nt;                  // UMMA N of this layer (multiple of 16, <= 256)
    int stages;              // smem ring depth
    int tmem_cols;           // power of two >= max(32, 2 * nt): two accumulators
};

cudaError_t tc_conv_init_device();
cudaError_t launch_tc_conv(const TcConvParams& p, int num_sms, cudaStream_t stream);
size_t tc_conv_smem_bytes(int nt, int stages);
int tc_conv_pick_stages(int nt, int k_chunks);
void tc_pack_weights(const float* w, int K, int cout, int ldw, int nt, std::vector<uint16_t>& out,
                     int* n_tiles_out, int* k_chunks_out);

}  // namespace bn
