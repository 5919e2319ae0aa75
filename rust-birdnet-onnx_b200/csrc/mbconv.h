// Fused MBConv block kernel (see mbconv.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace bn {

struct MbconvParams {
    // 5-D (channel, row, 1, 1, plane) fp16 tensor maps with box {64 channels, 64 rows}, SWIZZLE_128B
    // (tc_encode_tmap): the block input X and the depthwise output D, both [batch * npix][channels] hi/lo planes
    alignas(64) CUtensorMap xmap;
    alignas(64) CUtensorMap dmap;
    const void* we_pack;      // expand weights, tc_pack_weights(nt = group size): [group][k_chunk][hi|lo][G x 64] fp16
    const float* be;          // [cexp] expand bias
    const float* wd;          // [k*k][cexp] depthwise weights
    const float* bd;          // [cexp] depthwise bias
    const float* w1;          // [cexp][ldw1] squeeze-excite reduce (r outputs)
    const float* b1;          // [r]
    const float* w2;          // [r][ldw2] squeeze-excite expand (cexp outputs)
    const float* b2;          // [cexp]
    const void* wp_pack;      // projection weights, tc_pack_weights(nt = cout): [k_chunk][hi|lo][cout x 64] fp16 (gate applied to D)
    const float* wpT;         // projection weights [cout][cexp] FP32, output-channel major (gate applied to the weights)
    const float* bp;          // [cout]
    __half* d_hi;             // D planes (un-gated), lo = d_hi + d_plane
    size_t d_plane;
    const __half* res_hi;     // residual (= the block input when cin == cout) or nullptr
    size_t res_plane;
    __half* out_hi;           // block output planes [batch][npix][cout]
    size_t out_plane;
    int batch, h, w, k, cin, cexp, cout, r, ldw1, ldw2;
    int segs;                 // segments per CTA pass (set by launch_mbconv)
    unsigned long long* prof; // optional [12] phase cycle counters of CTA 0 (development aid), or nullptr
    int debug;                // development experiments (BN_MB_DEBUG bits: skip parts of the work to time the rest), 0 in production
};

struct MbLayout {
    int n_box, n_mt, kc_e, kc_p, npixp, n_da;
    uint32_t xa_bytes, we_stage, da_stage, wp_stage, patch_bytes, stg_pitch;
    uint32_t off_xa, off_we, off_da, off_wp, off_patch, off_wd, off_part, off_pool, off_gate, off_r, off_fc, total;
};

cudaError_t mbconv_init_device();
// stride-1 MBConv block with squeeze-excite this build has a fused kernel for?
bool mbconv_supported(int h, int w, int k, int stride, int cin, int cexp, int cout, int r);
// channel group size the kernel uses for this shape (the N tile the expand weights must be packed with), 0 = unsupported
int mbconv_group(int h, int w, int k, int cin, int cexp, int cout, int r);
cudaError_t launch_mbconv(const MbconvParams& p, int num_sms, cudaStream_t stream);

}  // namespace bn
