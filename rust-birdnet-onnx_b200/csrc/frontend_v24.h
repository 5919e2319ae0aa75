// Fused v2.4 audio front-end (see frontend_v24.cu): normaliser + both mel-spectrogram GEMMs in one tcgen05 kernel.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace bn {

constexpr int SV_MAX_KSTEPS = 192;

struct SpecBranchDev {
    const void* wpack;        // [k-step][cell 2][W_hi rows n_pad | W_lo rows n_pad][8] fp16 (spec_v24_pack), schedule order
    int hop, kcells, rows;    // samples per patch row, 8-sample cells per row, rows a 128-frame tile needs
    int blocks, split, pad;   // hop-sized blocks of a frame; column pairs < split carry every block, the rest one fewer; zero step
    int n_ksteps, ch;
    float exponent;
};

struct SpecV24Params {
    const float* audio;       // [batch][S] raw samples
    const uint32_t* minmax;   // [batch][2] order-preserving keys of min / max (k_minmax_partial)
    SpecBranchDev br[2];      // slot 0 = the branch with the longer K loop
    int n_br;
    __half* out_hi;           // [batch][n_mels][n_frames][n_ch] hi plane, lo = out_hi + out_plane
    size_t out_plane;
    int batch, S, n_frames, n_mels, n_pad, n_ch, tiles_per_seg;
    int row_pitch, n_stages;
    uint32_t patch_plane;
    float eps, half, two;
    int debug;                // development experiments (BN_FE_DEBUG: 1 = producers write nothing, 256 = no L2 prefetch), 0 in production
    unsigned long long* prof; // optional [16] role cycle counters of CTA 0 (development aid, BN_FE_PROFILE), or nullptr
};

struct SpecKStep { int m, j; bool zero; };
struct SpecBranchHost {
    int hop = 0, kcells = 0, blocks = 0, rows = 0, rows_read = 0, n_pad = 0, split = 0, pad = 0;
    std::vector<SpecKStep> table;
};

// K-step schedule of one branch (cell-column pair outer, block inner); false = shape outside what the kernel covers
bool spec_v24_plan(int n_fft, int hop, int n_mels, SpecBranchHost& out);
// shared-memory plan for the branches together; false = does not fit
bool spec_v24_layout(const SpecBranchHost* br, int n_br, int& row_pitch, uint32_t& patch_plane, int& n_stages, uint32_t& smem_bytes);
// basis [n_fft][ldb] FP32 -> packed hi / lo K steps in schedule order
void spec_v24_pack(const SpecBranchHost& b, const float* basis, int ldb, int n_fft, int n_mels, std::vector<uint16_t>& wpack);
cudaError_t spec_v24_init_device();
cudaError_t launch_spec_v24(const SpecV24Params& p, uint32_t smem_bytes, int num_sms, cudaStream_t stream);

}  // namespace bn
