// Launchers for the sm_100a kernels of the hot path.  All tensors are NHWC FP32 in HBM.
// Every launcher enqueues on `stream` and returns the cudaError_t of the launch.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace bn {

enum { KACT_NONE = 0, KACT_SILU = 1, KACT_SIGMOID = 2 };   // == plan.h Act

// must be called once per device before launch_topk / launch_range_filter (opt-in smem size)
cudaError_t init_kernels_for_device();

// C-ABI mirror (include/birdnet_b200.h: bn_pred)
struct Pred { uint32_t index; float confidence; };

// ---- front-end (row A7) -----------------------------------------------------------
// per-segment min/max normalise: y = ((x - min) / (max(x - min) + eps) - half) * two
cudaError_t launch_minmax_normalize(const float* x, float* y, int batch, int sample_count,
                                    float eps, float half, float two, cudaStream_t stream);

// framed DFT(real) x mel (pre-multiplied basis [n_fft][ldb]) -> square -> pow(exponent);
// writes channel `ch` of the NHWC spectrogram [B][n_mels][n_frames][n_ch]
cudaError_t launch_spectrogram_v24(const float* xnorm, const float* basis, int ldb, float* spec,
                                   int batch, int sample_count, int n_fft, int hop, int n_frames,
                                   int n_mels, int n_ch, int ch, float exponent, cudaStream_t stream);

// ---- CLI ingest moved on-device (section 8f row 1): i16 PCM -> overlapping FP32 segments ------------
// out[b][i] = pcm[first_pos + b*step + i - base] / 32768 (0 beyond n_total); `base` = recording index of pcm[0]
cudaError_t launch_pcm16_to_segments(const int16_t* pcm, uint64_t base, uint64_t n_total, uint64_t first_pos, uint64_t step,
                                     float* out, int batch, int sample_count, cudaStream_t stream);

// ---- CNN (row A8) -----------------------------------------------------------------
struct ConvParams {
    const float* in;        // [B][hin][win][cin]
    const float* in_scale;  // [B][cin] squeeze-excite gate or nullptr
    const float* weight;    // [k*k*cin][ldw]
    const float* bias;      // [cout]
    const float* residual;  // [B][hout][wout][cout] or nullptr (added after the activation)
    float* out;             // [B][hout][wout][cout]
    int batch, hin, win, cin, hout, wout, cout, ldw, k, stride, pad, act;
};
cudaError_t launch_conv_igemm(const ConvParams& p, cudaStream_t stream);

struct DwParams {
    const float* in;      // [B][hin][win][c]
    const float* weight;  // [k*k][c]
    const float* bias;    // [c]
    float* out;           // [B][hout][wout][c]
    int batch, hin, win, c, hout, wout, k, stride, pad, act;
};
cudaError_t launch_dwconv(const DwParams& p, cudaStream_t stream);

// global average pool: [B][hw][c] -> [B][c]
cudaError_t launch_gap(const float* in, float* out, int batch, int hw, int c, cudaStream_t stream);

// ---- "planes" storage variants (hi/lo fp16 planes, see tc_conv.h) ----------------------
struct PlanesPtr {
    __half* hi;        // lo = hi + plane
    size_t plane;
};
// normaliser that also emits, per spectrogram branch, the fp16 hi/lo "frame matrix" the
// tensor-core front-end GEMM reads: Xp[b][r][c] = y[r*hop + c] (c < hop, inside the segment) else 0,
// rows of row_stride = round_up(hop, 8) elements so every row start is 16-byte aligned and frame t
// is the contiguous run Xp[t*row_stride ...] (block-Toeplitz form of the overlapping frames).
struct FePlaneOut {
    __half* hi;
    size_t plane;       // lo = hi + plane
    int hop, row_stride, rows;
};
// y may be nullptr (normalised FP32 audio is only kept for tests); minmax_scratch: [batch][2] u32
cudaError_t launch_minmax_normalize_fe(const float* x, float* y, uint32_t* minmax_scratch, const FePlaneOut* outs, int n_outs,
                                       int batch, int sample_count, float eps, float half, float two, cudaStream_t stream);
// framed DFT x mel -> square -> pow, written as planes
cudaError_t launch_spectrogram_v24_planes(const float* xnorm, const float* basis, int ldb, PlanesPtr spec,
                                          int batch, int sample_count, int n_fft, int hop, int n_frames,
                                          int n_mels, int n_ch, int ch, float exponent, cudaStream_t stream);
struct ConvPlanesParams {
    PlanesPtr in;
    const float* in_scale;
    const float* weight;
    const float* bias;
    PlanesPtr residual;   // hi == nullptr -> none
    PlanesPtr out;
    int batch, hin, win, cin, hout, wout, cout, ldw, k, stride, pad, act;
};
cudaError_t launch_conv_igemm_planes(const ConvPlanesParams& p, cudaStream_t stream);
// direct conv for tiny Cin (stem): cout <= 32, no gate / residual
cudaError_t launch_stem_planes(const ConvPlanesParams& p, cudaStream_t stream);
struct DwPlanesParams {
    PlanesPtr in;
    const float* weight;  // [k*k][c]
    const float* bias;
    PlanesPtr out;
    float* pooled;        // [B][c] global average of the output (fused squeeze), or nullptr
    int batch, hin, win, c, hout, wout, k, stride, pad, act;
};
cudaError_t launch_dwconv_planes(const DwPlanesParams& p, cudaStream_t stream);
cudaError_t launch_gap_planes(PlanesPtr in, float* out, int batch, int hw, int c, cudaStream_t stream);
// fused squeeze-excite tail: two FCs on the pooled vector + in-place rescale of the planes tensor d
struct SeParams {
    const float* pooled;   // [B][c]
    const float* w1;       // [c][ldw1]  (r outputs)
    const float* b1;       // [r]
    const float* w2;       // [r][ldw2]  (c outputs)
    const float* b2;       // [c]
    float* gate_out;       // [B][c] or nullptr
    PlanesPtr d;           // [B][npix][c], rescaled in place
    int c, r, ldw1, ldw2, npix;
};
cudaError_t launch_se_scale(const SeParams& p, int batch, cudaStream_t stream);
// squeeze-excite as two launches: the gate (one CTA per segment, both FCs) and a pure streaming in-place
// rescale d[b][pix][c] *= gate[b][c] on the hi/lo planes
cudaError_t launch_se_gate(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                           float* gate, int batch, int c, int r, int ldw1, int ldw2, cudaStream_t stream);
cudaError_t launch_se_rescale(PlanesPtr d, const float* gate, int batch, int npix, int c, cudaStream_t stream);

// depthwise conv + SiLU + squeeze-excite in one kernel (dw_se.cu): one CTA per segment, un-gated output
// written, pooled, gated and rescaled in place while it is still L2-resident
struct DwSeParams {
    PlanesPtr in;            // [B][hin][win][c] expanded tensor as hi/lo planes (used when in_f32 == nullptr)
    const float* in_f32;     // the same tensor as plain FP32 (its producer wrote it for this kernel alone), or nullptr
    const float* weight;     // [k*k][c]
    const float* bias;       // [c]
    PlanesPtr out;           // [B][hout][wout][c] = silu(dw) * gate
    const float* w1;         // [c][ldw1] (r outputs)
    const float* b1;         // [r]
    const float* w2;         // [r][ldw2] (c outputs)
    const float* b2;         // [c]
    float* pooled_out;       // [B][c] or nullptr
    float* gate_out;         // [B][c] or nullptr
    int batch, hin, win, c, hout, wout, k, stride, pad, act, r, ldw1, ldw2;
    int debug;               // development experiments (BN_DW_DEBUG), 0 in production
};
cudaError_t dw_se_init_device();
bool dw_se_supported(const DwSeParams& p);
cudaError_t launch_dw_se(const DwSeParams& p, cudaStream_t stream);

// ---- log-mel front-end (row A9: BirdNET v3.0 / Perch v2), frontend_logmel.cu -------------------
struct LogmelParams {
    const float* audio;       // [B][sample_count] raw samples
    int batch, sample_count;
    const float* window;      // [n_fft]
    const float2* twiddle;    // [n_fft] exp(-2*pi*i*m/n_fft), built in float64
    const int* mel_lo;        // [n_mels] first bin of the filter
    const int* mel_cnt;       // [n_mels] bins in the filter
    const int* mel_off;       // [n_mels] offset of its weights in mel_w
    const float* mel_w;       // packed filter weights
    int n_fft, hop, n_frames, n_mels;
    int radix[8], n_stages;   // factorisation of n_fft / 2
    float log_floor, log_scale;
    PlanesPtr out;            // [B][n_frames][n_mels] hi/lo planes
    float* out_f32;           // optional FP32 copy of the same layout, or nullptr
};
cudaError_t logmel_init_device();
size_t logmel_smem_bytes(int n_fft);
bool logmel_factorize(int n_fft, int radix[8], int* n_stages);
cudaError_t launch_logmel(const LogmelParams& p, cudaStream_t stream);
void logmel_twiddles(int n_fft, std::vector<float>& out);   // interleaved re/im

// ---- epilogue (rows A4 / A6) ------------------------------------------------------
struct TopkParams {
    const float* logits;        // [B][n]
    int batch, n;
    uint32_t k;                 // already clamped to n by the caller; 0 => counts are all 0
    int has_min_conf;
    float min_conf;
    const uint8_t* range_state; // [n] 0 = absent (keep), 1 = keep * score, 2 = drop; or nullptr
    const float* range_score;   // [n]
    int rerank;
    Pred* out;                  // [B][k]
    uint32_t* out_count;        // [B]
    uint32_t* nonfinite;        // optional counter: += 1 per segment whose logits contain NaN / inf
};
cudaError_t launch_topk(const TopkParams& p, cudaStream_t stream);

// range filter applied to already-selected predictions (RangeFilter::filter_predictions,
// src/rangefilter.rs:333-386): in/out lists of `stride` slots per row, counts per row.
cudaError_t launch_range_filter(const Pred* in, const uint32_t* in_count, int rows, int stride,
                                const uint8_t* state, const float* score, int n, int rerank,
                                Pred* out, uint32_t* out_count, cudaStream_t stream);

}  // namespace bn
