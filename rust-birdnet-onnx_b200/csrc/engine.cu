// Engine: ONNX -> plan -> device weights; per-context staging slabs, streams and the
// enqueue of the whole hot path (H2D, front-end, CNN, epilogue, D2H).
//
// Reference call sites replaced: src/classifier.rs:340-383 (build), 504-574 (run_inference
// monitor), 676-727 (predict_batch), 826-867 (predict_batch_with_context);
// src/batch_context.rs:102-133, 188-338.
#include "engine.h"
#include "frontend_v24.h"
#include "hostcopy.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>

namespace bn {

// --------------------------------------------------------------------------------------
// error channel
// --------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local uint64_t g_detail[3] = {0, 0, 0};

int set_error(int status, const std::string& msg) {
    g_err = msg;
    g_detail[0] = g_detail[1] = g_detail[2] = 0;
    return status;
}
int set_error_detail(int status, const std::string& msg, uint64_t a, uint64_t b, uint64_t c) {
    g_err = msg;
    g_detail[0] = a; g_detail[1] = b; g_detail[2] = c;
    return status;
}
const std::string& last_error() { return g_err; }
const uint64_t* last_detail() { return g_detail; }

int cuda_fail(cudaError_t e, const char* what) {
    std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " (" + what + ")";
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice ||
        e == cudaErrorInitializationError)
        return set_error(BN_ERR_RUNTIME_INIT, msg);
    return set_error(BN_ERR_INFERENCE, msg);
}

RangeDev::~RangeDev() {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    if (state) cudaFree(state);
    if (score) cudaFree(score);
    cudaSetDevice(prev);
}

// --------------------------------------------------------------------------------------
// io info / detection glue
// --------------------------------------------------------------------------------------
static void copy_tensor_info(const std::string& name, const std::vector<int64_t>& dims, bn_tensor_info* ti) {
    memset(ti, 0, sizeof(*ti));
    strncpy(ti->name, name.c_str(), sizeof(ti->name) - 1);
    ti->rank = (int32_t)std::min<size_t>(dims.size(), BN_MAX_DIMS);
    for (int i = 0; i < ti->rank; ++i) ti->dims[i] = dims[i];
}

int fill_io_info(const Plan& plan, bn_io_info* out) {
    memset(out, 0, sizeof(*out));
    copy_tensor_info(plan.input_name, plan.input_dims, &out->input);
    out->n_outputs = (int32_t)std::min<size_t>(plan.outputs.size(), BN_MAX_OUTPUTS);
    for (int i = 0; i < out->n_outputs; ++i) copy_tensor_info(plan.outputs[i].name, plan.outputs[i].dims, &out->outputs[i]);
    out->model_type = plan.model_type;
    out->sample_rate = plan.model_type == MT_BIRDNET_V24 ? 48000u : 32000u;         // types.rs:17-22
    out->segment_duration = plan.model_type == MT_BIRDNET_V24 ? 3.0f : 5.0f;        // types.rs:26-31
    out->sample_count = (uint64_t)plan.sample_count;
    out->num_species = (uint64_t)plan.num_species;
    out->embedding_dim = (uint64_t)plan.embedding_dim;
    return BN_OK;
}

static int finish_plan(Plan& plan, int override_type) {
    std::vector<std::vector<int64_t>> oshapes;
    for (auto& o : plan.outputs) oshapes.push_back(o.dims);
    std::string reason;
    if (!detect_model_type(plan.input_dims, oshapes, override_type, &plan.model_type, &plan.sample_count,
                           &plan.num_species, &plan.embedding_dim, &reason))
        return set_error(BN_ERR_MODEL_DETECTION, reason);
    // output index map of src/classifier.rs:917-934
    int li = plan.model_type == MT_BIRDNET_V24 ? 0 : plan.model_type == MT_BIRDNET_V30 ? 1 : 3;
    plan.logits_tensor = plan.outputs[li].tensor;
    plan.embedding_tensor = plan.model_type == MT_BIRDNET_V24 ? -1 : plan.outputs[0].tensor;
    const TensorInfo& lt = plan.tensors[plan.logits_tensor];
    if ((int)lt.elems() != plan.num_species)
        return set_error(BN_ERR_MODEL_LOAD, "declared logits shape does not match the graph");
    if (plan.embedding_tensor >= 0 && (int)plan.tensors[plan.embedding_tensor].elems() != plan.embedding_dim)
        return set_error(BN_ERR_MODEL_LOAD, "declared embedding shape does not match the graph");
    if (plan.sample_count != plan.fe.sample_count)
        return set_error(BN_ERR_MODEL_LOAD, "input sample count does not match the front-end");
    return BN_OK;
}

int load_plan(const char* path, int override_type, Plan& plan) {
    if (!path) return set_error(BN_ERR_MODEL_PATH_REQUIRED, "model path required");
    try {
        OnnxModel m;
        load_onnx(path, m);
        build_plan(m, plan);
    } catch (const std::exception& ex) {
        return set_error(BN_ERR_MODEL_LOAD, ex.what());
    }
    return finish_plan(plan, override_type);
}

// --------------------------------------------------------------------------------------
// front-end basis: window[n] * cos(2*pi*f*n/N) projected on the mel matrix (float64 maths)
// --------------------------------------------------------------------------------------
static void build_real_mel_basis(const SpecBranch& br, std::vector<float>& basis, int& ldb) {
    const int N = br.n_fft, F = br.n_bins, Mm = br.n_mels;
    ldb = (Mm + 3) / 4 * 4;
    basis.assign((size_t)N * ldb, 0.f);
    std::vector<double> costab(N);
    for (int i = 0; i < N; ++i) costab[i] = cos(2.0 * M_PI * (double)i / (double)N);
    // sparse mel rows
    std::vector<std::vector<std::pair<int, double>>> nz(F);
    for (int f = 0; f < F; ++f)
        for (int m = 0; m < Mm; ++m) {
            float w = br.mel[(size_t)f * Mm + m];
            if (w != 0.f) nz[f].push_back({br.flip ? Mm - 1 - m : m, (double)w});
        }
    std::vector<double> acc(Mm);
    for (int n = 0; n < N; ++n) {
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int f = 0; f < F; ++f) {
            if (nz[f].empty()) continue;
            double c = costab[(int)(((long long)f * n) % N)];
            for (auto& pr : nz[f]) acc[pr.first] += c * pr.second;
        }
        for (int m = 0; m < Mm; ++m) basis[(size_t)n * ldb + m] = (float)(acc[m] * (double)br.window[n]);
    }
}

// MBConv block with squeeze-excite starting at op i (the 1x1 expand conv)?  ops i..i+5 = expand, depthwise, pool, FC silu,
// FC sigmoid, gated 1x1 projection; the expanded tensor and the depthwise output have no other reader.
static bool match_mbconv(const Plan& p, size_t i) {
    if (i + 5 >= p.ops.size()) return false;
    const PlanOp &ex = p.ops[i], &dw = p.ops[i + 1], &gp = p.ops[i + 2], &f1 = p.ops[i + 3], &f2 = p.ops[i + 4], &pr = p.ops[i + 5];
    if (ex.kind != OP_CONV || ex.k != 1 || ex.stride != 1 || ex.pad != 0 || ex.act != ACT_SILU || ex.in_scale >= 0 || ex.residual >= 0) return false;
    if (dw.kind != OP_DWCONV || dw.in != ex.out || dw.act != ACT_SILU || dw.pad != dw.k / 2 || dw.cout != ex.cout) return false;
    if (gp.kind != OP_GAP || gp.in != dw.out) return false;
    if (f1.kind != OP_LINEAR || f1.act != ACT_SILU || f1.in != gp.out || f1.cin != dw.cout || f1.in_scale >= 0 || f1.residual >= 0) return false;
    if (f2.kind != OP_LINEAR || f2.act != ACT_SIGMOID || f2.in != f1.out || f2.cout != dw.cout || f2.in_scale >= 0 || f2.residual >= 0) return false;
    if (pr.kind != OP_CONV || pr.k != 1 || pr.stride != 1 || pr.pad != 0 || pr.act != ACT_NONE || pr.in != dw.out || pr.in_scale != f2.out) return false;
    if (pr.residual >= 0 && (pr.residual != ex.in || pr.cout != ex.cin)) return false;
    // E, D, the pooled vector, the FC outputs: read by nothing else and not model outputs
    const int inner[5] = {ex.out, dw.out, gp.out, f1.out, f2.out};
    for (size_t q = 0; q < p.ops.size(); ++q) {
        if (q >= i && q <= i + 5) continue;
        for (int t : inner)
            if (p.ops[q].in == t || p.ops[q].residual == t || p.ops[q].in_scale == t) return false;
    }
    for (auto& o : p.outputs)
        for (int t : inner)
            if (p.root(o.tensor) == p.root(t)) return false;
    return mbconv_supported(dw.hin, dw.win, dw.k, dw.stride, ex.cin, ex.cout, pr.cout, f1.cout);
}

// UMMA N tile for a layer: the largest multiple of 16 (<= 128) that divides cout rounded up to 16
static int choose_nt(int cout) {
    const int c16 = (cout + 15) / 16 * 16;
    int best = 16;
    for (int nt = 16; nt <= 128; nt += 16)      // N = 2*NT per MMA (main | correction accumulators) must stay <= 256
        if (c16 % nt == 0) best = nt;
    return best;
}

// --------------------------------------------------------------------------------------
// engine
// --------------------------------------------------------------------------------------
int engine_create(const char* path, const bn_device_cfg* cfg, bn_engine** out) {
    if (!out) return set_error(BN_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    std::unique_ptr<bn_engine> e(new bn_engine());
    e->device = cfg ? cfg->device_id : 0;
    int ov = cfg ? cfg->model_type_override : BN_MODEL_AUTO;
    int st = load_plan(path, ov, e->plan);
    if (st != BN_OK) return st;
    fill_io_info(e->plan, &e->info);
    unsigned hw = std::thread::hardware_concurrency();
    e->pack_threads = cfg && cfg->pack_threads > 0 ? cfg->pack_threads : (int)std::max(1u, std::min(16u, hw / 2));

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_error(BN_ERR_RUNTIME_INIT, std::string("no CUDA device available: ") + cudaGetErrorString(ce));
    if (e->device < 0 || e->device >= ndev)
        return set_error(BN_ERR_RUNTIME_INIT, "device_id " + std::to_string(e->device) + " out of range (" + std::to_string(ndev) + " devices)");
    BN_CUDA(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    BN_CUDA(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major < 10)
        return set_error(BN_ERR_RUNTIME_INIT, std::string("device '") + prop.name + "' is not sm_100-class; this engine is built for sm_100a only");
    BN_CUDA(init_kernels_for_device());
    BN_CUDA(tc_conv_init_device());
    BN_CUDA(dw_se_init_device());
    BN_CUDA(mbconv_init_device());
    const char* env_tc = getenv("BN_DISABLE_TC");
    const bool tc_enabled = !(env_tc && env_tc[0] == '1');
    e->tc_mode = tc_enabled;
    { const char* ev = getenv("BN_DISABLE_TMA"); e->use_tma = !(ev && ev[0] == '1'); }
    e->num_sms = prop.multiProcessorCount;
    const char* env_st = getenv("BN_TC_STAGES");
    const int forced_stages = env_st ? atoi(env_st) : 0;

    // measured (profiles/r01_stage_times_b256.txt): wide 3x3 tiles (N >= 64) run faster with 8 epilogue warps and a
    // deeper operand ring, everything else with 16
    const char* env_epi = getenv("BN_EPI_WARPS");
    auto epi_rule = [&](int k, int nt) -> int {
        if (env_epi) return atoi(env_epi);
        return (k == 3 && nt >= 64) ? 8 : 16;
    };
    Plan& p = e->plan;
    e->dev_ops.resize(p.ops.size());
    for (size_t i = 0; i < p.ops.size(); ++i) {
        PlanOp& op = p.ops[i];
        if (op.kind == OP_GAP) continue;
        BN_CUDA(cudaMalloc(&e->dev_ops[i].weight, op.weight.size() * sizeof(float)));
        BN_CUDA(cudaMemcpy(e->dev_ops[i].weight, op.weight.data(), op.weight.size() * sizeof(float), cudaMemcpyHostToDevice));
        BN_CUDA(cudaMalloc(&e->dev_ops[i].bias, op.bias.size() * sizeof(float)));
        BN_CUDA(cudaMemcpy(e->dev_ops[i].bias, op.bias.data(), op.bias.size() * sizeof(float), cudaMemcpyHostToDevice));
        // dense contractions with >= 64 output channels (or any spatial conv) go to the tensor cores
        const bool dense = op.kind == OP_CONV || (op.kind == OP_LINEAR && op.cout >= 64);
        if (tc_enabled && dense && (op.cin % 8) == 0 && op.cout >= 16) {
            DevOp& d = e->dev_ops[i];
            const int K = op.k * op.k * op.cin;
            d.nt = choose_nt(op.cout);
            // 3x3 stride-1 layers on wide images: halo mode if some N tile lets the weights stay resident
            const char* env_halo = getenv("BN_DISABLE_HALO");
            if (!(env_halo && env_halo[0] == '1') && op.kind == OP_CONV && op.in_scale < 0) {
                const int c16 = (op.cout + 15) / 16 * 16;
                static const int halo_nt_max = [] { const char* ev = getenv("BN_HALO_NT_MAX"); return ev ? atoi(ev) : 128; }();
                for (int nt = std::min(d.nt, halo_nt_max); nt >= 16; nt -= 16) {
                    if (c16 % nt) continue;
                    const int slots = tc_conv_halo_slots(op.k, op.stride, op.pad, op.cin, op.wout, op.win, nt, (K + 63) / 64, epi_rule(op.k, nt));
                    if (slots >= 2) { d.nt = nt; d.halo_slots = slots; break; }
                }
            }
            d.epi_warps = epi_rule(op.k, d.nt);
            std::vector<uint16_t> pack;
            tc_pack_weights(op.weight.data(), K, op.cout, op.ldw, d.nt, pack, &d.n_tiles, &d.k_chunks);
            d.stages = forced_stages > 1 ? forced_stages : tc_conv_pick_stages(d.nt, d.k_chunks, d.epi_warps);
            d.tmem_cols = 32;
            // 2 buffers x (main | correction); narrow tiles with 16 epilogue warps get four buffers (k_tc_conv: n_acc)
            { const char* ev = getenv("BN_TC_ACC4"); const bool acc4 = !(ev && ev[0] == '0') && d.nt <= 32 && d.epi_warps == 16;
              while (d.tmem_cols < (acc4 ? 8 : 4) * d.nt) d.tmem_cols <<= 1; }
            BN_CUDA(cudaMalloc(&d.wpack, pack.size() * sizeof(uint16_t)));
            BN_CUDA(cudaMemcpy(d.wpack, pack.data(), pack.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
            d.use_tc = true;
        }
        // fused MBConv block: expand weights packed per channel group, projection weights output-channel major
        static const bool no_mbconv = [] { const char* ev = getenv("BN_DISABLE_MBCONV"); return ev && ev[0] == '1'; }();
        if (tc_enabled && !no_mbconv && e->dev_ops[i].use_tc && match_mbconv(p, i)) {
            DevOp& d = e->dev_ops[i];
            const PlanOp &dw = p.ops[i + 1], &pr = p.ops[i + 5];
            d.mb_group = mbconv_group(dw.hin, dw.win, dw.k, op.cin, op.cout, pr.cout, p.ops[i + 3].cout);
            std::vector<uint16_t> pack;
            int nt_ = 0, kc_ = 0;
            tc_pack_weights(op.weight.data(), op.cin, op.cout, op.ldw, d.mb_group, pack, &nt_, &kc_);
            BN_CUDA(cudaMalloc(&d.mb_we_pack, pack.size() * sizeof(uint16_t)));
            BN_CUDA(cudaMemcpy(d.mb_we_pack, pack.data(), pack.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
            std::vector<float> wt((size_t)pr.cout * pr.cin);
            for (int c = 0; c < pr.cin; ++c)
                for (int n = 0; n < pr.cout; ++n) wt[(size_t)n * pr.cin + c] = pr.weight[(size_t)c * pr.ldw + n];
            BN_CUDA(cudaMalloc(&d.mb_wpT, wt.size() * sizeof(float)));
            BN_CUDA(cudaMemcpy(d.mb_wpT, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
            std::vector<uint16_t> ppack;                      // one N tile = all output channels: [k_chunk][hi|lo][cout x 64]
            tc_pack_weights(pr.weight.data(), pr.cin, pr.cout, pr.ldw, pr.cout, ppack, &nt_, &kc_);
            BN_CUDA(cudaMalloc(&d.mb_wp_pack, ppack.size() * sizeof(uint16_t)));
            BN_CUDA(cudaMemcpy(d.mb_wp_pack, ppack.data(), ppack.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        }
        std::vector<float>().swap(op.weight);   // host copy no longer needed
    }
    if (p.fe.kind == FE_BIRDNET_V24) {
        for (auto& br : p.fe.branches) {
            std::vector<float> basis;
            int ldb = 0;
            build_real_mel_basis(br, basis, ldb);
            float* d = nullptr;
            BN_CUDA(cudaMalloc(&d, basis.size() * sizeof(float)));
            BN_CUDA(cudaMemcpy(d, basis.data(), basis.size() * sizeof(float), cudaMemcpyHostToDevice));
            e->d_basis.push_back(d);
            e->ldb.push_back(ldb);
            if (e->tc_mode) {
                // A[t][k'] = Xp[t*row_stride + k'], k' = j*row_stride + c  <->  n = j*hop + c of the frame
                bn_engine::FeTc ft;
                ft.hop = br.hop;
                ft.row_stride = (br.hop + 7) / 8 * 8;
                const int J = ft.row_stride == br.hop ? 0 : (br.n_fft + br.hop - 1) / br.hop;
                ft.K = J ? J * ft.row_stride : br.n_fft;
                ft.rows = J ? br.n_frames + J - 1 : (p.sample_count + ft.row_stride - 1) / ft.row_stride + 1;
                std::vector<float> wb((size_t)ft.K * ldb, 0.f);
                for (int kp = 0; kp < ft.K; ++kp) {
                    int n = kp;
                    if (J) {
                        const int j = kp / ft.row_stride, cc = kp - j * ft.row_stride;
                        n = cc < br.hop ? j * br.hop + cc : -1;
                    }
                    if (n >= 0 && n < br.n_fft)
                        for (int m = 0; m < br.n_mels; ++m) wb[(size_t)kp * ldb + m] = basis[(size_t)n * ldb + m];
                }
                ft.nt = choose_nt(br.n_mels);
                std::vector<uint16_t> pack;
                tc_pack_weights(wb.data(), ft.K, br.n_mels, ldb, ft.nt, pack, &ft.n_tiles, &ft.k_chunks);
                ft.stages = tc_conv_pick_stages(ft.nt, ft.k_chunks);
                ft.tmem_cols = 32;
                while (ft.tmem_cols < 4 * ft.nt) ft.tmem_cols <<= 1;
                BN_CUDA(cudaMalloc(&ft.wpack, pack.size() * sizeof(uint16_t)));
                BN_CUDA(cudaMemcpy(ft.wpack, pack.data(), pack.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                e->fe_tc.push_back(ft);
            }
        }
        // one kernel for normaliser + both spectrogram GEMMs (BN_DISABLE_FE_FUSED=1 keeps the frame-matrix path for A/B runs)
        const char* nofe = getenv("BN_DISABLE_FE_FUSED");
        const size_t nb = p.fe.branches.size();
        if (e->tc_mode && p.fe.normalize && !(nofe && nofe[0] == '1') && nb == 2) {      // the two-branch form is the one the GPU tests cover
            SpecBranchHost hb[2];
            bool ok = true;
            for (size_t bi = 0; bi < nb && ok; ++bi) {
                const SpecBranch& br = p.fe.branches[bi];
                ok = spec_v24_plan(br.n_fft, br.hop, br.n_mels, hb[bi]) && br.n_frames == p.fe.branches[0].n_frames &&
                     br.n_mels == p.fe.branches[0].n_mels;
            }
            auto& fv = e->fe_v24;
            int order[2] = {0, 1};
            if (ok && nb == 2 && hb[1].table.size() > hb[0].table.size()) { order[0] = 1; order[1] = 0; std::swap(hb[0], hb[1]); }
            ok = ok && spec_v24_layout(hb, (int)nb, fv.row_pitch, fv.patch_plane, fv.n_stages, fv.smem_bytes);
            if (ok) {
                BN_CUDA(spec_v24_init_device());
                for (size_t sl = 0; sl < nb; ++sl) {
                    const int bi = order[sl];
                    const SpecBranch& br = p.fe.branches[bi];
                    std::vector<float> basis;
                    int ldb = 0;
                    build_real_mel_basis(br, basis, ldb);
                    std::vector<uint16_t> pack;
                    spec_v24_pack(hb[sl], basis.data(), ldb, br.n_fft, br.n_mels, pack);
                    BN_CUDA(cudaMalloc(&fv.wpack[sl], pack.size() * sizeof(uint16_t)));
                    BN_CUDA(cudaMemcpy(fv.wpack[sl], pack.data(), pack.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    fv.slot_branch[sl] = bi;
                    fv.hop[sl] = hb[sl].hop; fv.kcells[sl] = hb[sl].kcells; fv.rows[sl] = hb[sl].rows;
                    fv.blocks[sl] = hb[sl].blocks; fv.split[sl] = hb[sl].split; fv.pad[sl] = hb[sl].pad;
                    fv.n_ksteps[sl] = (int)hb[sl].table.size();
                }
                fv.n_br = (int)nb;
                fv.n_pad = hb[0].n_pad;
                fv.on = true;
            }
        }
    } else {
        if (!tc_enabled) return set_error(BN_ERR_MODEL_LOAD, "the log-mel front-end needs the tensor-core (planes) path; unset BN_DISABLE_TC");
        const SpecBranch& br = p.fe.branches[0];
        auto& lm = e->fe_lm;
        if (br.n_bins != br.n_fft / 2 + 1) return set_error(BN_ERR_MODEL_LOAD, "log-mel front-end: mel matrix rows != n_fft/2 + 1");
        if (!logmel_factorize(br.n_fft, lm.radix, &lm.n_stages) || logmel_smem_bytes(br.n_fft) > 96 * 1024)
            return set_error(BN_ERR_MODEL_LOAD, "log-mel front-end: unsupported STFT length " + std::to_string(br.n_fft));
        BN_CUDA(logmel_init_device());
        std::vector<float> tw;
        logmel_twiddles(br.n_fft, tw);
        // banded form of the mel matrix: filter m = bins [lo, lo + cnt) (zeros inside the band are kept)
        std::vector<int> lo(br.n_mels, 0), cnt(br.n_mels, 0), off(br.n_mels, 0);
        std::vector<float> wv;
        for (int m = 0; m < br.n_mels; ++m) {
            int first = -1, last = -1;
            for (int f = 0; f < br.n_bins; ++f)
                if (br.mel[(size_t)f * br.n_mels + m] != 0.f) { if (first < 0) first = f; last = f; }
            off[m] = (int)wv.size();
            if (first >= 0) {
                lo[m] = first; cnt[m] = last - first + 1;
                for (int f = first; f <= last; ++f) wv.push_back(br.mel[(size_t)f * br.n_mels + m]);
            }
        }
        if (wv.empty()) wv.push_back(0.f);
        auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
            cudaError_t ce = cudaMalloc(dst, bytes);
            if (ce != cudaSuccess) return ce;
            return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
        };
        BN_CUDA(up(br.window.data(), br.window.size() * sizeof(float), (void**)&lm.window));
        BN_CUDA(up(tw.data(), tw.size() * sizeof(float), (void**)&lm.twiddle));
        BN_CUDA(up(lo.data(), lo.size() * sizeof(int), (void**)&lm.mel_lo));
        BN_CUDA(up(cnt.data(), cnt.size() * sizeof(int), (void**)&lm.mel_cnt));
        BN_CUDA(up(off.data(), off.size() * sizeof(int), (void**)&lm.mel_off));
        BN_CUDA(up(wv.data(), wv.size() * sizeof(float), (void**)&lm.mel_w));
    }
    { const char* ev = getenv("BN_COMPUTE_LANES"); if (ev && atoi(ev) >= 1 && atoi(ev) <= bn_engine::MAX_LANES) e->n_lanes = atoi(ev); }
    for (int l = 0; l < e->n_lanes; ++l) BN_CUDA(cudaStreamCreateWithFlags(&e->compute[l], cudaStreamNonBlocking));
    BN_CUDA(cudaStreamCreateWithFlags(&e->h2d, cudaStreamNonBlocking));
    *out = e.release();
    return BN_OK;
}

}  // namespace bn

bn_engine::~bn_engine() {
    cudaSetDevice(device);
    for (auto* c : run_free) delete c;       // every lease has been returned by the time the engine is destroyed
    for (auto& d : dev_ops) {
        if (d.weight) cudaFree(d.weight);
        if (d.bias) cudaFree(d.bias);
        if (d.wpack) cudaFree(d.wpack);
        if (d.mb_we_pack) cudaFree(d.mb_we_pack);
        if (d.mb_wp_pack) cudaFree(d.mb_wp_pack);
        if (d.mb_wpT) cudaFree(d.mb_wpT);
    }
    for (auto* b : d_basis) cudaFree(b);
    for (auto& f : fe_tc) if (f.wpack) cudaFree(f.wpack);
    for (int i = 0; i < 2; ++i) if (fe_v24.wpack[i]) cudaFree(fe_v24.wpack[i]);
    if (fe_lm.window) cudaFree(fe_lm.window);
    if (fe_lm.twiddle) cudaFree(fe_lm.twiddle);
    if (fe_lm.mel_lo) cudaFree(fe_lm.mel_lo);
    if (fe_lm.mel_cnt) cudaFree(fe_lm.mel_cnt);
    if (fe_lm.mel_off) cudaFree(fe_lm.mel_off);
    if (fe_lm.mel_w) cudaFree(fe_lm.mel_w);
    for (auto& l : compute) if (l) cudaStreamDestroy(l);
    if (h2d) cudaStreamDestroy(h2d);
}

bn_ctx::~bn_ctx() {
    if (!eng) return;
    // everything below works from this context's own copies: a context destroyed after its engine (interpreter
    // shutdown can do that to a binding) only gets CUDA errors back for the engine-owned lane, never a stale pointer
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (in_stream && in_stream != stream) cudaStreamSynchronize(in_stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    if (h_in) cudaFreeHost(h_in);
    if (h_pcm) cudaFreeHost(h_pcm);
    if (d_pcm) cudaFree(d_pcm);
    if (d_in) cudaFree(d_in);
    if (d_norm) cudaFree(d_norm);
    if (d_minmax) cudaFree(d_minmax);
    for (auto* x : d_xp) if (x) cudaFree(x);
    for (size_t i = 0; i < d_tensor.size(); ++i)
        if (d_tensor[i] && i < owns_tensor.size() && owns_tensor[i]) cudaFree(d_tensor[i]);
    if (h_logits) cudaFreeHost(h_logits);
    if (h_emb) cudaFreeHost(h_emb);
    if (d_topk) cudaFree(d_topk);
    if (d_count) cudaFree(d_count);
    if (h_topk) cudaFreeHost(h_topk);
    if (h_count) cudaFreeHost(h_count);
    for (auto ev : prof_events) cudaEventDestroy(ev);
    if (done) cudaEventDestroy(done);
    if (in_stream && in_stream != stream) cudaStreamDestroy(in_stream);
    if (stream && owns_stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_results) cudaEventDestroy(ev_results);
    if (ev_fetched) cudaEventDestroy(ev_fetched);
}

namespace bn {

int ctx_create(bn_engine* e, uint64_t max_batch, bn_ctx** out) {
    if (!e || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    BN_CUDA(cudaSetDevice(e->device));
    std::unique_ptr<bn_ctx> c(new bn_ctx());
    c->eng = e;
    c->device = e->device;
    c->max_batch = max_batch;
    const Plan& p = e->plan;
    const size_t S = (size_t)p.sample_count;
    const size_t mb = std::max<uint64_t>(max_batch, 1);
    static const bool shared_lane = [] { const char* ev = getenv("BN_SHARED_COMPUTE"); return !(ev && ev[0] == '0'); }();
    if (shared_lane && e->compute[0]) {
        c->lane = e->next_lane.fetch_add(1) % e->n_lanes;
        c->stream = e->compute[c->lane];
        c->owns_stream = false;
        BN_CUDA(cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking));
    } else {
        BN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->in_stream = c->stream;
    }
    BN_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    BN_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    // blocking sync: several driver threads per GPU (and 8 ranks per box) wait here; they must sleep, not spin on the host cores
    BN_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming | cudaEventBlockingSync));
    BN_CUDA(cudaEventCreateWithFlags(&c->ev_results, cudaEventDisableTiming));
    BN_CUDA(cudaEventCreateWithFlags(&c->ev_fetched, cudaEventDisableTiming));
    BN_CUDA(cudaHostAlloc(&c->h_in, mb * S * sizeof(float), cudaHostAllocDefault));
    memset(c->h_in, 0, mb * S * sizeof(float));    // vec![0.0f32; max*sample_count], batch_context.rs:122
    BN_CUDA(cudaMalloc(&c->d_in, mb * S * sizeof(float)));
    if (p.fe.normalize) BN_CUDA(cudaMalloc(&c->d_norm, mb * S * sizeof(float)));
    BN_CUDA(cudaMalloc(&c->d_minmax, mb * 2 * sizeof(uint32_t)));
    { const char* kn = getenv("BN_KEEP_NORMALIZED"); c->keep_normalized = !e->tc_mode || (kn && kn[0] == '1'); }
    if (!e->fe_v24.on)
    for (auto& ft : e->fe_tc) {
        __half* x = nullptr;
        BN_CUDA(cudaMalloc(&x, 2 * mb * (size_t)ft.rows * ft.row_stride * sizeof(__half)));
        c->d_xp.push_back(x);
    }
    c->d_tensor.assign(p.tensors.size(), nullptr);
    c->owns_tensor.assign(p.tensors.size(), 0);
    for (size_t i = 0; i < p.tensors.size(); ++i) {
        const TensorInfo& t = p.tensors[i];
        if (t.alias_of >= 0 || t.scale_base >= 0) continue;
        BN_CUDA(cudaMalloc(&c->d_tensor[i], mb * t.elems() * sizeof(float)));
        c->owns_tensor[i] = 1;
    }
    for (size_t i = 0; i < p.tensors.size(); ++i)
        if (p.tensors[i].alias_of >= 0) c->d_tensor[i] = c->d_tensor[p.root((int)i)];
    BN_CUDA(cudaHostAlloc(&c->h_logits, mb * p.num_species * sizeof(float), cudaHostAllocDefault));
    if (p.embedding_dim > 0) BN_CUDA(cudaHostAlloc(&c->h_emb, mb * p.embedding_dim * sizeof(float), cudaHostAllocDefault));
    BN_CUDA(cudaMalloc(&c->d_count, (mb + 1) * sizeof(uint32_t)));          // [mb] per-segment counts + the non-finite counter
    BN_CUDA(cudaHostAlloc(&c->h_count, (mb + 1) * sizeof(uint32_t), cudaHostAllocDefault));
    c->h_count[mb] = 0;
    *out = c.release();
    return BN_OK;
}

static int ensure_topk_capacity(bn_ctx* c, uint64_t k) {
    if (k <= c->topk_cap && c->d_topk) return BN_OK;
    const size_t mb = std::max<uint64_t>(c->max_batch, 1);
    if (c->d_topk) { cudaFree(c->d_topk); c->d_topk = nullptr; }
    if (c->h_topk) { cudaFreeHost(c->h_topk); c->h_topk = nullptr; }
    uint64_t cap = std::max<uint64_t>(k, 1);
    BN_CUDA(cudaMalloc(&c->d_topk, mb * cap * sizeof(Pred)));
    BN_CUDA(cudaHostAlloc(&c->h_topk, mb * cap * sizeof(Pred), cudaHostAllocDefault));
    c->topk_cap = cap;
    return BN_OK;
}

// ---- profiling helpers -------------------------------------------------------------------
static void prof_mark(bn_ctx* c, const char* name) {
    if (!c->profiling) return;
    size_t i = c->prof_names.size();
    if (i >= c->prof_events.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        c->prof_events.push_back(ev);
    }
    cudaEventRecord(c->prof_events[i], c->stream);
    c->prof_names.push_back(name);
}


// ---- tensor-core mode: spatial tensors are hi/lo fp16 planes, vectors ([B][C]) stay FP32 ------
static inline bool is_spatial(const Plan& p, int t) { return p.tensors[t].H * p.tensors[t].W > 1; }
static inline PlanesPtr planes_of(bn_ctx* c, int t) {
    const Plan& p = c->eng->plan;
    PlanesPtr r;
    r.hi = reinterpret_cast<__half*>(c->d_tensor[t]);
    r.plane = (size_t)std::max<uint64_t>(c->max_batch, 1) * p.tensors[p.root(t)].elems();
    return r;
}

static void prof_mark(bn_ctx* c, const char* name);
static int wait_for_fetch(bn_ctx* c);

// squeeze-excite tail starting at op i?  (FC silu -> FC sigmoid -> conv gated by it, input = the
// tensor whose pool feeds the first FC).  Returns the index of the gated conv or -1.
static int match_se_tail(const Plan& p, size_t i) {
    if (i + 2 >= p.ops.size()) return -1;
    const PlanOp &f1 = p.ops[i], &f2 = p.ops[i + 1];
    if (f1.kind != OP_LINEAR || f1.act != ACT_SILU || f2.kind != OP_LINEAR || f2.act != ACT_SIGMOID) return -1;
    if (f2.in != f1.out || f1.in_scale >= 0 || f2.in_scale >= 0 || f1.residual >= 0 || f2.residual >= 0) return -1;
    // the pooled vector must come from a GAP op
    int gap = -1;
    for (size_t j = 0; j < i; ++j)
        if (p.ops[j].kind == OP_GAP && p.ops[j].out == f1.in) gap = (int)j;
    if (gap < 0) return -1;
    const int d = p.ops[gap].in;
    for (size_t j = i + 2; j < p.ops.size() && j < i + 4; ++j) {
        const PlanOp& cv = p.ops[j];
        if (cv.in_scale == f2.out && cv.in == d && (cv.kind == OP_CONV)) {
            // d may only be read by its pool and by this conv (it is rescaled in place)
            for (size_t q = 0; q < p.ops.size(); ++q) {
                if ((int)q == gap || q == j) continue;
                if (p.ops[q].in == d || p.ops[q].residual == d) return -1;
            }
            for (auto& o : p.outputs) if (p.root(o.tensor) == p.root(d)) return -1;
            return (int)j;
        }
    }
    return -1;
}

static bool no_dw_se_flag() {
    static const bool v = [] { const char* ev = getenv("BN_DISABLE_DW_SE"); return ev && ev[0] == '1'; }();
    return v;
}

static int enqueue_ops_tc(bn_ctx* c, int B, uint64_t& launches) {
    bn_engine* e = c->eng;
    const Plan& p = e->plan;
    cudaStream_t s = c->stream;
    int prescaled_conv = -1;                   // conv whose gated input was already rescaled in place
    // FP32 hand-off: a tensor-core conv whose output is read by exactly one op, a depthwise conv that runs in the
    // fused dw + squeeze-excite kernel, writes plain FP32 (dw_se.cu lands it in shared memory with cp.async)
    static const bool no_f32_handoff = [] { const char* ev = getenv("BN_DISABLE_F32_HANDOFF"); return ev && ev[0] == '1'; }();
    auto sole_dw_consumer = [&](int t) -> int {
        int found = -1, uses = 0;
        for (size_t q = 0; q < p.ops.size(); ++q) {
            const PlanOp& o = p.ops[q];
            if (o.in == t) { ++uses; if (o.kind == OP_DWCONV) found = (int)q; }
            if (o.residual == t || o.in_scale == t) ++uses;
        }
        for (auto& o : p.outputs) if (p.root(o.tensor) == p.root(t)) ++uses;
        return uses == 1 ? found : -1;
    };
    auto producer_of = [&](int t) -> int {
        for (size_t q = 0; q < p.ops.size(); ++q) if (p.ops[q].out == t) return (int)q;
        return -1;
    };
    // would the depthwise conv at `i` run fused, and with which input form?  fills sp on success
    auto dw_se_plan = [&](size_t i, DwSeParams& sp) -> int {
        const PlanOp& op = p.ops[i];
        if (op.kind != OP_DWCONV || no_dw_se_flag() || i + 3 >= p.ops.size()) return -1;
        if (!(p.ops[i + 1].kind == OP_GAP && p.ops[i + 1].in == op.out)) return -1;
        const int cv = match_se_tail(p, i + 2);
        if (cv < 0 || p.ops[cv].in != op.out || p.ops[i + 2].in != p.ops[i + 1].out) return -1;
        const PlanOp &f1 = p.ops[i + 2], &f2 = p.ops[i + 3];
        if (f2.cout != op.cout || f1.cin != op.cout) return -1;
        sp = DwSeParams{};
        sp.in = planes_of(c, op.in); sp.weight = e->dev_ops[i].weight; sp.bias = e->dev_ops[i].bias; sp.out = planes_of(c, op.out);
        sp.w1 = e->dev_ops[i + 2].weight; sp.b1 = e->dev_ops[i + 2].bias; sp.ldw1 = f1.ldw;
        sp.w2 = e->dev_ops[i + 3].weight; sp.b2 = e->dev_ops[i + 3].bias; sp.ldw2 = f2.ldw;
        sp.pooled_out = c->d_tensor[p.ops[i + 1].out]; sp.gate_out = c->d_tensor[f2.out];
        sp.batch = B; sp.hin = op.hin; sp.win = op.win; sp.c = op.cout; sp.hout = op.hout; sp.wout = op.wout;
        sp.k = op.k; sp.stride = op.stride; sp.pad = op.pad; sp.act = op.act; sp.r = f1.cout;
        // FP32 input when the producer is a tensor-core conv with a vectorisable epilogue and this is its only reader
        const int prod = producer_of(op.in);
        if (!no_f32_handoff && prod >= 0 && e->dev_ops[prod].use_tc && p.ops[prod].kind == OP_CONV && (p.ops[prod].cout & 15) == 0 &&
            is_spatial(p, op.in) && sole_dw_consumer(op.in) == (int)i) {
            sp.in_f32 = c->d_tensor[op.in];
            if (dw_se_supported(sp)) return cv;
            sp.in_f32 = nullptr;
        }
        return dw_se_supported(sp) ? cv : -1;
    };
    int fused_se_fc = -1;                      // first FC of a squeeze-excite tail that ran inside the depthwise kernel
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const PlanOp& op = p.ops[i];
        const DevOp& d = e->dev_ops[i];
        // the previous run's results may still be on their way to the host: wait before overwriting them
        if (c->fetch_pending && op.out >= 0 &&
            (p.root(op.out) == p.root(p.logits_tensor) || (p.embedding_tensor >= 0 && p.root(op.out) == p.root(p.embedding_tensor)))) {
            const int ws = wait_for_fetch(c);
            if (ws != BN_OK) return ws;
        }
        if (d.mb_group > 0) {
            // expand -> depthwise -> squeeze-excite -> projection in one kernel (mbconv.cu): E never leaves the SM
            const PlanOp &dw = p.ops[i + 1], &f1 = p.ops[i + 3], &f2 = p.ops[i + 4], &pr = p.ops[i + 5];
            if (c->mb_state.empty()) { c->mb_xmaps.resize(p.ops.size()); c->mb_dmaps.resize(p.ops.size()); c->mb_state.assign(p.ops.size(), 0); }
            const PlanesPtr xp = planes_of(c, op.in), dp = planes_of(c, dw.out);
            if (c->mb_state[i] == 0) {
                const uint64_t rows = (uint64_t)std::max<uint64_t>(c->max_batch, 1) * dw.hin * dw.win;
                const uint32_t box[5] = {64, 64, 1, 1, 1};
                const uint32_t es[5] = {1, 1, 1, 1, 1};
                const uint64_t dx[5] = {(uint64_t)op.cin, rows, 1, 1, 2};
                const uint64_t sx[4] = {(uint64_t)op.cin * 2, rows * op.cin * 2, rows * op.cin * 2, (uint64_t)xp.plane * 2};
                const uint64_t dd[5] = {(uint64_t)dw.cout, rows, 1, 1, 2};
                const uint64_t sd[4] = {(uint64_t)dw.cout * 2, rows * dw.cout * 2, rows * dw.cout * 2, (uint64_t)dp.plane * 2};
                const bool ok = tc_encode_tmap(&c->mb_xmaps[i], xp.hi, dx, sx, box, es, 64) && tc_encode_tmap(&c->mb_dmaps[i], dp.hi, dd, sd, box, es, 64);
                c->mb_state[i] = ok ? 1 : 2;
            }
            if (c->mb_state[i] == 1) {
                MbconvParams mp{};
                mp.xmap = c->mb_xmaps[i]; mp.dmap = c->mb_dmaps[i];
                mp.we_pack = d.mb_we_pack; mp.be = d.bias;
                mp.wd = e->dev_ops[i + 1].weight; mp.bd = e->dev_ops[i + 1].bias;
                mp.w1 = e->dev_ops[i + 3].weight; mp.b1 = e->dev_ops[i + 3].bias; mp.ldw1 = f1.ldw;
                mp.w2 = e->dev_ops[i + 4].weight; mp.b2 = e->dev_ops[i + 4].bias; mp.ldw2 = f2.ldw;
                mp.wp_pack = d.mb_wp_pack; mp.wpT = d.mb_wpT; mp.bp = e->dev_ops[i + 5].bias;
                mp.d_hi = dp.hi; mp.d_plane = dp.plane;
                if (pr.residual >= 0) { const PlanesPtr rp = planes_of(c, pr.residual); mp.res_hi = rp.hi; mp.res_plane = rp.plane; }
                const PlanesPtr op_ = planes_of(c, pr.out);
                mp.out_hi = op_.hi; mp.out_plane = op_.plane;
                mp.prof = c->profiling && getenv("BN_MB_PROFILE") ? tc_conv_prof_slot((int)i) : nullptr;
                { static const int dbg = [] { const char* ev = getenv("BN_MB_DEBUG"); return ev ? atoi(ev) : 0; }(); mp.debug = dbg; }
                mp.batch = B; mp.h = dw.hin; mp.w = dw.win; mp.k = dw.k; mp.cin = op.cin; mp.cexp = op.cout; mp.cout = pr.cout; mp.r = f1.cout;
                std::string nm = op.name;
                const size_t dot = nm.find(".expand");
                if (dot != std::string::npos) nm = nm.substr(0, dot);
                prof_mark(c, (nm + ".mbconv").c_str());
                BN_CUDA(launch_mbconv(mp, e->num_sms, s));
                ++launches;
                i += 5;
                continue;
            }
        }
        if (op.kind == OP_LINEAR && (int)i == fused_se_fc) {
            prescaled_conv = match_se_tail(p, i);
            ++i;                                 // both FCs ran inside launch_dw_se
            continue;
        }
        if (op.kind == OP_LINEAR) {
            const int cv = match_se_tail(p, i);
            if (cv >= 0 && (p.tensors[p.ops[cv].in].C % 8) == 0) {
                const PlanOp& f2 = p.ops[i + 1];
                const TensorInfo& dt = p.tensors[p.ops[cv].in];
                prof_mark(c, op.name.c_str());
                if (op.cout <= 256) {                        // gate kernel + streaming rescale
                    BN_CUDA(launch_se_gate(c->d_tensor[op.in], d.weight, d.bias, e->dev_ops[i + 1].weight, e->dev_ops[i + 1].bias,
                                           c->d_tensor[f2.out], B, dt.C, op.cout, op.ldw, f2.ldw, s));
                    BN_CUDA(launch_se_rescale(planes_of(c, p.ops[cv].in), c->d_tensor[f2.out], B, dt.H * dt.W, dt.C, s));
                    launches += 2;
                    prescaled_conv = cv;
                    ++i;
                    continue;
                }
                SeParams sp{};
                sp.pooled = c->d_tensor[op.in];
                sp.w1 = d.weight; sp.b1 = d.bias; sp.ldw1 = op.ldw;
                sp.w2 = e->dev_ops[i + 1].weight; sp.b2 = e->dev_ops[i + 1].bias; sp.ldw2 = f2.ldw;
                sp.gate_out = c->d_tensor[f2.out];
                sp.d = planes_of(c, p.ops[cv].in);
                sp.c = dt.C; sp.r = op.cout; sp.npix = dt.H * dt.W;
                BN_CUDA(launch_se_scale(sp, B, s));
                ++launches;
                prescaled_conv = cv;
                ++i;                             // the second FC ran inside the fused kernel
                continue;
            }
        }
        if (op.kind == OP_GAP) {
            // squeeze fused into the preceding depthwise conv?
            if (i > 0 && p.ops[i - 1].kind == OP_DWCONV && p.ops[i - 1].out == op.in) continue;
            prof_mark(c, op.name.c_str());
            BN_CUDA(launch_gap_planes(planes_of(c, op.in), c->d_tensor[op.out], B, op.hin * op.win, op.cin, s));
            ++launches;
            continue;
        }
        prof_mark(c, op.name.c_str());
        if (op.kind == OP_DWCONV) {
            float* pooled = nullptr;
            if (i + 1 < p.ops.size() && p.ops[i + 1].kind == OP_GAP && p.ops[i + 1].in == op.out) pooled = c->d_tensor[p.ops[i + 1].out];
            // dw -> pool -> FC silu -> FC sigmoid -> gated conv: one kernel (dw_se.cu)
            {
                DwSeParams sp;
                if (dw_se_plan(i, sp) >= 0) {
                    BN_CUDA(launch_dw_se(sp, s));
                    ++launches;
                    fused_se_fc = (int)i + 2;
                    continue;
                }
            }
            DwPlanesParams dp{planes_of(c, op.in), d.weight, d.bias, planes_of(c, op.out), pooled,
                              B, op.hin, op.win, op.cout, op.hout, op.wout, op.k, op.stride, op.pad, op.act};
            BN_CUDA(launch_dwconv_planes(dp, s));
        } else if (d.use_tc) {
            TcConvParams tp{};
            const bool in_sp = is_spatial(p, op.in);
            if (in_sp) {
                PlanesPtr ip = planes_of(c, op.in);
                tp.in_hi = ip.hi; tp.in_plane = ip.plane;
                const bool gated = op.in_scale >= 0 && (int)i != prescaled_conv;
                tp.in_mode = gated ? TC_IN_PLANES_SCALED : (d.halo_slots > 0 ? TC_IN_HALO : TC_IN_PLANES);
                tp.in_scale = gated ? c->d_tensor[op.in_scale] : nullptr;
                // 1x1 stride-1 layers: the A operand is a plain [rows][cin] matrix per plane -> one TMA box per
                // plane and K chunk lands it in the SWIZZLE_128B layout, no per-thread cp.async address work
                if (!gated && e->use_tma && op.k == 1 && op.stride == 1 && op.pad == 0 && op.cin <= 256) {   // measured: deep-K projections are faster with the cp.async gather
                    if (c->tmap_state.empty()) { c->tmaps.resize(p.ops.size()); c->tmap_state.assign(p.ops.size(), 0); }
                    if (c->tmap_state[i] == 0) {
                        const uint64_t rows = (uint64_t)std::max<uint64_t>(c->max_batch, 1) * op.hin * op.win;
                        const uint64_t dims[5] = {(uint64_t)op.cin, rows, 1, 1, 2};
                        const uint64_t strides[4] = {(uint64_t)op.cin * 2, rows * op.cin * 2, rows * op.cin * 2, (uint64_t)ip.plane * 2};
                        const uint32_t box[5] = {64, 128, 1, 1, 1};
                        const uint32_t es[5] = {1, 1, 1, 1, 1};
                        c->tmap_state[i] = tc_encode_tmap(&c->tmaps[i], ip.hi, dims, strides, box, es, 64) ? 1 : 2;
                    }
                    if (c->tmap_state[i] == 1) {
                        tp.tmap = c->tmaps[i];
                        tp.in_mode = TC_IN_TMA;
                        tp.kb = 64; tp.flat = 1;
                    }
                }
            } else {
                tp.in_f32 = c->d_tensor[op.in];
                tp.in_mode = TC_IN_F32;
                if (op.in_scale >= 0) return set_error(BN_ERR_INFERENCE, "gated vector input is not supported");
            }
            if (op.residual >= 0) {
                if (!is_spatial(p, op.residual)) return set_error(BN_ERR_INFERENCE, "vector residual is not supported");
                PlanesPtr rp = planes_of(c, op.residual);
                tp.res_hi = rp.hi; tp.res_plane = rp.plane;
            }
            tp.bias = d.bias;
            bool f32_rows = false;
            if (is_spatial(p, op.out) && op.kind == OP_CONV && (op.cout & 15) == 0) {
                if (c->f32_out_cache.empty()) c->f32_out_cache.assign(p.ops.size(), -1);
                if (c->f32_out_cache[i] < 0) {               // graph-only decision: made once per context
                    const int dwi = sole_dw_consumer(op.out);
                    DwSeParams sp;
                    c->f32_out_cache[i] = (dwi >= 0 && dw_se_plan((size_t)dwi, sp) >= 0 && sp.in_f32 != nullptr) ? 1 : 0;
                }
                f32_rows = c->f32_out_cache[i] == 1;
            }
            if (f32_rows) {
                tp.out_f32 = c->d_tensor[op.out];
                tp.out_f32_rows = 1;
            } else if (is_spatial(p, op.out)) {
                PlanesPtr o = planes_of(c, op.out);
                tp.out_hi = o.hi; tp.out_plane = o.plane;
            } else {
                tp.out_f32 = c->d_tensor[op.out];
            }
            // TMA-store epilogue: spatial outputs with >= 32 channels and no residual
            static const bool no_out_tma = [] { const char* ev = getenv("BN_DISABLE_OUT_TMA"); return ev && ev[0] == '1'; }();
            if (!no_out_tma && e->use_tma && is_spatial(p, op.out) && (op.cout & 15) == 0 && op.cout >= 32) {
                if (c->omap_state.empty()) { c->omaps.resize(p.ops.size()); c->omap_state.assign(p.ops.size(), 0); }
                const uint8_t want = f32_rows ? 1 : 3;                   // the map depends on the output form
                if (c->omap_state[i] != want && c->omap_state[i] != 2) {
                    const uint64_t rows = (uint64_t)std::max<uint64_t>(c->max_batch, 1) * op.hout * op.wout;
                    const bool ok = f32_rows ? tc_encode_out_tmap(&c->omaps[i], c->d_tensor[op.out], rows, (uint64_t)op.cout, true, 0)
                                             : tc_encode_out_tmap(&c->omaps[i], tp.out_hi, rows, (uint64_t)op.cout, false, tp.out_plane);
                    c->omap_state[i] = ok ? want : 2;
                }
                if (c->omap_state[i] == want) { tp.omap = c->omaps[i]; tp.out_tma = 1; }
            }
            tp.wpack = d.wpack;
            tp.batch = B; tp.hin = op.hin; tp.win = op.win; tp.cin = op.cin;
            tp.hout = op.hout; tp.wout = op.wout; tp.cout = op.cout;
            tp.k = op.k; tp.stride = op.stride; tp.pad = op.pad; tp.act = op.act;
            tp.K = op.k * op.k * op.cin;
            tp.M = B * op.hout * op.wout;
            tp.pix_stride = op.cin; tp.seg_stride = op.hin * op.win * op.cin; tp.tab_cin = op.cin;
            tp.k_chunks = d.k_chunks; tp.n_tiles = d.n_tiles; tp.m_tiles = (tp.M + 127) / 128;
            tp.tiles_per_seg = tp.m_tiles; tp.pix_per_seg = tp.M;
            tp.nt = d.nt; tp.stages = tp.in_mode == TC_IN_HALO ? d.halo_slots : d.stages; tp.tmem_cols = d.tmem_cols;
            tp.epi_warps = d.epi_warps;
            tp.prof = c->profiling && getenv("BN_TC_PROFILE") ? tc_conv_prof_slot((int)i) : nullptr;
            BN_CUDA(launch_tc_conv(tp, e->num_sms, s));
        } else if (is_spatial(p, op.in) || is_spatial(p, op.out)) {
            ConvPlanesParams cp{};
            cp.in = planes_of(c, op.in);
            cp.in_scale = (op.in_scale >= 0 && (int)i != prescaled_conv) ? c->d_tensor[op.in_scale] : nullptr;
            cp.weight = d.weight; cp.bias = d.bias;
            if (op.residual >= 0) cp.residual = planes_of(c, op.residual);
            cp.out = planes_of(c, op.out);
            cp.batch = B; cp.hin = op.hin; cp.win = op.win; cp.cin = op.cin; cp.hout = op.hout; cp.wout = op.wout;
            cp.cout = op.cout; cp.ldw = op.ldw; cp.k = op.k; cp.stride = op.stride; cp.pad = op.pad; cp.act = op.act;
            if (op.cout <= 32 && (op.cout & 7) == 0 && op.in_scale < 0 && op.residual < 0 && (op.cin == 1 || op.cin == 2) && op.k == 3 && op.stride == 2)
                BN_CUDA(launch_stem_planes(cp, s));
            else
                BN_CUDA(launch_conv_igemm_planes(cp, s));
        } else {
            ConvParams cp{c->d_tensor[op.in], op.in_scale >= 0 ? c->d_tensor[op.in_scale] : nullptr, d.weight, d.bias,
                          op.residual >= 0 ? c->d_tensor[op.residual] : nullptr, c->d_tensor[op.out],
                          B, op.hin, op.win, op.cin, op.hout, op.wout, op.cout, op.ldw, op.k, op.stride, op.pad, op.act};
            BN_CUDA(launch_conv_igemm(cp, s));
        }
        ++launches;
    }
    return BN_OK;
}

// ---- the forward pass (device side), input already in c->d_in or `d_audio` -----------------
static int enqueue_forward(bn_ctx* c, const float* d_audio, int B, const PostCfg& post, uint64_t k_eff) {
    bn_engine* e = c->eng;
    const Plan& p = e->plan;
    cudaStream_t s = c->stream;
    uint64_t launches = 0;
    const float* fe_in = d_audio;
    if (p.fe.kind != FE_LOGMEL) prof_mark(c, "normalize");
    const bool fe_on_tc = e->tc_mode && !e->fe_tc.empty() && p.fe.normalize;
    const bool fe_fused = fe_on_tc && e->fe_v24.on;
    if (fe_fused) {
        // min / max per segment (and, for tests only, the FP32 normalised copy); the spectrogram kernel normalises on the fly
        BN_CUDA(launch_minmax_normalize_fe(d_audio, c->keep_normalized ? c->d_norm : nullptr, c->d_minmax, nullptr, 0,
                                           B, p.sample_count, p.fe.eps, p.fe.half, p.fe.two, s));
        launches += c->keep_normalized ? 3 : 2;
        prof_mark(c, "spectrogram");
        const auto& fv = e->fe_v24;
        PlanesPtr o = planes_of(c, p.fe.out_tensor);
        SpecV24Params sp{};
        sp.audio = d_audio; sp.minmax = c->d_minmax; sp.n_br = fv.n_br;
        for (int sl = 0; sl < fv.n_br; ++sl) {
            const SpecBranch& br = p.fe.branches[fv.slot_branch[sl]];
            sp.br[sl] = SpecBranchDev{fv.wpack[sl], fv.hop[sl], fv.kcells[sl], fv.rows[sl], fv.blocks[sl], fv.split[sl], fv.pad[sl],
                                      fv.n_ksteps[sl], fv.slot_branch[sl], br.exponent};
        }
        const SpecBranch& b0 = p.fe.branches[0];
        sp.out_hi = o.hi; sp.out_plane = o.plane;
        sp.batch = B; sp.S = p.sample_count; sp.n_frames = b0.n_frames; sp.n_mels = b0.n_mels; sp.n_pad = fv.n_pad;
        sp.n_ch = (int)p.fe.branches.size(); sp.tiles_per_seg = (b0.n_frames + 127) / 128;
        sp.row_pitch = fv.row_pitch; sp.n_stages = fv.n_stages; sp.patch_plane = fv.patch_plane;
        { static const int ns = [] { const char* ev = getenv("BN_FE_STAGES"); return ev ? atoi(ev) : 0; }(); if (ns >= 2 && ns < sp.n_stages) sp.n_stages = ns; }
        sp.eps = p.fe.eps; sp.half = p.fe.half; sp.two = p.fe.two;
        { static const int dbg = [] { const char* ev = getenv("BN_FE_DEBUG"); return ev ? atoi(ev) : 0; }(); sp.debug = dbg; }
        sp.prof = c->profiling && getenv("BN_FE_PROFILE") ? tc_conv_prof_slot(127) : nullptr;
        BN_CUDA(launch_spec_v24(sp, fv.smem_bytes, e->num_sms, s));
        ++launches;
        fe_in = c->d_norm;
    } else if (fe_on_tc) {
        const size_t mb = std::max<uint64_t>(c->max_batch, 1);
        FePlaneOut outs[2];
        for (size_t bi = 0; bi < e->fe_tc.size(); ++bi) {
            const auto& ft = e->fe_tc[bi];
            outs[bi] = FePlaneOut{c->d_xp[bi], mb * (size_t)ft.rows * ft.row_stride, ft.hop, ft.row_stride, ft.rows};
        }
        BN_CUDA(launch_minmax_normalize_fe(d_audio, c->keep_normalized ? c->d_norm : nullptr, c->d_minmax, outs,
                                           (int)e->fe_tc.size(), B, p.sample_count, p.fe.eps, p.fe.half, p.fe.two, s));
        launches += 3;
        fe_in = c->d_norm;
    } else if (p.fe.normalize) {
        BN_CUDA(launch_minmax_normalize(d_audio, c->d_norm, B, p.sample_count, p.fe.eps, p.fe.half, p.fe.two, s));
        ++launches;
        fe_in = c->d_norm;
    }
    float* spec = c->d_tensor[p.fe.out_tensor];
    if (p.fe.kind == FE_LOGMEL) {
        const SpecBranch& br = p.fe.branches[0];
        const auto& lm = e->fe_lm;
        prof_mark(c, "logmel");
        LogmelParams lp{};
        lp.audio = d_audio; lp.batch = B; lp.sample_count = p.sample_count;
        lp.window = lm.window; lp.twiddle = reinterpret_cast<const float2*>(lm.twiddle);
        lp.mel_lo = lm.mel_lo; lp.mel_cnt = lm.mel_cnt; lp.mel_off = lm.mel_off; lp.mel_w = lm.mel_w;
        lp.n_fft = br.n_fft; lp.hop = br.hop; lp.n_frames = br.n_frames; lp.n_mels = br.n_mels;
        for (int i = 0; i < 8; ++i) lp.radix[i] = lm.radix[i];
        lp.n_stages = lm.n_stages;
        lp.log_floor = p.fe.log_floor; lp.log_scale = p.fe.log_scale;
        lp.out = planes_of(c, p.fe.out_tensor);
        lp.out_f32 = nullptr;
        BN_CUDA(launch_logmel(lp, s));
        ++launches;
    } else if (!fe_fused)
    for (size_t bi = 0; bi < p.fe.branches.size(); ++bi) {
        const SpecBranch& br = p.fe.branches[bi];
        prof_mark(c, bi == 0 ? "spectrogram0" : "spectrogram1");
        if (fe_on_tc) {
            const auto& ft = e->fe_tc[bi];
            const size_t mb = std::max<uint64_t>(c->max_batch, 1);
            PlanesPtr o = planes_of(c, p.fe.out_tensor);
            TcConvParams tp{};
            tp.in_hi = c->d_xp[bi];
            tp.in_plane = mb * (size_t)ft.rows * ft.row_stride;
            tp.in_mode = TC_IN_PLANES;
            tp.out_hi = o.hi; tp.out_plane = o.plane;
            tp.spec_nframes = br.n_frames; tp.spec_nch = (int)p.fe.branches.size(); tp.spec_ch = (int)bi;
            tp.spec_exponent = br.exponent;
            tp.wpack = ft.wpack;
            tp.batch = B; tp.hin = 1; tp.win = br.n_frames; tp.cin = ft.K; tp.hout = 1; tp.wout = br.n_frames;
            tp.cout = br.n_mels; tp.k = 1; tp.stride = 1; tp.pad = 0; tp.act = ACT_NONE;
            tp.K = ft.K; tp.M = B * br.n_frames; tp.k_chunks = ft.k_chunks; tp.n_tiles = ft.n_tiles;
            tp.m_tiles = (tp.M + 127) / 128;
            tp.pix_stride = ft.row_stride; tp.seg_stride = ft.rows * ft.row_stride; tp.tab_cin = ft.K;
            tp.nt = ft.nt; tp.stages = ft.stages; tp.tmem_cols = ft.tmem_cols;
            tp.prof = c->profiling && getenv("BN_TC_PROFILE") ? tc_conv_prof_slot(120 + (int)bi) : nullptr;
            BN_CUDA(launch_tc_conv(tp, e->num_sms, s));
        } else if (e->tc_mode)
            BN_CUDA(launch_spectrogram_v24_planes(fe_in, e->d_basis[bi], e->ldb[bi], planes_of(c, p.fe.out_tensor), B, p.sample_count,
                                                  br.n_fft, br.hop, br.n_frames, br.n_mels, (int)p.fe.branches.size(), (int)bi, br.exponent, s));
        else
            BN_CUDA(launch_spectrogram_v24(fe_in, e->d_basis[bi], e->ldb[bi], spec, B, p.sample_count, br.n_fft, br.hop,
                                           br.n_frames, br.n_mels, (int)p.fe.branches.size(), (int)bi, br.exponent, s));
        ++launches;
    }
    if (e->tc_mode) {
        int st = enqueue_ops_tc(c, B, launches);
        if (st != BN_OK) return st;
    } else {
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const PlanOp& op = p.ops[i];
        prof_mark(c, op.name.c_str());
        if (op.kind == OP_GAP) {
            BN_CUDA(launch_gap(c->d_tensor[op.in], c->d_tensor[op.out], B, op.hin * op.win, op.cin, s));
        } else if (op.kind == OP_DWCONV) {
            DwParams d{c->d_tensor[op.in], e->dev_ops[i].weight, e->dev_ops[i].bias, c->d_tensor[op.out],
                       B, op.hin, op.win, op.cout, op.hout, op.wout, op.k, op.stride, op.pad, op.act};
            BN_CUDA(launch_dwconv(d, s));
        } else {
            ConvParams cp{c->d_tensor[op.in], op.in_scale >= 0 ? c->d_tensor[op.in_scale] : nullptr,
                          e->dev_ops[i].weight, e->dev_ops[i].bias,
                          op.residual >= 0 ? c->d_tensor[op.residual] : nullptr, c->d_tensor[op.out],
                          B, op.hin, op.win, op.cin, op.hout, op.wout, op.cout, op.ldw, op.k, op.stride, op.pad, op.act};
            BN_CUDA(launch_conv_igemm(cp, s));
        }
        ++launches;
    }
    }
    prof_mark(c, "topk_epilogue");
    { const int ws = wait_for_fetch(c); if (ws != BN_OK) return ws; }     // top-k slots and counts are fetched too
    TopkParams tp{};
    tp.logits = c->d_tensor[p.logits_tensor];
    tp.batch = B;
    tp.n = p.num_species;
    tp.k = (uint32_t)k_eff;
    tp.has_min_conf = post.has_min_conf;
    tp.min_conf = post.min_conf;
    tp.range_state = post.range ? post.range->state : nullptr;
    tp.range_score = post.range ? post.range->score : nullptr;
    tp.rerank = post.range ? post.range->rerank : 0;
    tp.out = c->d_topk;
    tp.out_count = c->d_count;
    // d_count[max_batch] counts the segments of this run whose logits are not all finite (an activation that left the
    // fp16 range of the hi/lo operand format turns into NaN downstream; NaN audio does the same in the reference)
    tp.nonfinite = c->d_count + std::max<uint64_t>(c->max_batch, 1);
    BN_CUDA(cudaMemsetAsync(tp.nonfinite, 0, sizeof(uint32_t), s));
    BN_CUDA(launch_topk(tp, s));
    ++launches;
    prof_mark(c, "d2h");
    c->last_launches = launches;
    return BN_OK;
}

// Results leave on the copy stream, so the next batch's kernels (same compute stream) start while the 6.7 MB of
// logits are still crossing PCIe; the compute stream only waits for that copy right before it overwrites the result
// buffers again (wait_for_fetch, called in front of the first op that writes an output tensor).
static int enqueue_fetch(bn_ctx* c, int B, uint64_t k_eff) {
    const Plan& p = c->eng->plan;
    cudaStream_t s = c->profiling ? c->stream : c->copy_stream;      // profiling keeps everything on one timeline
    if (!c->profiling) {
        BN_CUDA(cudaEventRecord(c->ev_results, c->stream));
        BN_CUDA(cudaStreamWaitEvent(s, c->ev_results, 0));
    }
    BN_CUDA(cudaMemcpyAsync(c->h_logits, c->d_tensor[p.logits_tensor], (size_t)B * p.num_species * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (p.embedding_tensor >= 0)
        BN_CUDA(cudaMemcpyAsync(c->h_emb, c->d_tensor[p.embedding_tensor], (size_t)B * p.embedding_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (k_eff > 0) BN_CUDA(cudaMemcpyAsync(c->h_topk, c->d_topk, (size_t)B * k_eff * sizeof(Pred), cudaMemcpyDeviceToHost, s));
    BN_CUDA(cudaMemcpyAsync(c->h_count, c->d_count, (size_t)B * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    { const size_t mb = std::max<uint64_t>(c->max_batch, 1);
      BN_CUDA(cudaMemcpyAsync(c->h_count + mb, c->d_count + mb, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)); }
    if (!c->profiling) {
        BN_CUDA(cudaEventRecord(c->ev_fetched, s));
        c->fetch_pending = true;
    }
    return BN_OK;
}

// the `done` event of a run: after the fetch when there is one
static int record_done(bn_ctx* c, bool fetched) {
    BN_CUDA(cudaEventRecord(c->done, (fetched && !c->profiling) ? c->copy_stream : c->stream));
    return BN_OK;
}

static int wait_for_fetch(bn_ctx* c) {
    if (c->fetch_pending) {
        BN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fetched, 0));
        c->fetch_pending = false;
    }
    return BN_OK;
}

// Monitor loop of Classifier::run_inference (src/classifier.rs:527-554): completed first, then
// cancellation, then timeout.  A finished run wins over a fired monitor (classifier.rs:569).
static int wait_done(bn_ctx* c, const bn_run_opts* opts) {
    using clock = std::chrono::steady_clock;
    const bool monitor = opts && (opts->cancel_flag || opts->has_timeout);
    if (!monitor) {
        BN_CUDA(cudaEventSynchronize(c->done));
        return BN_OK;
    }
    const auto t0 = clock::now();
    int spins = 0;
    while (true) {
        cudaError_t q = cudaEventQuery(c->done);
        if (q == cudaSuccess) return BN_OK;
        if (q != cudaErrorNotReady) return cuda_fail(q, "cudaEventQuery");
        if (opts->cancel_flag && *opts->cancel_flag != 0) {
            c->draining = true;
            return set_error(BN_ERR_CANCELLED, "inference was cancelled");
        }
        if (opts->has_timeout) {
            uint64_t el = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(clock::now() - t0).count();
            if (el >= opts->timeout_ns) {
                c->draining = true;
                return set_error_detail(BN_ERR_TIMEOUT, "inference timed out", opts->timeout_ns, 0, 0);
            }
        }
        if (++spins < 200) std::this_thread::yield();
        else std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

static void fill_outputs(bn_ctx* c, uint64_t B, uint64_t k_eff, bn_outputs* out) {
    const Plan& p = c->eng->plan;
    memset(out, 0, sizeof(*out));
    out->batch = B;
    out->num_species = (uint64_t)p.num_species;
    out->logits = c->h_logits;
    out->embedding_dim = (uint64_t)p.embedding_dim;
    out->embeddings = p.embedding_tensor >= 0 ? c->h_emb : nullptr;
    out->topk_stride = k_eff;
    out->topk_count = c->h_count;
    out->topk = reinterpret_cast<const bn_pred*>(c->h_topk);
}

static int begin_run(bn_ctx* c, PostCfg& post, uint64_t& k_eff, const bn_run_opts* opts) {
    bn_engine* e = c->eng;
    BN_CUDA(cudaSetDevice(e->device));
    if (c->draining) {                       // a timed-out / cancelled run may still be in flight
        BN_CUDA(cudaStreamSynchronize(c->in_stream));
        if (!c->owns_stream) BN_CUDA(cudaStreamSynchronize(e->h2d));
        BN_CUDA(cudaStreamSynchronize(c->stream));
        BN_CUDA(cudaStreamSynchronize(c->copy_stream));
        c->draining = false;
    }
    {
        std::lock_guard<std::mutex> lk(e->post_mu);
        post = e->post;
    }
    if (post.range && post.range->n != (uint64_t)e->plan.num_species)
        return set_error(BN_ERR_INFERENCE, "range filter has " + std::to_string(post.range->n) + " classes, model has " + std::to_string(e->plan.num_species));
    k_eff = std::min<uint64_t>(post.top_k, (uint64_t)e->plan.num_species);   // postprocess.rs:50
    int st = ensure_topk_capacity(c, k_eff);
    if (st != BN_OK) return st;
    c->range_in_flight = post.range;
    c->prof_names.clear();
    // a token cancelled before the run starts terminates it at the first poll (classifier.rs:536-541)
    if (opts && opts->cancel_flag && *opts->cancel_flag != 0) return set_error(BN_ERR_CANCELLED, "inference was cancelled");
    return BN_OK;
}

int ctx_enqueue_device(bn_ctx* c, const float* d_audio, uint64_t batch, bool fetch, const bn_run_opts* opts) {
    if (!c) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) return BN_OK;
    if (batch > c->max_batch)
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    PostCfg post;
    uint64_t k_eff = 0;
    int st = begin_run(c, post, k_eff, opts);
    if (st != BN_OK) return st;
    {
        std::unique_lock<std::mutex> lane(c->eng->launch_mu[c->lane], std::defer_lock);
        if (!c->owns_stream) lane.lock();                  // a whole batch enters the shared lane at a time
        st = enqueue_forward(c, d_audio, (int)batch, post, k_eff);
        if (st != BN_OK) return st;
        if (fetch) { st = enqueue_fetch(c, (int)batch, k_eff); if (st != BN_OK) return st; }
        prof_mark(c, "end");
        st = record_done(c, fetch);
        if (st != BN_OK) return st;
    }
    c->pending_batch = batch;
    c->pending_k = k_eff;
    return BN_OK;
}

int ctx_wait(bn_ctx* c, const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    BN_CUDA(cudaSetDevice(c->eng->device));
    int st = wait_done(c, opts);
    if (st != BN_OK) return st;
    fill_outputs(c, c->pending_batch, c->pending_k, out);
    if (c->profiling) {
        size_t n = c->prof_names.size();
        c->prof_ms.assign(n ? n - 1 : 0, 0.f);
        for (size_t i = 0; i + 1 < n; ++i) cudaEventElapsedTime(&c->prof_ms[i], c->prof_events[i], c->prof_events[i + 1]);
    }
    return BN_OK;
}

int ctx_run_device(bn_ctx* c, const float* d_audio, uint64_t batch, bool fetch, const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) { memset(out, 0, sizeof(*out)); return BN_OK; }
    int st = ctx_enqueue_device(c, d_audio, batch, fetch, opts);
    if (st != BN_OK) return st;
    return ctx_wait(c, opts, out);
}

// ---- host staging: gather caller slices into the pinned slab and ship them chunk by chunk ----
// every slice (first and last byte) lies in page-locked host memory: the DMA engine can read it where it is
static bool segments_page_locked(const float* const* seg_ptrs, uint64_t B, size_t seg_bytes) {
    auto is_pinned = [](const void* ptr) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    if (!is_pinned(seg_ptrs[0])) return false;               // pageable callers pay for one query only
    for (uint64_t i = 0; i < B; ++i)                          // a slice may straddle the end of a registered range
        if (!is_pinned(seg_ptrs[i]) || !is_pinned(reinterpret_cast<const char*>(seg_ptrs[i]) + seg_bytes - 1)) return false;
    return true;
}

// DMA straight from the caller's page-locked slices, contiguous runs as one copy
static int copy_page_locked_segments(bn_ctx* c, const float* const* seg_ptrs, uint64_t B, cudaStream_t s) {
    const size_t S = (size_t)c->eng->plan.sample_count;
    const size_t seg_bytes = S * sizeof(float);
    uint64_t i = 0;
    while (i < B) {
        uint64_t j = i + 1;
        while (j < B && seg_ptrs[j] == seg_ptrs[j - 1] + S && (j - i) < 32) ++j;     // <= 32 segments (18 MB) per copy
        BN_CUDA(cudaMemcpyAsync(c->d_in + i * S, seg_ptrs[i], (j - i) * seg_bytes, cudaMemcpyHostToDevice, s));
        i = j;
    }
    return BN_OK;
}

// pageable slices: gather into the pinned slab and ship chunk by chunk on this context's input stream
static int stage_input(bn_ctx* c, const float* const* seg_ptrs, uint64_t B) {
    bn_engine* e = c->eng;
    const size_t S = (size_t)e->plan.sample_count;
    const size_t seg_bytes = S * sizeof(float);
    const uint64_t chunk = 8;
    const uint64_t n_chunks = (B + chunk - 1) / chunk;
    int T = (int)std::min<uint64_t>((uint64_t)e->pack_threads, n_chunks);
    if (T <= 1) {
        for (uint64_t ci = 0; ci < n_chunks; ++ci) {
            uint64_t lo = ci * chunk, hi = std::min(B, lo + chunk);
            for (uint64_t i = lo; i < hi; ++i) stream_copy(c->h_in + i * S, seg_ptrs[i], seg_bytes);   // batch_context.rs:209-211
            BN_CUDA(cudaMemcpyAsync(c->d_in + lo * S, c->h_in + lo * S, (hi - lo) * seg_bytes, cudaMemcpyHostToDevice, c->in_stream));
        }
        return BN_OK;
    }
    std::atomic<uint64_t> next{0};
    std::atomic<int> err{(int)cudaSuccess};
    auto worker = [&]() {
        cudaSetDevice(e->device);
        while (true) {
            uint64_t ci = next.fetch_add(1);
            if (ci >= n_chunks) break;
            uint64_t lo = ci * chunk, hi = std::min(B, lo + chunk);
            for (uint64_t i = lo; i < hi; ++i) stream_copy(c->h_in + i * S, seg_ptrs[i], seg_bytes);
            cudaError_t ce = cudaMemcpyAsync(c->d_in + lo * S, c->h_in + lo * S, (hi - lo) * seg_bytes, cudaMemcpyHostToDevice, c->in_stream);
            if (ce != cudaSuccess) err.store((int)ce);
        }
    };
    std::vector<std::thread> th;
    th.reserve(T - 1);
    for (int t = 0; t < T - 1; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    if (err.load() != (int)cudaSuccess) return cuda_fail((cudaError_t)err.load(), "cudaMemcpyAsync(H2D)");
    return BN_OK;
}

// BN_TRACE_RUN=1 (dev aid): one stderr line per bn_ctx_run with host timestamps and device timestamps of the H2D copy,
// the forward pass and the fetch, all in ms since the first traced call
struct RunTrace {
    cudaEvent_t ref = nullptr;
    std::chrono::steady_clock::time_point host0;
    std::mutex mu;
};
static RunTrace g_trace;
static bool trace_enabled() {
    static const bool v = [] { const char* ev = getenv("BN_TRACE_RUN"); return ev && ev[0] == '1'; }();
    return v;
}
static double host_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_trace.host0).count(); }

int ctx_run_host(bn_ctx* c, const float* const* seg_ptrs, const uint64_t* seg_lens, uint64_t batch,
                 bool check_max_first, const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) { memset(out, 0, sizeof(*out)); return BN_OK; }          // classifier.rs:681-683, 832-834
    if (!seg_ptrs || !seg_lens) return set_error(BN_ERR_INVALID_ARGUMENT, "null segment array");
    const uint64_t S = (uint64_t)c->eng->plan.sample_count;
    if (check_max_first && batch > c->max_batch)                               // batch_context.rs:191-196
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    for (uint64_t i = 0; i < batch; ++i)                                       // classifier.rs:688-696, batch_context.rs:199-206
        if (seg_lens[i] != S)
            return set_error_detail(BN_ERR_BATCH_INPUT_SIZE, "batch input size mismatch", i, S, seg_lens[i]);
    if (batch > c->max_batch)
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    PostCfg post;
    uint64_t k_eff = 0;
    int st = begin_run(c, post, k_eff, opts);
    if (st != BN_OK) return st;
    const bool tr = trace_enabled() && !c->owns_stream;
    cudaEvent_t te[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double th[5] = {0, 0, 0, 0, 0};
    if (tr) {
        std::lock_guard<std::mutex> lk(g_trace.mu);
        if (!g_trace.ref) {
            cudaEventCreate(&g_trace.ref);
            cudaDeviceSynchronize();
            cudaEventRecord(g_trace.ref, c->eng->h2d);
            cudaEventSynchronize(g_trace.ref);
            g_trace.host0 = std::chrono::steady_clock::now();
        }
        for (auto& e : te) cudaEventCreate(&e);
        th[0] = host_ms();
    }
    const bool in_place = segments_page_locked(seg_ptrs, batch, (size_t)S * sizeof(float));
    c->last_in_place = in_place;
    if (tr) th[1] = host_ms();
    if (c->owns_stream) prof_mark(c, "h2d");
    if (!in_place || c->owns_stream) {
        st = in_place ? copy_page_locked_segments(c, seg_ptrs, batch, c->in_stream) : stage_input(c, seg_ptrs, batch);
        if (st != BN_OK) return st;
    }
    {
        std::unique_lock<std::mutex> lane(c->eng->launch_mu[c->lane], std::defer_lock);
        if (!c->owns_stream) {                             // a whole batch enters the shared lanes at a time
            cudaStream_t src = c->in_stream;
            if (in_place) {
                // page-locked input goes through the engine's one H2D lane: copies of concurrent callers run one after
                // the other in the order their kernels will, instead of sharing PCIe and all arriving late
                lane.lock();
                std::lock_guard<std::mutex> hl(c->eng->h2d_mu);     // always lane -> h2d
                if (tr) { th[2] = host_ms(); cudaEventRecord(te[0], c->eng->h2d); }
                st = copy_page_locked_segments(c, seg_ptrs, batch, c->eng->h2d);
                if (st != BN_OK) return st;
                if (tr) cudaEventRecord(te[1], c->eng->h2d);
                BN_CUDA(cudaEventRecord(c->ev_in, c->eng->h2d));
            } else
                BN_CUDA(cudaEventRecord(c->ev_in, src));
            if (!in_place) lane.lock();
            prof_mark(c, "h2d");                           // on the lane: the time the lane waits for this batch's input
            BN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in, 0));
            if (tr) cudaEventRecord(te[2], c->stream);
        }
        st = enqueue_forward(c, c->d_in, (int)batch, post, k_eff);
        if (st != BN_OK) return st;
        if (tr) cudaEventRecord(te[3], c->stream);
        st = enqueue_fetch(c, (int)batch, k_eff);
        if (st != BN_OK) return st;
        prof_mark(c, "end");
        st = record_done(c, true);
        if (st != BN_OK) return st;
        if (tr) { cudaEventRecord(te[4], c->copy_stream); th[3] = host_ms(); }
    }
    st = wait_done(c, opts);
    if (st != BN_OK) return st;
    if (tr) {
        th[4] = host_ms();
        float g[5] = {0, 0, 0, 0, 0};
        cudaEventSynchronize(te[4]);
        for (int i = 0; i < 5; ++i) if (in_place || i >= 2) cudaEventElapsedTime(&g[i], g_trace.ref, te[i]);
        fprintf(stderr, "[trace] ctx=%p host: enter %.2f checked %.2f locked %.2f enqueued %.2f done %.2f | gpu: h2d %.2f-%.2f fwd %.2f-%.2f fetched %.2f\n",
                (void*)c, th[0], th[1], th[2], th[3], th[4], g[0], g[1], g[2], g[3], g[4]);
        for (auto& e : te) cudaEventDestroy(e);
    }
    fill_outputs(c, batch, k_eff, out);
    if (c->profiling) {
        size_t n = c->prof_names.size();
        c->prof_ms.assign(n ? n - 1 : 0, 0.f);
        for (size_t i = 0; i + 1 < n; ++i) cudaEventElapsedTime(&c->prof_ms[i], c->prof_events[i], c->prof_events[i + 1]);
    }
    return BN_OK;
}

// CLI ingest (src/bin/birdnet-analyze.rs:653-743) with the conversion and the chunking on the device: the
// recording crosses PCIe once as 16-bit PCM (2 B/sample, overlap not duplicated) instead of as FP32 segments.
int ctx_run_pcm16(bn_ctx* c, const int16_t* pcm, uint64_t n_samples, uint64_t first_pos, uint64_t step, uint64_t batch,
                  const bn_run_opts* opts, bn_outputs* out) {
    if (!c || !out) return set_error(BN_ERR_INVALID_ARGUMENT, "null argument");
    if (batch == 0) { memset(out, 0, sizeof(*out)); return BN_OK; }
    if (!pcm) return set_error(BN_ERR_INVALID_ARGUMENT, "null pcm buffer");
    bn_engine* e = c->eng;
    const uint64_t S = (uint64_t)e->plan.sample_count;
    if (batch > c->max_batch)
        return set_error(BN_ERR_INFERENCE, "batch size " + std::to_string(batch) + " exceeds context max " + std::to_string(c->max_batch));
    if (step == 0 || step > S) return set_error(BN_ERR_INVALID_ARGUMENT, "step must be in [1, sample_count]");
    if (first_pos >= n_samples) return set_error(BN_ERR_INVALID_ARGUMENT, "first segment starts past the end of the recording");
    PostCfg post;
    uint64_t k_eff = 0;
    int st = begin_run(c, post, k_eff, opts);
    if (st != BN_OK) return st;
    const size_t cap = (size_t)c->max_batch * S;
    if (!c->h_pcm) {
        BN_CUDA(cudaHostAlloc(&c->h_pcm, cap * sizeof(int16_t), cudaHostAllocDefault));
        BN_CUDA(cudaMalloc(&c->d_pcm, cap * sizeof(int16_t)));
    }
    const uint64_t hi = std::min<uint64_t>(n_samples, first_pos + (batch - 1) * step + S);
    const size_t n = (size_t)(hi - first_pos);                      // <= (batch-1)*step + S <= max_batch * S
    prof_mark(c, "h2d");
    bool pinned = false;                                            // recording already page-locked: DMA in place
    {
        cudaPointerAttributes a0{}, a1{};
        if (cudaPointerGetAttributes(&a0, pcm + first_pos) == cudaSuccess && cudaPointerGetAttributes(&a1, pcm + first_pos + n - 1) == cudaSuccess)
            pinned = a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
        else
            cudaGetLastError();
    }
    const size_t piece = (size_t)4 << 20;                           // samples per pipelined copy (8 MiB)
    for (size_t o = 0; o < n; o += piece) {
        const size_t m = std::min(piece, n - o);
        const int16_t* src = pcm + first_pos + o;
        if (!pinned) { stream_copy(c->h_pcm + o, src, m * sizeof(int16_t)); src = c->h_pcm + o; }
        if (pinned && !c->owns_stream) continue;                    // goes through the engine's H2D lane below
        BN_CUDA(cudaMemcpyAsync(c->d_pcm + o, src, m * sizeof(int16_t), cudaMemcpyHostToDevice, c->in_stream));
    }
    {
        std::unique_lock<std::mutex> lane(c->eng->launch_mu[c->lane], std::defer_lock);
        if (!c->owns_stream) {
            cudaStream_t src_stream = c->in_stream;
            if (pinned) {
                lane.lock();
                std::lock_guard<std::mutex> hl(c->eng->h2d_mu);     // always lane -> h2d
                for (size_t o = 0; o < n; o += piece)
                    BN_CUDA(cudaMemcpyAsync(c->d_pcm + o, pcm + first_pos + o, std::min(piece, n - o) * sizeof(int16_t),
                                            cudaMemcpyHostToDevice, c->eng->h2d));
                BN_CUDA(cudaEventRecord(c->ev_in, c->eng->h2d));
            } else
                BN_CUDA(cudaEventRecord(c->ev_in, src_stream));
            if (!pinned) lane.lock();
            BN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in, 0));
        }
        BN_CUDA(launch_pcm16_to_segments(c->d_pcm, first_pos, n_samples, first_pos, step, c->d_in, (int)batch, (int)S, c->stream));
        st = enqueue_forward(c, c->d_in, (int)batch, post, k_eff);
        if (st != BN_OK) return st;
        c->last_launches += 1;
        st = enqueue_fetch(c, (int)batch, k_eff);
        if (st != BN_OK) return st;
        prof_mark(c, "end");
        st = record_done(c, true);
        if (st != BN_OK) return st;
    }
    st = wait_done(c, opts);
    if (st != BN_OK) return st;
    fill_outputs(c, batch, k_eff, out);
    if (c->profiling) {
        size_t np = c->prof_names.size();
        c->prof_ms.assign(np ? np - 1 : 0, 0.f);
        for (size_t i = 0; i + 1 < np; ++i) cudaEventElapsedTime(&c->prof_ms[i], c->prof_events[i], c->prof_events[i + 1]);
    }
    return BN_OK;
}

}  // namespace bn
