// See plan.h.
#include "plan.h"

#include <cmath>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

namespace bn {
namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error("unsupported graph: " + msg); }

struct Matcher {
    const OnnxModel& m;
    Plan& plan;
    std::map<std::string, int> tensor_id;                    // value name -> tensor id
    std::map<std::string, std::vector<int>> consumers;       // value name -> node indices
    std::vector<bool> consumed;

    Matcher(const OnnxModel& m_, Plan& p) : m(m_), plan(p), consumed(m_.nodes.size(), false) {
        for (size_t i = 0; i < m.nodes.size(); ++i)
            for (auto& in : m.nodes[i].inputs) {
                auto& v = consumers[in];
                if (v.empty() || v.back() != (int)i) v.push_back((int)i);   // Mul(x, x) counts once
            }
    }

    int new_tensor(const std::string& name, int C, int H, int W) {
        TensorInfo t;
        t.name = name; t.C = C; t.H = H; t.W = W;
        plan.tensors.push_back(t);
        int id = (int)plan.tensors.size() - 1;
        tensor_id[name] = id;
        return id;
    }
    int lookup(const std::string& name) const {
        auto it = tensor_id.find(name);
        if (it == tensor_id.end()) fail("value '" + name + "' is used before it is produced");
        return it->second;
    }
    const OnnxTensor& need_init(const std::string& name) const {
        auto* t = m.init(name);
        if (!t) fail("expected '" + name + "' to be an initializer");
        return *t;
    }
    // the single consumer node of `value`, or -1
    int sole_consumer(const std::string& value) const {
        auto it = consumers.find(value);
        if (it == consumers.end() || it->second.size() != 1) return -1;
        return it->second[0];
    }

    // ---------------------------------------------------------------- front-end
    // Follows the single-consumer chain that starts at an STFT output and fills a SpecBranch.
    // Returns the name of the last value of the chain (the branch's NCHW tensor or "spec").
    std::string match_branch(int stft_idx, SpecBranch& br, bool& is_logmel, std::string& spectro_name) {
        const OnnxNode& st = m.nodes[stft_idx];
        if (st.inputs.size() < 4) fail("STFT needs signal, frame_step, window, frame_length");
        if (st.attr_i("onesided", 1) != 1) fail("STFT must be onesided");
        br.hop = (int)need_init(st.inputs[1]).i64();
        const OnnxTensor& win = need_init(st.inputs[2]);
        br.n_fft = (int)need_init(st.inputs[3]).i64();
        if ((int)win.numel() != br.n_fft) fail("STFT window length != frame_length");
        br.window.assign(win.f32(), win.f32() + br.n_fft);
        br.n_bins = br.n_fft / 2 + 1;
        consumed[stft_idx] = true;

        // torch's exporter keeps torch.stft's [B, bins, frames, 2] layout: Transpose(0,2,1,3) in front of the gathers and
        // Transpose(0,2,1) behind each of them; the pair cancels (the in-house writer emits neither)
        std::string stft_out = st.outputs[0];
        bool swapped = false;
        {
            const int t = sole_consumer(stft_out);
            if (t >= 0 && m.nodes[t].op == "Transpose") {
                const OnnxAttr* pm = m.nodes[t].attr("perm");
                if (!pm || pm->ints != std::vector<int64_t>{0, 2, 1, 3}) fail("STFT output: only Transpose(0,2,1,3) is understood");
                consumed[t] = true;
                stft_out = m.nodes[t].outputs[0];
                swapped = true;
            }
        }
        auto unswap = [&](std::string& v) {
            if (!swapped || v.empty()) return;
            const int t = sole_consumer(v);
            const OnnxAttr* pm = t >= 0 ? m.nodes[t].attr("perm") : nullptr;
            if (t < 0 || m.nodes[t].op != "Transpose" || !pm || pm->ints != std::vector<int64_t>{0, 2, 1})
                fail("STFT output: expected Transpose(0,2,1) behind the real / imaginary Gather");
            consumed[t] = true;
            v = m.nodes[t].outputs[0];
        };
        // real / imag gathers
        auto& cons = consumers[stft_out];
        std::string re, im;
        for (int ci : cons) {
            const OnnxNode& g = m.nodes[ci];
            if (g.op != "Gather" || g.attr_i("axis", 0) != 3) fail("STFT output must feed Gather(axis=3)");
            int64_t idx = need_init(g.inputs[1]).i64();
            (idx == 0 ? re : im) = g.outputs[0];
            consumed[ci] = true;
        }
        if (re.empty()) fail("STFT real part is unused");
        unswap(re);
        unswap(im);
        std::string cur;
        if (im.empty()) {
            is_logmel = false;
            cur = re;
        } else {
            // re*re + im*im -> Sqrt
            is_logmel = true;
            int a = sole_consumer(re), b = sole_consumer(im);
            if (a < 0 || b < 0 || m.nodes[a].op != "Mul" || m.nodes[b].op != "Mul") fail("log-mel: expected re*re and im*im");
            consumed[a] = consumed[b] = true;
            int s = sole_consumer(m.nodes[a].outputs[0]);
            if (s < 0 || m.nodes[s].op != "Add") fail("log-mel: expected Add(re^2, im^2)");
            consumed[s] = true;
            int q = sole_consumer(m.nodes[s].outputs[0]);
            if (q < 0 || m.nodes[q].op != "Sqrt") fail("log-mel: expected Sqrt");
            consumed[q] = true;
            cur = m.nodes[q].outputs[0];
        }
        // mel projection
        int mm = sole_consumer(cur);
        if (mm < 0 || m.nodes[mm].op != "MatMul") fail("front-end: expected MatMul with the mel matrix");
        const OnnxTensor& mel = need_init(m.nodes[mm].inputs[1]);
        if (mel.dims.size() != 2 || mel.dims[0] != br.n_bins) fail("mel matrix must be [n_fft/2+1, n_mels]");
        br.n_mels = (int)mel.dims[1];
        br.mel.assign(mel.f32(), mel.f32() + mel.numel());
        consumed[mm] = true;
        cur = m.nodes[mm].outputs[0];

        if (!is_logmel) {
            // Mul(x,x) -> Pow(exponent) -> [Slice flip] -> Transpose -> Unsqueeze
            int sq = -1;
            for (int ci : consumers[cur]) { sq = ci; }
            if (sq < 0 || m.nodes[sq].op != "Mul" || m.nodes[sq].inputs[0] != cur || m.nodes[sq].inputs[1] != cur)
                fail("v2.4 front-end: expected Mul(mel, mel)");
            consumed[sq] = true;
            int pw = sole_consumer(m.nodes[sq].outputs[0]);
            if (pw < 0 || m.nodes[pw].op != "Pow") fail("v2.4 front-end: expected Pow");
            br.exponent = need_init(m.nodes[pw].inputs[1]).f32_at();
            consumed[pw] = true;
            cur = m.nodes[pw].outputs[0];
            int nx = sole_consumer(cur);
            if (nx >= 0 && m.nodes[nx].op == "Slice") {
                const OnnxNode& sl = m.nodes[nx];
                if (sl.inputs.size() < 5 || need_init(sl.inputs[3]).i64() != 2 || need_init(sl.inputs[4]).i64() != -1)
                    fail("v2.4 front-end: Slice must flip axis 2 with step -1");
                br.flip = true;
                consumed[nx] = true;
                cur = sl.outputs[0];
                nx = sole_consumer(cur);
            }
            if (nx < 0 || m.nodes[nx].op != "Transpose") fail("v2.4 front-end: expected Transpose");
            consumed[nx] = true;
            cur = m.nodes[nx].outputs[0];
            nx = sole_consumer(cur);
            if (nx < 0 || m.nodes[nx].op != "Unsqueeze") fail("v2.4 front-end: expected Unsqueeze");
            consumed[nx] = true;
            return m.nodes[nx].outputs[0];
        }
        // Add(floor) -> Log -> Mul(scale) -> Unsqueeze
        int ad = sole_consumer(cur);
        if (ad < 0 || m.nodes[ad].op != "Add") fail("log-mel: expected Add(log_floor)");
        plan.fe.log_floor = need_init(m.nodes[ad].inputs[1]).f32_at();
        consumed[ad] = true;
        int lg = sole_consumer(m.nodes[ad].outputs[0]);
        if (lg < 0 || m.nodes[lg].op != "Log") fail("log-mel: expected Log");
        consumed[lg] = true;
        int sc = sole_consumer(m.nodes[lg].outputs[0]);
        if (sc < 0 || m.nodes[sc].op != "Mul") fail("log-mel: expected Mul(log_scale)");
        plan.fe.log_scale = need_init(m.nodes[sc].inputs[1]).f32_at();
        consumed[sc] = true;
        spectro_name = m.nodes[sc].outputs[0];
        int un = -1;
        for (int ci : consumers[spectro_name])
            if (m.nodes[ci].op == "Unsqueeze") un = ci;
        if (un < 0) fail("log-mel: expected Unsqueeze to NCHW");
        consumed[un] = true;
        return m.nodes[un].outputs[0];
    }

    void match_frontend() {
        FrontEndPlan& fe = plan.fe;
        if (m.inputs.size() != 1) fail("model must have exactly one input");
        plan.input_name = m.inputs[0].name;
        plan.input_dims = m.inputs[0].dims;
        const auto& id = plan.input_dims;
        if (id.size() != 2 && id.size() != 3) fail("input must be [batch, samples] or [batch, 1, samples]");
        int64_t sc = id.back();
        if (sc <= 0) fail("input sample count must be static");
        fe.sample_count = (int)sc;

        std::string cur = plan.input_name;
        int n0 = sole_consumer(cur);
        // optional per-segment min/max normalisation
        bool has_min = false;
        for (int ci : consumers[cur]) if (m.nodes[ci].op == "ReduceMin") has_min = true;
        if (has_min) {
            // ReduceMin, Sub, ReduceMax, Add(eps), Div, Sub(half), Mul(two)
            const char* seq[] = {"ReduceMin", "Sub", "ReduceMax", "Add", "Div", "Sub", "Mul"};
            size_t i = 0;
            std::vector<int> idx;
            for (size_t n = 0; n < m.nodes.size() && i < 7; ++n) {
                if (consumed[n]) continue;
                if (m.nodes[n].op != seq[i]) fail(std::string("normaliser: expected ") + seq[i] + ", found " + m.nodes[n].op);
                consumed[n] = true;
                idx.push_back((int)n);
                ++i;
            }
            if (i != 7) fail("normaliser: incomplete");
            fe.normalize = true;
            fe.eps = need_init(m.nodes[idx[3]].inputs[1]).f32_at();
            fe.half = need_init(m.nodes[idx[5]].inputs[1]).f32_at();
            fe.two = need_init(m.nodes[idx[6]].inputs[1]).f32_at();
            cur = m.nodes[idx[6]].outputs[0];
            n0 = sole_consumer(cur);
        }
        if (n0 >= 0 && m.nodes[n0].op == "Pad") {
            const OnnxTensor& pads = need_init(m.nodes[n0].inputs[1]);
            if (pads.numel() != 4 || pads.i64(0) || pads.i64(1) || pads.i64(2)) fail("Pad must only pad the end of axis 1");
            fe.pad_end = (int)pads.i64(3);
            consumed[n0] = true;
            cur = m.nodes[n0].outputs[0];
            n0 = sole_consumer(cur);
        }
        // ONNX STFT takes [B,S,1]; torch's exporter feeds the [B,S] signal directly (ONNX Runtime accepts both)
        if (n0 >= 0 && m.nodes[n0].op == "Unsqueeze") {
            consumed[n0] = true;
            cur = m.nodes[n0].outputs[0];
        }

        std::vector<std::string> branch_out;
        bool any_logmel = false;
        std::string spectro_name;
        for (int ci : consumers[cur]) {
            if (m.nodes[ci].op != "STFT") fail("front-end: signal must feed STFT nodes only");
            SpecBranch br;
            bool lm = false;
            branch_out.push_back(match_branch(ci, br, lm, spectro_name));
            any_logmel |= lm;
            br.n_frames = 1 + (fe.sample_count + fe.pad_end - br.n_fft) / br.hop;
            fe.branches.push_back(std::move(br));
        }
        if (fe.branches.empty()) fail("front-end: no STFT node found");
        fe.kind = any_logmel ? FE_LOGMEL : FE_BIRDNET_V24;
        std::string spec_name;
        if (fe.kind == FE_BIRDNET_V24) {
            if (!fe.normalize) fail("v2.4 front-end requires the min/max normaliser");
            int cc = sole_consumer(branch_out[0]);
            if (cc < 0 || m.nodes[cc].op != "Concat" || m.nodes[cc].attr_i("axis", 0) != 1) fail("v2.4 front-end: expected Concat(axis=1)");
            if (m.nodes[cc].inputs != branch_out) fail("v2.4 front-end: Concat inputs must be the spectrogram branches in order");
            consumed[cc] = true;
            spec_name = m.nodes[cc].outputs[0];
            for (auto& b : fe.branches)
                if (b.n_mels != fe.branches[0].n_mels || b.n_frames != fe.branches[0].n_frames)
                    fail("v2.4 front-end: branches must produce equal [mels, frames]");
            fe.out_tensor = new_tensor(spec_name, (int)fe.branches.size(), fe.branches[0].n_mels, fe.branches[0].n_frames);
        } else {
            if (fe.branches.size() != 1) fail("log-mel front-end supports one STFT branch");
            spec_name = branch_out[0];
            fe.out_tensor = new_tensor(spec_name, 1, fe.branches[0].n_frames, fe.branches[0].n_mels);
            int sid = new_tensor(spectro_name, 1, fe.branches[0].n_frames, fe.branches[0].n_mels);
            plan.tensors[sid].alias_of = fe.out_tensor;
            fe.spectrogram_tensor = sid;
        }
    }

    // ---------------------------------------------------------------- CNN
    void match_cnn() {
        for (size_t n = 0; n < m.nodes.size(); ++n) {
            if (consumed[n]) continue;
            const OnnxNode& nd = m.nodes[n];
            consumed[n] = true;
            if (nd.op == "Conv") match_conv((int)n);
            else if (nd.op == "Gemm") match_gemm((int)n);
            else if (nd.op == "GlobalAveragePool" || nd.op == "ReduceMean") {
                if (nd.op == "ReduceMean") {             // torch: x.mean(dim=(2,3), keepdim=True)
                    const OnnxAttr* ax = nd.attr("axes");
                    if (!ax || ax->ints != std::vector<int64_t>{2, 3} || nd.attr_i("keepdims", 1) != 1)
                        fail("ReduceMean must average axes (2,3) with keepdims=1 (a global average pool)");
                }
                int in = resolve_plain(nd.inputs[0], "GlobalAveragePool");
                PlanOp op;
                op.kind = OP_GAP; op.name = nd.name; op.in = in;
                const TensorInfo& ti = plan.tensors[in];
                op.cin = op.cout = ti.C; op.hin = ti.H; op.win = ti.W;
                op.out = new_tensor(nd.outputs[0], ti.C, 1, 1);
                plan.ops.push_back(std::move(op));
            } else if (nd.op == "Mul") {
                int a = lookup(nd.inputs[0]), b = lookup(nd.inputs[1]);
                const TensorInfo &ta = plan.tensors[a], &tb = plan.tensors[b];
                if (tb.H == 1 && tb.W == 1 && tb.C == ta.C && (ta.H > 1 || ta.W > 1)) {
                    int v = new_tensor(nd.outputs[0], ta.C, ta.H, ta.W);
                    plan.tensors[v].scale_base = a; plan.tensors[v].scale_vec = b;
                } else if (ta.H == 1 && ta.W == 1 && ta.C == tb.C && (tb.H > 1 || tb.W > 1)) {
                    int v = new_tensor(nd.outputs[0], tb.C, tb.H, tb.W);
                    plan.tensors[v].scale_base = b; plan.tensors[v].scale_vec = a;
                } else fail("Mul '" + nd.name + "' is not a squeeze-excite gate");
            } else if (nd.op == "Flatten") {
                int in = resolve_plain(nd.inputs[0], "Flatten");
                const TensorInfo& ti = plan.tensors[in];
                if (ti.H != 1 || ti.W != 1) fail("Flatten is only supported after global pooling");
                int v = new_tensor(nd.outputs[0], ti.C, 1, 1);
                plan.tensors[v].alias_of = in;
            } else if (nd.op == "Transpose") {
                auto* perm = nd.attr("perm");
                if (!perm || perm->ints != std::vector<int64_t>{0, 2, 3, 1}) fail("Transpose must be NCHW->NHWC");
                int in = resolve_plain(nd.inputs[0], "Transpose");
                const TensorInfo& ti = plan.tensors[in];
                int v = new_tensor(nd.outputs[0], ti.C, ti.H, ti.W);   // engine layout is NHWC already
                plan.tensors[v].alias_of = in;
            } else {
                fail("op '" + nd.op + "' (node '" + nd.name + "') is not supported");
            }
        }
    }

    int resolve_plain(const std::string& name, const char* who) {
        int t = lookup(name);
        if (plan.tensors[t].scale_base >= 0) fail(std::string(who) + " cannot consume a gated (virtual) tensor");
        return t;
    }

    // after a producer node: fold Sigmoid+Mul (SiLU) / Sigmoid; returns the final value name
    std::string fold_activation(const std::string& raw, int& act) {
        act = ACT_NONE;
        auto it = consumers.find(raw);
        if (it == consumers.end()) return raw;
        const auto& cs = it->second;
        int sig = -1, mul = -1;
        for (int ci : cs) {
            if (m.nodes[ci].op == "Sigmoid") sig = ci;
            else if (m.nodes[ci].op == "Mul") mul = ci;
        }
        if (sig >= 0 && mul >= 0 && cs.size() == 2) {
            const OnnxNode& mu = m.nodes[mul];
            const std::string& so = m.nodes[sig].outputs[0];
            bool ok = (mu.inputs[0] == raw && mu.inputs[1] == so) || (mu.inputs[1] == raw && mu.inputs[0] == so);
            if (ok && sole_consumer(so) == mul) {
                consumed[sig] = consumed[mul] = true;
                act = ACT_SILU;
                return mu.outputs[0];
            }
        }
        if (sig >= 0 && cs.size() == 1) {
            consumed[sig] = true;
            act = ACT_SIGMOID;
            return m.nodes[sig].outputs[0];
        }
        return raw;
    }

    // residual: the activated value feeds exactly one Add whose other operand already exists
    std::string fold_residual(const std::string& val, int& residual) {
        residual = -1;
        int ad = sole_consumer(val);
        if (ad < 0 || m.nodes[ad].op != "Add" || consumed[ad]) return val;
        const OnnxNode& a = m.nodes[ad];
        const std::string& other = a.inputs[0] == val ? a.inputs[1] : a.inputs[0];
        auto it = tensor_id.find(other);
        if (it == tensor_id.end() || plan.tensors[it->second].scale_base >= 0) return val;
        residual = it->second;
        consumed[ad] = true;
        return a.outputs[0];
    }

    // Operand format of the tensor-core path: x = fp16 hi + fp16 lo (tc_common.cuh split8).  It carries ~22 mantissa
    // bits but only the fp16 RANGE: |x| > 65504 would become inf - inf = NaN downstream.  Weights are checked here, at
    // load time (the reference reports unusable models through ModelLoad as well); activations are the caller's
    // business and are flagged at run time (bn_ctx_nonfinite_segments).
    void check_weight_range(const OnnxTensor& W, const std::string& layer) {
        const float* w = W.f32();
        const size_t n = W.numel();
        for (size_t i = 0; i < n; ++i)
            if (!(std::fabs(w[i]) <= 65504.0f))
                fail("layer '" + layer + "': weight " + std::to_string(w[i]) + " is outside the fp16 range (|w| <= 65504) of the hi/lo operand format");
    }

    void match_conv(int n) {
        const OnnxNode& nd = m.nodes[n];
        if (nd.inputs.size() != 3) fail("Conv '" + nd.name + "' must have a bias (BatchNorm folded)");
        const OnnxTensor& W = need_init(nd.inputs[1]);
        const OnnxTensor& Bv = need_init(nd.inputs[2]);
        if (W.dims.size() != 4 || W.dims[2] != W.dims[3]) fail("Conv weights must be [cout,cin/g,k,k]");
        if (W.dims[0] < 1 || W.dims[0] > (1 << 20) || W.dims[1] < 1 || W.dims[1] > (1 << 20) || W.dims[2] < 1 || W.dims[2] > 15)
            fail("Conv '" + nd.name + "': weight shape out of range");
        check_weight_range(W, nd.name);
        PlanOp op;
        op.name = nd.name;
        op.cout = (int)W.dims[0];
        int cpg = (int)W.dims[1];
        op.k = (int)W.dims[2];
        int group = (int)nd.attr_i("group", 1);
        auto* st = nd.attr("strides");
        auto* pd = nd.attr("pads");
        auto* dl = nd.attr("dilations");
        op.stride = st && !st->ints.empty() ? (int)st->ints[0] : 1;
        if (st && st->ints.size() == 2 && st->ints[0] != st->ints[1]) fail("Conv strides must be square");
        op.pad = pd && !pd->ints.empty() ? (int)pd->ints[0] : 0;
        if (pd) for (auto p : pd->ints) if (p != op.pad) fail("Conv pads must be symmetric");
        if (dl) for (auto d : dl->ints) if (d != 1) fail("Conv dilation must be 1");
        op.cin = cpg * group;

        int t = lookup(nd.inputs[0]);
        if (plan.tensors[t].scale_base >= 0) {
            op.in = plan.tensors[t].scale_base;
            op.in_scale = plan.tensors[t].scale_vec;
        } else op.in = t;
        const TensorInfo& ti = plan.tensors[op.in];
        if (ti.C != op.cin) fail("Conv '" + nd.name + "': input channels mismatch");
        op.hin = ti.H; op.win = ti.W;
        op.hout = (ti.H + 2 * op.pad - op.k) / op.stride + 1;
        op.wout = (ti.W + 2 * op.pad - op.k) / op.stride + 1;
        if ((int)Bv.numel() != op.cout) fail("Conv bias size mismatch");
        op.bias.assign(Bv.f32(), Bv.f32() + op.cout);
        const float* w = W.f32();
        const int k2 = op.k * op.k;
        if (group == 1) {
            op.kind = (ti.H == 1 && ti.W == 1 && op.k == 1) ? OP_LINEAR : OP_CONV;
            op.ldw = (op.cout + 3) / 4 * 4;
            op.weight.assign((size_t)k2 * op.cin * op.ldw, 0.f);
            for (int co = 0; co < op.cout; ++co)
                for (int ci = 0; ci < op.cin; ++ci)
                    for (int kk = 0; kk < k2; ++kk)
                        op.weight[((size_t)kk * op.cin + ci) * op.ldw + co] = w[((size_t)co * op.cin + ci) * k2 + kk];
        } else {
            if (group != op.cin || op.cout != op.cin || cpg != 1) fail("grouped Conv must be depthwise");
            if (op.in_scale >= 0) fail("depthwise Conv cannot take a gated input");
            op.kind = OP_DWCONV;
            op.ldw = op.cout;
            op.weight.assign((size_t)k2 * op.cout, 0.f);
            for (int c = 0; c < op.cout; ++c)
                for (int kk = 0; kk < k2; ++kk) op.weight[(size_t)kk * op.cout + c] = w[(size_t)c * k2 + kk];
        }
        std::string val = fold_activation(nd.outputs[0], op.act);
        val = fold_residual(val, op.residual);
        op.out = new_tensor(val, op.cout, op.hout, op.wout);
        if (op.residual >= 0) {
            const TensorInfo& r = plan.tensors[op.residual];
            if (r.C != op.cout || r.H != op.hout || r.W != op.wout) fail("residual shape mismatch at '" + nd.name + "'");
        }
        plan.ops.push_back(std::move(op));
    }

    void match_gemm(int n) {
        const OnnxNode& nd = m.nodes[n];
        if (nd.inputs.size() != 3) fail("Gemm must have a bias");
        if (nd.attr_i("transA", 0) != 0) fail("Gemm transA unsupported");
        auto* al = nd.attr("alpha");
        auto* be = nd.attr("beta");
        if ((al && al->f != 1.f) || (be && be->f != 1.f)) fail("Gemm alpha/beta must be 1");
        bool transB = nd.attr_i("transB", 0) != 0;
        const OnnxTensor& W = need_init(nd.inputs[1]);
        const OnnxTensor& Bv = need_init(nd.inputs[2]);
        if (W.dims.size() != 2) fail("Gemm weight must be 2-D");
        if (W.dims[0] < 1 || W.dims[0] > (1 << 24) || W.dims[1] < 1 || W.dims[1] > (1 << 24)) fail("Gemm '" + nd.name + "': weight shape out of range");
        check_weight_range(W, nd.name);
        PlanOp op;
        op.kind = OP_LINEAR;
        op.name = nd.name;
        op.cout = (int)(transB ? W.dims[0] : W.dims[1]);
        op.cin = (int)(transB ? W.dims[1] : W.dims[0]);
        op.in = resolve_plain(nd.inputs[0], "Gemm");
        const TensorInfo& ti = plan.tensors[op.in];
        if (ti.H != 1 || ti.W != 1 || ti.C != op.cin) fail("Gemm input must be a [B, cin] vector");
        op.ldw = (op.cout + 3) / 4 * 4;
        op.weight.assign((size_t)op.cin * op.ldw, 0.f);
        const float* w = W.f32();
        for (int co = 0; co < op.cout; ++co)
            for (int ci = 0; ci < op.cin; ++ci)
                op.weight[(size_t)ci * op.ldw + co] = transB ? w[(size_t)co * op.cin + ci] : w[(size_t)ci * op.cout + co];
        if ((int)Bv.numel() != op.cout) fail("Gemm bias size mismatch");
        op.bias.assign(Bv.f32(), Bv.f32() + op.cout);
        std::string val = fold_activation(nd.outputs[0], op.act);
        op.out = new_tensor(val, op.cout, 1, 1);
        plan.ops.push_back(std::move(op));
    }

    void match_outputs() {
        std::vector<std::vector<int64_t>> shapes;
        for (auto& o : m.outputs) {
            OutputInfo oi;
            oi.name = o.name;
            oi.dims = o.dims;
            oi.tensor = lookup(o.name);
            if (plan.tensors[oi.tensor].scale_base >= 0) fail("graph output '" + o.name + "' is a gated tensor");
            plan.outputs.push_back(oi);
            shapes.push_back(o.dims);
        }
    }
};

std::string debug_i64_slice(const std::vector<int64_t>& v) {   // Rust {:?} of &[i64]
    std::ostringstream os;
    os << "[";
    for (size_t i = 0; i < v.size(); ++i) os << (i ? ", " : "") << v[i];
    os << "]";
    return os.str();
}

const char* model_type_debug(int t) {
    return t == MT_BIRDNET_V24 ? "BirdNetV24" : t == MT_BIRDNET_V30 ? "BirdNetV30" : "PerchV2";
}

}  // namespace

bool detect_model_type(const std::vector<int64_t>& input_shape,
                       const std::vector<std::vector<int64_t>>& output_shapes, int override_type,
                       int* model_type, int* sample_count, int* num_species, int* embedding_dim,
                       std::string* reason) {
    // extract_sample_count  (detection.rs:149-163)
    int64_t sc;
    if (input_shape.size() == 2) sc = input_shape[1];
    else if (input_shape.size() == 3) sc = input_shape[2];
    else { *reason = "unexpected input shape: " + debug_i64_slice(input_shape); return false; }
    if (sc < 0) { *reason = "invalid sample count: " + std::to_string(sc); return false; }
    auto last_dim = [&](const std::vector<int64_t>& s, int* out) -> bool {   // detection.rs:166-174
        if (s.empty()) { *reason = "empty output shape"; return false; }
        if (s.back() < 0) { *reason = "invalid dimension: " + std::to_string(s.back()); return false; }
        *out = (int)s.back();
        return true;
    };
    const size_t nout = output_shapes.size();
    *embedding_dim = 0;
    if (override_type >= 0) {                                              // detection.rs:83-145
        static const int64_t expected[3] = {144000, 160000, 160000};
        if (override_type > 2) { *reason = "unknown model type override"; return false; }
        if (sc != expected[override_type]) {
            *reason = std::string("model type ") + model_type_debug(override_type) + " expects " +
                      std::to_string(expected[override_type]) + " samples, but model has " + std::to_string(sc);
            return false;
        }
        if (override_type == MT_BIRDNET_V24) {
            if (nout != 1) { *reason = "`BirdNET` v2.4 expects 1 output, got " + std::to_string(nout); return false; }
            if (!last_dim(output_shapes[0], num_species)) return false;
        } else if (override_type == MT_BIRDNET_V30) {
            if (nout != 2) { *reason = "`BirdNET` v3.0 expects 2 outputs, got " + std::to_string(nout); return false; }
            if (!last_dim(output_shapes[0], embedding_dim)) return false;
            if (!last_dim(output_shapes[1], num_species)) return false;
        } else {
            if (nout != 4) { *reason = "`Perch` v2 expects 4 outputs, got " + std::to_string(nout); return false; }
            if (!last_dim(output_shapes[0], embedding_dim)) return false;
            if (!last_dim(output_shapes[3], num_species)) return false;
        }
        *model_type = override_type;
        *sample_count = (int)sc;
        return true;
    }
    if (sc == 144000 && nout == 1) {
        if (!last_dim(output_shapes[0], num_species)) return false;
        *model_type = MT_BIRDNET_V24;
    } else if (sc == 160000 && nout == 2) {
        if (!last_dim(output_shapes[0], embedding_dim)) return false;
        if (!last_dim(output_shapes[1], num_species)) return false;
        *model_type = MT_BIRDNET_V30;
    } else if (sc == 160000 && nout == 4) {
        if (!last_dim(output_shapes[0], embedding_dim)) return false;
        if (!last_dim(output_shapes[3], num_species)) return false;
        *model_type = MT_PERCH_V2;
    } else {
        *reason = "unsupported model: " + std::to_string(sc) + " samples, " + std::to_string(nout) +
                  " outputs (expected 144000/1, 160000/2, or 160000/4)";
        return false;
    }
    *sample_count = (int)sc;
    return true;
}

void build_plan(const OnnxModel& m, Plan& plan) {
    Matcher mt(m, plan);
    mt.match_frontend();
    mt.match_cnn();
    mt.match_outputs();
}

}  // namespace bn
