// Layer plan: the ONNX graph pattern-matched into the engine's fused stages.
//
// The reference never sees inside the model: ONNX Runtime's graph optimiser does Conv+BN+
// activation fusion for it.  Here the same role is played by
// build_plan(): it walks the node list once and emits
//   front-end  (normalise -> framed windowed DFT -> mel -> compress | log)      rows A7/A9
//   CONV / DWCONV / LINEAR ops with fused bias + SiLU/sigmoid + residual + SE gate   row A8
//   GAP ops (squeeze-excite pooling, global pool)
// plus the output map the reference relies on (src/classifier.rs:917-934).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "onnx_reader.h"

namespace bn {

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_SIGMOID = 2 };
enum FrontEndKind : int { FE_BIRDNET_V24 = 0, FE_LOGMEL = 1 };
enum ModelTypeId : int { MT_BIRDNET_V24 = 0, MT_BIRDNET_V30 = 1, MT_PERCH_V2 = 2 };

struct SpecBranch {
    int n_fft = 0, hop = 0, n_bins = 0, n_mels = 0, n_frames = 0;
    float exponent = 1.f;              // v24: pow(mel^2, exponent)
    bool flip = false;
    std::vector<float> window;         // [n_fft]
    std::vector<float> mel;            // [n_bins][n_mels] as in the file (not flipped)
};

struct FrontEndPlan {
    int kind = FE_BIRDNET_V24;
    int sample_count = 0;
    int pad_end = 0;
    bool normalize = false;
    float eps = 1e-6f, half = 0.5f, two = 2.0f;
    float log_floor = 0.f, log_scale = 1.f;
    std::vector<SpecBranch> branches;
    int out_tensor = -1;               // id of "spec"
    int spectrogram_tensor = -1;       // logmel: [frames][mels] alias (Perch output 2)
};

struct TensorInfo {
    std::string name;
    int C = 0, H = 0, W = 0;           // per-segment NHWC extents
    int alias_of = -1;                 // shares storage with another tensor
    int scale_base = -1, scale_vec = -1;   // virtual tensor: base (*) per-channel gate
    size_t elems() const { return (size_t)C * H * W; }
};

enum OpKind : int { OP_CONV = 0, OP_DWCONV = 1, OP_LINEAR = 2, OP_GAP = 3 };

struct PlanOp {
    int kind = OP_CONV;
    std::string name;
    int in = -1, out = -1;
    int in_scale = -1;                 // [C] gate multiplied into the input (SE)
    int residual = -1;                 // tensor added after the activation
    int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
    int hin = 1, win = 1, hout = 1, wout = 1;
    int act = ACT_NONE;
    int ldw = 0;                       // padded cout (multiple of 4) of the engine weight layout
    std::vector<float> weight;         // CONV/LINEAR: [k*k*cin][ldw]; DWCONV: [k*k][cout]
    std::vector<float> bias;           // [cout]
    uint64_t macs() const {
        if (kind == OP_GAP) return 0;
        uint64_t px = (uint64_t)hout * wout;
        if (kind == OP_DWCONV) return px * cout * k * k;
        return px * cout * (uint64_t)cin * k * k;
    }
};

struct OutputInfo {
    std::string name;
    int tensor = -1;
    std::vector<int64_t> dims;         // as declared in the file (-1 = dynamic)
};

struct Plan {
    FrontEndPlan fe;
    std::vector<TensorInfo> tensors;
    std::vector<PlanOp> ops;
    std::vector<OutputInfo> outputs;
    std::string input_name;
    std::vector<int64_t> input_dims;
    // derived by detect_model_type()  (reference: src/detection.rs:15-80)
    int model_type = -1;
    int sample_count = 0, num_species = 0, embedding_dim = 0;
    int logits_tensor = -1, embedding_tensor = -1;

    int root(int t) const {
        while (tensors[t].alias_of >= 0) t = tensors[t].alias_of;
        return t;
    }
};

// Throws std::runtime_error("...") when the graph uses something the engine cannot run.
void build_plan(const OnnxModel& m, Plan& plan);

// Reference: src/detection.rs:15-174.  override_type < 0 means auto-detect.  On failure
// returns false and fills `reason` with the reference's exact message text.
bool detect_model_type(const std::vector<int64_t>& input_shape,
                       const std::vector<std::vector<int64_t>>& output_shapes, int override_type,
                       int* model_type, int* sample_count, int* num_species, int* embedding_dim,
                       std::string* reason);

}  // namespace bn
