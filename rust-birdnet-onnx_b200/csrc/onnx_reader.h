// Minimal ONNX ModelProto reader (no libprotobuf): walks the protobuf wire format and
// extracts graph nodes, initializers, and input/output value infos.
//
// Replaces, for this engine, what `Session::builder()...commit_from_file(path)` does inside
// ONNX Runtime for the reference (reference: src/classifier.rs:340-350) and what
// `session.inputs()/outputs()` expose (src/classifier.rs:387-420).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace bn {

struct OnnxTensor {
    std::string name;
    std::vector<int64_t> dims;
    int data_type = 0;              // 1 = FLOAT, 7 = INT64
    const uint8_t* raw = nullptr;   // points into the file buffer owned by OnnxModel
    size_t raw_len = 0;
    std::vector<float> f32_fallback;    // when the tensor used float_data instead of raw_data
    std::vector<int64_t> i64_fallback;

    size_t numel() const {
        size_t n = 1;
        for (auto d : dims) n *= (size_t)d;
        return n;
    }
    // numel() contiguous, 4-byte-aligned floats (parse_tensor guarantees exactly numel() values exist)
    const float* f32() const;
    int64_t i64(size_t i = 0) const;
    float f32_at(size_t i = 0) const;
};

struct OnnxAttr {
    std::string name;
    int64_t i = 0;
    float f = 0.f;
    std::string s;
    std::vector<int64_t> ints;
    std::vector<float> floats;
    bool has_t = false;
    OnnxTensor t;                   // TENSOR attribute (Constant nodes)
};

struct OnnxNode {
    std::string op, name;
    std::vector<std::string> inputs, outputs;
    std::vector<OnnxAttr> attrs;
    const OnnxAttr* attr(const std::string& n) const {
        for (auto& a : attrs)
            if (a.name == n) return &a;
        return nullptr;
    }
    int64_t attr_i(const std::string& n, int64_t def) const {
        auto* a = attr(n);
        return a ? a->i : def;
    }
};

struct OnnxValueInfo {
    std::string name;
    std::vector<int64_t> dims;   // -1 for symbolic / dynamic
};

struct OnnxModel {
    std::vector<uint8_t> file;   // owns the bytes every raw tensor points into
    std::vector<OnnxNode> nodes;
    std::map<std::string, OnnxTensor> initializers;
    std::vector<OnnxValueInfo> inputs, outputs;
    int64_t opset = 0;
    int64_t ir_version = 0;
    std::string graph_name;

    const OnnxTensor* init(const std::string& n) const {
        auto it = initializers.find(n);
        return it == initializers.end() ? nullptr : &it->second;
    }
};

// Throws std::runtime_error with a descriptive message on malformed input.
void load_onnx(const std::string& path, OnnxModel& out);
void parse_onnx(OnnxModel& m);   // parses m.file in place

}  // namespace bn
