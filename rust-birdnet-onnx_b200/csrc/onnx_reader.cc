// See onnx_reader.h.  Field numbers are those of onnx.proto3 (IR version 8, opset 17).
#include "onnx_reader.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace bn {
namespace {

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool done() const { return p >= end; }
};

uint64_t read_varint(Cursor& c) {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
        if (c.p >= c.end) throw std::runtime_error("onnx: truncated varint");
        uint8_t b = *c.p++;
        v |= (uint64_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) return v;
        shift += 7;
        if (shift > 63) throw std::runtime_error("onnx: varint too long");
    }
}

struct Field {
    uint32_t num;
    uint32_t wt;
    uint64_t varint;          // wt 0
    const uint8_t* data;      // wt 1,2,5
    size_t len;
};

bool next_field(Cursor& c, Field& f) {
    if (c.done()) return false;
    uint64_t key = read_varint(c);
    f.num = (uint32_t)(key >> 3);
    f.wt = (uint32_t)(key & 7);
    f.varint = 0;
    f.data = nullptr;
    f.len = 0;
    switch (f.wt) {
        case 0: f.varint = read_varint(c); break;
        case 1:
            if (c.end - c.p < 8) throw std::runtime_error("onnx: truncated fixed64");
            f.data = c.p; f.len = 8; c.p += 8; break;
        case 2: {
            uint64_t n = read_varint(c);
            if ((uint64_t)(c.end - c.p) < n) throw std::runtime_error("onnx: truncated bytes field");
            f.data = c.p; f.len = (size_t)n; c.p += n; break;
        }
        case 5:
            if (c.end - c.p < 4) throw std::runtime_error("onnx: truncated fixed32");
            f.data = c.p; f.len = 4; c.p += 4; break;
        default: throw std::runtime_error("onnx: unsupported wire type");
    }
    return true;
}

std::string str(const Field& f) { return std::string((const char*)f.data, f.len); }
float f32_of(const Field& f) { float v; memcpy(&v, f.data, 4); return v; }

void parse_tensor(const uint8_t* p, size_t n, OnnxTensor& t) {
    Cursor c{p, p + n};
    Field f;
    while (next_field(c, f)) {
        switch (f.num) {
            case 1:   // dims (possibly packed)
                if (f.wt == 2) { Cursor q{f.data, f.data + f.len}; while (!q.done()) t.dims.push_back((int64_t)read_varint(q)); }
                else t.dims.push_back((int64_t)f.varint);
                break;
            case 2: t.data_type = (int)f.varint; break;
            case 4:   // float_data
                if (f.wt == 2) { for (size_t i = 0; i + 4 <= f.len; i += 4) { float v; memcpy(&v, f.data + i, 4); t.f32_fallback.push_back(v); } }
                else t.f32_fallback.push_back(f32_of(f));
                break;
            case 7:   // int64_data
                if (f.wt == 2) { Cursor q{f.data, f.data + f.len}; while (!q.done()) t.i64_fallback.push_back((int64_t)read_varint(q)); }
                else t.i64_fallback.push_back((int64_t)f.varint);
                break;
            case 8: t.name = str(f); break;
            case 9: t.raw = f.data; t.raw_len = f.len; break;
            default: break;
        }
    }
    if (t.data_type != 1 && t.data_type != 7)
        throw std::runtime_error("onnx: initializer '" + t.name + "' has unsupported data_type " + std::to_string(t.data_type));
    size_t esz = t.data_type == 1 ? 4 : 8;
    if (t.raw && t.raw_len != t.numel() * esz)
        throw std::runtime_error("onnx: initializer '" + t.name + "' raw_data size mismatch");
}

void parse_attr(const uint8_t* p, size_t n, OnnxAttr& a) {
    Cursor c{p, p + n};
    Field f;
    while (next_field(c, f)) {
        switch (f.num) {
            case 1: a.name = str(f); break;
            case 2: a.f = f32_of(f); break;
            case 3: a.i = (int64_t)f.varint; break;
            case 4: a.s = str(f); break;
            case 7:
                if (f.wt == 2) { for (size_t i = 0; i + 4 <= f.len; i += 4) { float v; memcpy(&v, f.data + i, 4); a.floats.push_back(v); } }
                else a.floats.push_back(f32_of(f));
                break;
            case 8:
                if (f.wt == 2) { Cursor q{f.data, f.data + f.len}; while (!q.done()) a.ints.push_back((int64_t)read_varint(q)); }
                else a.ints.push_back((int64_t)f.varint);
                break;
            default: break;
        }
    }
}

void parse_node(const uint8_t* p, size_t n, OnnxNode& nd) {
    Cursor c{p, p + n};
    Field f;
    while (next_field(c, f)) {
        switch (f.num) {
            case 1: nd.inputs.push_back(str(f)); break;
            case 2: nd.outputs.push_back(str(f)); break;
            case 3: nd.name = str(f); break;
            case 4: nd.op = str(f); break;
            case 5: { OnnxAttr a; parse_attr(f.data, f.len, a); nd.attrs.push_back(std::move(a)); break; }
            default: break;
        }
    }
}

void parse_value_info(const uint8_t* p, size_t n, OnnxValueInfo& vi) {
    Cursor c{p, p + n};
    Field f;
    while (next_field(c, f)) {
        if (f.num == 1) vi.name = str(f);
        else if (f.num == 2 && f.wt == 2) {              // TypeProto
            Cursor c2{f.data, f.data + f.len};
            Field f2;
            while (next_field(c2, f2)) {
                if (f2.num != 1 || f2.wt != 2) continue;  // tensor_type
                Cursor c3{f2.data, f2.data + f2.len};
                Field f3;
                while (next_field(c3, f3)) {
                    if (f3.num != 2 || f3.wt != 2) continue;   // shape
                    Cursor c4{f3.data, f3.data + f3.len};
                    Field f4;
                    while (next_field(c4, f4)) {
                        if (f4.num != 1 || f4.wt != 2) continue;   // dim
                        int64_t d = -1;
                        Cursor c5{f4.data, f4.data + f4.len};
                        Field f5;
                        while (next_field(c5, f5))
                            if (f5.num == 1 && f5.wt == 0) d = (int64_t)f5.varint;
                        vi.dims.push_back(d);
                    }
                }
            }
        }
    }
}

void parse_graph(const uint8_t* p, size_t n, OnnxModel& m) {
    Cursor c{p, p + n};
    Field f;
    while (next_field(c, f)) {
        switch (f.num) {
            case 1: { OnnxNode nd; parse_node(f.data, f.len, nd); m.nodes.push_back(std::move(nd)); break; }
            case 2: m.graph_name = str(f); break;
            case 5: { OnnxTensor t; parse_tensor(f.data, f.len, t); std::string nm = t.name; m.initializers.emplace(nm, std::move(t)); break; }
            case 11: { OnnxValueInfo vi; parse_value_info(f.data, f.len, vi); m.inputs.push_back(std::move(vi)); break; }
            case 12: { OnnxValueInfo vi; parse_value_info(f.data, f.len, vi); m.outputs.push_back(std::move(vi)); break; }
            default: break;
        }
    }
}

}  // namespace

int64_t OnnxTensor::i64(size_t i) const {
    if (data_type != 7) throw std::runtime_error("onnx: tensor '" + name + "' is not int64");
    if (raw) { int64_t v; memcpy(&v, raw + 8 * i, 8); return v; }
    return i64_fallback.at(i);
}

float OnnxTensor::f32_at(size_t i) const {
    if (data_type != 1) throw std::runtime_error("onnx: tensor '" + name + "' is not float");
    return f32()[i];
}

void parse_onnx(OnnxModel& m) {
    Cursor c{m.file.data(), m.file.data() + m.file.size()};
    Field f;
    bool have_graph = false;
    while (next_field(c, f)) {
        if (f.num == 1 && f.wt == 0) m.ir_version = (int64_t)f.varint;
        else if (f.num == 7 && f.wt == 2) { parse_graph(f.data, f.len, m); have_graph = true; }
        else if (f.num == 8 && f.wt == 2) {
            Cursor c2{f.data, f.data + f.len};
            Field f2;
            std::string domain;
            int64_t ver = 0;
            while (next_field(c2, f2)) {
                if (f2.num == 1) domain = str(f2);
                else if (f2.num == 2) ver = (int64_t)f2.varint;
            }
            if (domain.empty() || domain == "ai.onnx") m.opset = ver;
        }
    }
    if (!have_graph) throw std::runtime_error("onnx: no graph in model file");
    // initializers that are also listed as graph inputs (older exporters) are not real inputs
    std::vector<OnnxValueInfo> real;
    for (auto& vi : m.inputs)
        if (!m.init(vi.name)) real.push_back(vi);
    m.inputs.swap(real);
}

void load_onnx(const std::string& path, OnnxModel& out) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("cannot open '" + path + "': " + strerror(errno));
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (sz <= 0) { fclose(fp); throw std::runtime_error("model file '" + path + "' is empty"); }
    out.file.resize((size_t)sz);
    size_t got = fread(out.file.data(), 1, (size_t)sz, fp);
    fclose(fp);
    if (got != (size_t)sz) throw std::runtime_error("short read on '" + path + "'");
    parse_onnx(out);
}

}  // namespace bn
