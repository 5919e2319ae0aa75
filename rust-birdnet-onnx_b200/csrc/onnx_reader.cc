// See onnx_reader.h.  Field numbers are those of onnx.proto3 (IR version 8, opset 17).
#include "onnx_reader.h"

#include <cerrno>
#include <climits>
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

// The file is untrusted: every accessor checks the wire type before it touches f.data / f.varint.
std::string str(const Field& f) {
    if (f.wt != 2) throw std::runtime_error("onnx: string field " + std::to_string(f.num) + " has wire type " + std::to_string(f.wt));
    return std::string((const char*)f.data, f.len);
}
float f32_of(const Field& f) {
    if (f.wt != 5 || f.len != 4) throw std::runtime_error("onnx: float field " + std::to_string(f.num) + " has wire type " + std::to_string(f.wt));
    float v;
    memcpy(&v, f.data, 4);
    return v;
}
uint64_t varint_of(const Field& f) {
    if (f.wt != 0) throw std::runtime_error("onnx: integer field " + std::to_string(f.num) + " has wire type " + std::to_string(f.wt));
    return f.varint;
}
void packed_floats(const Field& f, std::vector<float>& out) {
    if (f.wt == 2) {
        if (f.len % 4) throw std::runtime_error("onnx: packed float field length is not a multiple of 4");
        for (size_t i = 0; i < f.len; i += 4) { float v; memcpy(&v, f.data + i, 4); out.push_back(v); }
    } else out.push_back(f32_of(f));
}
void packed_ints(const Field& f, std::vector<int64_t>& out) {
    if (f.wt == 2) { Cursor q{f.data, f.data + f.len}; while (!q.done()) out.push_back((int64_t)read_varint(q)); }
    else out.push_back((int64_t)varint_of(f));
}
constexpr uint64_t kMaxElems = 1ull << 32;      // no tensor of a model this engine can run is larger

void parse_tensor(const uint8_t* p, size_t len, OnnxTensor& t) {
    Cursor c{p, p + len};
    Field f;
    while (next_field(c, f)) {
        switch (f.num) {
            case 1: packed_ints(f, t.dims); break;          // dims (possibly packed)
            case 2: t.data_type = (int)varint_of(f); break;
            case 4: packed_floats(f, t.f32_fallback); break;   // float_data
            case 7: packed_ints(f, t.i64_fallback); break;     // int64_data
            case 8: t.name = str(f); break;
            case 9:
                if (f.wt != 2) throw std::runtime_error("onnx: raw_data has wire type " + std::to_string(f.wt));
                t.raw = f.data; t.raw_len = f.len;
                break;
            default: break;
        }
    }
    if (t.data_type != 1 && t.data_type != 7)
        throw std::runtime_error("onnx: initializer '" + t.name + "' has unsupported data_type " + std::to_string(t.data_type));
    // element count: non-negative dims, no overflow, bounded
    uint64_t n = 1;
    for (auto d : t.dims) {
        if (d < 0) throw std::runtime_error("onnx: initializer '" + t.name + "' has a negative dimension");
        if (d != 0 && n > kMaxElems / (uint64_t)d) throw std::runtime_error("onnx: initializer '" + t.name + "' is too large");
        n *= (uint64_t)d;
    }
    const size_t esz = t.data_type == 1 ? 4 : 8;
    // exactly one payload of exactly numel elements: every later read of numel() elements is in bounds
    if (t.raw) {
        if (t.raw_len != n * esz) throw std::runtime_error("onnx: initializer '" + t.name + "' raw_data size mismatch");
        if (reinterpret_cast<uintptr_t>(t.raw) % esz) {          // protobuf gives no alignment: move to aligned storage
            if (t.data_type == 1) { t.f32_fallback.resize(n); memcpy(t.f32_fallback.data(), t.raw, n * esz); }
            else { t.i64_fallback.resize(n); memcpy(t.i64_fallback.data(), t.raw, n * esz); }
            t.raw = nullptr;
            t.raw_len = 0;
        }
    } else if ((t.data_type == 1 ? t.f32_fallback.size() : t.i64_fallback.size()) != n) {
        throw std::runtime_error("onnx: initializer '" + t.name + "' holds " +
                                 std::to_string(t.data_type == 1 ? t.f32_fallback.size() : t.i64_fallback.size()) +
                                 " values for " + std::to_string(n) + " elements");
    }
}

void parse_attr(const uint8_t* p, size_t n, OnnxAttr& a) {
    Cursor c{p, p + n};
    Field f;
    while (next_field(c, f)) {
        switch (f.num) {
            case 1: a.name = str(f); break;
            case 2: a.f = f32_of(f); break;
            case 3: a.i = (int64_t)varint_of(f); break;
            case 4: a.s = str(f); break;
            case 5:                                   // t: TensorProto (Constant nodes of exporters that do not use initializers)
                if (f.wt != 2) throw std::runtime_error("onnx: tensor attribute has wire type " + std::to_string(f.wt));
                parse_tensor(f.data, f.len, a.t);
                a.has_t = true;
                break;
            case 7: packed_floats(f, a.floats); break;
            case 8: packed_ints(f, a.ints); break;
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
            case 5: {
                if (f.wt != 2) throw std::runtime_error("onnx: attribute has wire type " + std::to_string(f.wt));
                OnnxAttr a; parse_attr(f.data, f.len, a); nd.attrs.push_back(std::move(a)); break;
            }
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
            case 1: case 5: case 11: case 12:
                if (f.wt != 2) throw std::runtime_error("onnx: graph field " + std::to_string(f.num) + " has wire type " + std::to_string(f.wt));
                break;
            default: break;
        }
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

// Shape arithmetic that exporters leave in the graph (torch lowers F.pad's amounts to ConstantOfShape / Concat / Reshape /
// Slice / Transpose / Cast over tiny int64 tensors): nodes of that vocabulary whose inputs are all constants are evaluated
// here and become initializers, as ONNX Runtime's constant folding does.  Tensors of rank <= 2 and <= 64 elements only.
struct SmallI64 { std::vector<int64_t> dims, v; };
bool small_i64(const OnnxModel& m, const std::string& name, SmallI64& out) {
    const OnnxTensor* t = m.init(name);
    if (!t || t->data_type != 7 || t->dims.size() > 2 || t->numel() > 64) return false;
    out.dims = t->dims;
    out.v.resize(t->numel());
    for (size_t i = 0; i < out.v.size(); ++i) out.v[i] = t->i64(i);
    return true;
}
void fold_int64_constants(OnnxModel& m) {
    bool changed = true;
    while (changed) {
        changed = false;
        for (size_t n = 0; n < m.nodes.size(); ++n) {
            const OnnxNode& nd = m.nodes[n];
            if (nd.outputs.size() != 1) continue;
            SmallI64 r;
            bool ok = false;
            if (nd.op == "ConstantOfShape" && nd.inputs.size() == 1) {
                SmallI64 sh;
                const OnnxAttr* v = nd.attr("value");
                if (small_i64(m, nd.inputs[0], sh) && sh.dims.size() <= 1 && sh.v.size() <= 2 && v && v->has_t && v->t.data_type == 7 && v->t.numel() == 1) {
                    size_t cnt = 1;
                    for (auto d : sh.v) { if (d < 0 || d > 64) { cnt = 0; break; } cnt *= (size_t)d; }
                    if (cnt <= 64) { r.dims = sh.v; r.v.assign(cnt, v->t.i64(0)); ok = true; }
                }
            } else if (nd.op == "Concat" && nd.attr_i("axis", 0) == 0 && !nd.inputs.empty()) {
                ok = true;
                for (auto& in : nd.inputs) {
                    SmallI64 a;
                    if (!small_i64(m, in, a) || a.dims.size() > 1) { ok = false; break; }
                    r.v.insert(r.v.end(), a.v.begin(), a.v.end());
                }
                r.dims = {(int64_t)r.v.size()};
            } else if (nd.op == "Reshape" && nd.inputs.size() == 2) {
                SmallI64 a, sh;
                if (small_i64(m, nd.inputs[0], a) && small_i64(m, nd.inputs[1], sh) && sh.dims.size() <= 1 && sh.v.size() <= 2) {
                    int64_t known = 1, infer = -1;
                    for (size_t i = 0; i < sh.v.size(); ++i) { if (sh.v[i] == -1) infer = (int64_t)i; else known *= sh.v[i]; }
                    r.dims = sh.v;
                    if (infer >= 0 && known > 0) r.dims[infer] = (int64_t)a.v.size() / known;
                    int64_t tot = 1;
                    for (auto d : r.dims) tot *= d;
                    if (tot == (int64_t)a.v.size()) { r.v = a.v; ok = true; }
                }
            } else if (nd.op == "Transpose") {
                SmallI64 a;
                const OnnxAttr* pm = nd.attr("perm");
                if (small_i64(m, nd.inputs[0], a) && a.dims.size() == 2 && pm && pm->ints == std::vector<int64_t>{1, 0}) {
                    const int64_t R = a.dims[0], Cc = a.dims[1];
                    r.dims = {Cc, R};
                    r.v.resize(a.v.size());
                    for (int64_t i = 0; i < R; ++i) for (int64_t j = 0; j < Cc; ++j) r.v[j * R + i] = a.v[i * Cc + j];
                    ok = true;
                }
            } else if (nd.op == "Slice" && nd.inputs.size() >= 3) {
                SmallI64 a, st, en, ax, sp;
                if (small_i64(m, nd.inputs[0], a) && a.dims.size() >= 1 && small_i64(m, nd.inputs[1], st) && small_i64(m, nd.inputs[2], en) &&
                    st.v.size() == 1 && en.v.size() == 1) {
                    int64_t axis = 0, step = 1;
                    if (nd.inputs.size() >= 4 && !nd.inputs[3].empty()) { if (!small_i64(m, nd.inputs[3], ax) || ax.v.size() != 1) continue; axis = ax.v[0]; }
                    if (nd.inputs.size() >= 5 && !nd.inputs[4].empty()) { if (!small_i64(m, nd.inputs[4], sp) || sp.v.size() != 1) continue; step = sp.v[0]; }
                    if (axis < 0) axis += (int64_t)a.dims.size();
                    if (axis == 0 && (step == 1 || step == -1)) {
                        const int64_t R = a.dims[0], Cc = a.dims.size() == 2 ? a.dims[1] : 1;
                        auto clampi = [&](int64_t x, int64_t lo, int64_t hi) { return x < lo ? lo : (x > hi ? hi : x); };
                        int64_t s0 = st.v[0], e0 = en.v[0];
                        if (s0 < 0) s0 += R;
                        if (e0 < 0 && e0 > INT64_MIN / 2) e0 += R;
                        std::vector<int64_t> rows;
                        if (step == 1) { s0 = clampi(s0, 0, R); e0 = clampi(e0, 0, R); for (int64_t i = s0; i < e0; ++i) rows.push_back(i); }
                        else { s0 = clampi(s0, -1, R - 1); e0 = clampi(e0, -1, R - 1); for (int64_t i = s0; i > e0; --i) rows.push_back(i); }
                        for (auto i : rows) r.v.insert(r.v.end(), a.v.begin() + i * Cc, a.v.begin() + (i + 1) * Cc);
                        r.dims = a.dims;
                        r.dims[0] = (int64_t)rows.size();
                        ok = true;
                    }
                }
            } else if (nd.op == "Cast" && nd.attr_i("to", 0) == 7) {
                ok = small_i64(m, nd.inputs[0], r);
            }
            if (!ok) continue;
            OnnxTensor t;
            t.name = nd.outputs[0];
            t.data_type = 7;
            t.dims = r.dims;
            t.i64_fallback = r.v;
            m.initializers[t.name] = std::move(t);
            m.nodes.erase(m.nodes.begin() + (long)n);
            changed = true;
            break;
        }
    }
}

}  // namespace

int64_t OnnxTensor::i64(size_t i) const {
    if (data_type != 7) throw std::runtime_error("onnx: tensor '" + name + "' is not int64");
    if (i >= numel()) throw std::runtime_error("onnx: index " + std::to_string(i) + " out of range for tensor '" + name + "'");
    if (raw) { int64_t v; memcpy(&v, raw + 8 * i, 8); return v; }
    return i64_fallback.at(i);
}

float OnnxTensor::f32_at(size_t i) const {
    if (data_type != 1) throw std::runtime_error("onnx: tensor '" + name + "' is not float");
    if (i >= numel()) throw std::runtime_error("onnx: index " + std::to_string(i) + " out of range for tensor '" + name + "'");
    if (raw) { float v; memcpy(&v, raw + 4 * i, 4); return v; }
    return f32_fallback[i];
}

const float* OnnxTensor::f32() const {
    if (data_type != 1) throw std::runtime_error("onnx: tensor '" + name + "' is not float");
    return raw ? reinterpret_cast<const float*>(raw) : f32_fallback.data();   // raw is 4-byte aligned (parse_tensor)
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
                else if (f2.num == 2) ver = (int64_t)varint_of(f2);
            }
            if (domain.empty() || domain == "ai.onnx") m.opset = ver;
        }
    }
    if (!have_graph) throw std::runtime_error("onnx: no graph in model file");
    // Constant nodes (torch's exporter emits scalars and small tables this way) become initializers under their output
    // name, so the planner sees one dialect: constants are always initializers
    {
        std::vector<OnnxNode> kept;
        kept.reserve(m.nodes.size());
        for (auto& nd : m.nodes) {
            if (nd.op == "Constant" && nd.outputs.size() == 1) {
                const OnnxAttr* v = nd.attr("value");
                OnnxTensor t;
                if (v && v->has_t) t = v->t;
                else if (const OnnxAttr* vf = nd.attr("value_float")) { t.data_type = 1; t.f32_fallback = {vf->f}; }
                else if (const OnnxAttr* vi = nd.attr("value_int")) { t.data_type = 7; t.i64_fallback = {vi->i}; }
                else if (const OnnxAttr* vfs = nd.attr("value_floats")) { t.data_type = 1; t.f32_fallback = vfs->floats; t.dims = {(int64_t)vfs->floats.size()}; }
                else if (const OnnxAttr* vis = nd.attr("value_ints")) { t.data_type = 7; t.i64_fallback = vis->ints; t.dims = {(int64_t)vis->ints.size()}; }
                else throw std::runtime_error("onnx: Constant node '" + nd.name + "' has no supported value attribute");
                t.name = nd.outputs[0];
                m.initializers[t.name] = std::move(t);
                continue;
            }
            kept.push_back(std::move(nd));
        }
        m.nodes.swap(kept);
    }
    fold_int64_constants(m);
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
    long sz = ftell(fp);           // a directory reports LONG_MAX / -1 here
    fseek(fp, 0, SEEK_SET);
    if (sz <= 0 || (unsigned long)sz > (1ul << 34)) { fclose(fp); throw std::runtime_error("model file '" + path + "' is empty or not a regular file"); }
    out.file.resize((size_t)sz);
    size_t got = fread(out.file.data(), 1, (size_t)sz, fp);
    fclose(fp);
    if (got != (size_t)sz) throw std::runtime_error("short read on '" + path + "'");
    parse_onnx(out);
}

}  // namespace bn
