// Reader of frozen TensorFlow graphs (binary GraphDef), host code only — no TensorFlow, no protobuf library.
// Replaces the weight-loading half of `load_graph` (reference integration_prediction_neural_network.cpp:29-69): after
// `freeze_graph` (freezing_graph_pnn.py:129-139) every variable is a `Const` node named like the variable, its value a
// TensorProto.  Field numbers: GraphDef.node = 1; NodeDef.name = 1, .op = 2, .attr = 5 (map entry key = 1, value = 2);
// AttrValue.tensor = 8; TensorProto.dtype = 1, .tensor_shape = 2 (dim = 2 {size = 1}), .tensor_content = 4, .float_val = 5.
#include "pnn_internal.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace pnn {
namespace {

struct Span {
    const uint8_t* p;
    size_t n;
};

struct Field {
    uint32_t number;
    int wire;
    uint64_t value;   // wire 0
    Span bytes;       // wire 1, 2, 5
};

uint64_t varint(const Span& s, size_t* pos) {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
        if (*pos >= s.n) throw std::runtime_error("truncated varint");
        const uint8_t b = s.p[(*pos)++];
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return v;
    }
    throw std::runtime_error("varint too long");
}

bool next_field(const Span& s, size_t* pos, Field* f) {
    if (*pos >= s.n) return false;
    const uint64_t tag = varint(s, pos);
    f->number = (uint32_t)(tag >> 3);
    f->wire = (int)(tag & 7);
    size_t len = 0;
    switch (f->wire) {
        case 0: f->value = varint(s, pos); return true;
        case 1: len = 8; break;
        case 2: len = (size_t)varint(s, pos); break;
        case 5: len = 4; break;
        default: throw std::runtime_error("unsupported protobuf wire type " + std::to_string(f->wire));
    }
    if (len > s.n - *pos) throw std::runtime_error("truncated protobuf field");
    f->bytes = Span{s.p + *pos, len};
    *pos += len;
    return true;
}

// -> false when the tensor is not float32
bool parse_tensor(const Span& t, std::vector<float>* values, std::vector<int>* dims) {
    uint64_t dtype = 0;
    Span content{nullptr, 0};
    std::vector<float> float_val;
    size_t pos = 0;
    Field f;
    while (next_field(t, &pos, &f)) {
        if (f.number == 1 && f.wire == 0) dtype = f.value;
        else if (f.number == 2 && f.wire == 2) {
            size_t p2 = 0;
            Field d;
            while (next_field(f.bytes, &p2, &d)) {
                if (d.number != 2 || d.wire != 2) continue;
                size_t p3 = 0;
                Field sz;
                int size = 0;
                while (next_field(d.bytes, &p3, &sz)) {
                    if (sz.number == 1 && sz.wire == 0) size = (int)sz.value;
                }
                dims->push_back(size);
            }
        } else if (f.number == 4 && f.wire == 2) content = f.bytes;
        else if (f.number == 5 && f.wire == 2) {
            for (size_t i = 0; i + 4 <= f.bytes.n; i += 4) {
                float v;
                memcpy(&v, f.bytes.p + i, 4);
                float_val.push_back(v);
            }
        } else if (f.number == 5 && f.wire == 5) {
            float v;
            memcpy(&v, f.bytes.p, 4);
            float_val.push_back(v);
        }
    }
    if (dtype != 1) return false;   // DT_FLOAT
    size_t count = 1;
    for (int d : *dims) {
        if (d < 0) throw std::runtime_error("negative tensor dimension");
        count *= (size_t)d;
    }
    if (content.n) {
        if (content.n != count * 4) throw std::runtime_error("tensor_content does not match the tensor shape");
        values->resize(count);
        memcpy(values->data(), content.p, content.n);
    } else if (float_val.size() == count) {
        *values = float_val;
    } else if (float_val.size() == 1) {
        values->assign(count, float_val[0]);   // a constant-filled tensor is stored as one value
    } else if (float_val.empty()) {
        values->assign(count, 0.f);
    } else {
        throw std::runtime_error("float_val does not match the tensor shape");
    }
    return true;
}

}  // namespace

void read_frozen_graph(const std::vector<char>& data, const std::string& path, FlatFile* out) {
    const Span g{reinterpret_cast<const uint8_t*>(data.data()), data.size()};
    if (g.n == 0 || g.p[0] != 0x0a) throw std::runtime_error("not a binary GraphDef");
    size_t pos = 0;
    Field node;
    const std::string fc_prefix = "fully_connected/", conv_prefix = "convolutional/";
    while (next_field(g, &pos, &node)) {
        if (node.number != 1 || node.wire != 2) continue;   // versions, library
        std::string name, op;
        Span tensor{nullptr, 0};
        size_t p2 = 0;
        Field f;
        while (next_field(node.bytes, &p2, &f)) {
            if (f.wire != 2) continue;
            if (f.number == 1) name.assign(reinterpret_cast<const char*>(f.bytes.p), f.bytes.n);
            else if (f.number == 2) op.assign(reinterpret_cast<const char*>(f.bytes.p), f.bytes.n);
            else if (f.number == 5) {
                std::string key;
                Span attr{nullptr, 0};
                size_t p3 = 0;
                Field e;
                while (next_field(f.bytes, &p3, &e)) {
                    if (e.wire != 2) continue;
                    if (e.number == 1) key.assign(reinterpret_cast<const char*>(e.bytes.p), e.bytes.n);
                    else if (e.number == 2) attr = e.bytes;
                }
                if (key == "value" && attr.p) {
                    size_t p4 = 0;
                    Field a;
                    while (next_field(attr, &p4, &a)) {
                        if (a.number == 8 && a.wire == 2) tensor = a.bytes;
                    }
                }
            }
        }
        if (op != "Const" || !tensor.p) continue;
        if (name.compare(0, fc_prefix.size(), fc_prefix) != 0 && name.compare(0, conv_prefix.size(), conv_prefix) != 0) continue;
        std::vector<float> values;
        std::vector<int> dims;
        if (!parse_tensor(tensor, &values, &dims)) continue;
        out->t[name] = std::move(values);
        out->shape[name] = dims;
    }
    // width and kind from the constants (reference PredictionNeuralNetwork.py:126-132: strides per width)
    auto fc0 = out->shape.find("fully_connected/weights_0");
    if (fc0 != out->shape.end() && fc0->second.size() == 2) {
        const int w = (int)std::lround(std::sqrt(fc0->second[0] / 5.0));
        if (5 * w * w != fc0->second[0]) throw std::runtime_error("fully_connected/weights_0 has " + std::to_string(fc0->second[0]) + " rows");
        out->width = w;
        out->is_fc = true;
        return;
    }
    int n_layers = 0;
    while (out->shape.count("convolutional/merger/transpose_convolution_" + std::to_string(n_layers) + "/weights")) ++n_layers;
    auto kernel_of = [&](int i) {
        auto it = out->shape.find("convolutional/branch_above/convolution_" + std::to_string(i) + "/weights");
        return it == out->shape.end() || it->second.size() != 4 ? 0 : it->second[0];
    };
    out->is_fc = false;
    if (n_layers == 2) out->width = kernel_of(0) == 3 ? 4 : 8;          // strides (1, 1) / (2, 1)
    else if (n_layers == 4) out->width = 16;                            // (2, 1, 2, 1)
    else if (n_layers == 5) out->width = kernel_of(2) == 3 ? 32 : 64;   // (2, 2, 1, 2, 1) / (2, 2, 2, 2, 1)
    else throw std::runtime_error("no PNN constants found in \"" + path + "\"");
}

}  // namespace pnn
