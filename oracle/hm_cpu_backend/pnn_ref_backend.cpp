// libpnn_ref: CPU backend behind the C ABI of include/pnn_cuda.h -- TEST / BASELINE INFRASTRUCTURE, not product.
//
// What it is for (SURVEY.md section 8d "CPU baseline beside it"): the reference's HM builds run
// the PNN graphs through the TensorFlow 1.9 C++ CPU runtime inside the codec process
// (hevc/hm_16_15_substitution/source/Lib/TLibCommon/TComPrediction.cpp:556-614, Session::Run on a batch of one).
// TensorFlow cannot be installed here, so the "TF-CPU build" of the codec is stood in for by the SAME unmodified codec
// sources and the SAME link seam (hm/shim/) linked against this library instead of libpnn_cuda: the fp32 oracle of
// oracle/nets.py restated on libtorch-CPU (oneDNN / MKL kernels, all host threads: the same arithmetic class as
// TensorFlow-Eigen), in process, batch of one.  Only the entry points the link seam uses are provided.
// PNN_REF_NULL=1 makes every prediction return zeros at once: the codec's own time (the Amdahl floor of any backend).
//
// Layer semantics follow oracle/nets.py line by line (which cites pnn/components.py, pnn/tfutils.py).
#include "../../include/pnn_cuda.h"

#include <ATen/ATen.h>
#include <ATen/Parallel.h>
#include <c10/core/InferenceMode.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

struct RefNet {
    int width = 0;
    bool is_fc = false;
    std::map<std::string, at::Tensor> t;       // reference layouts, converted to torch layouts at load
    std::vector<int> strides;
};

std::vector<int> strides_branch(int w) {       // pnn/PredictionNeuralNetwork.py:126-132
    switch (w) {
        case 4: return {1, 1};
        case 8: return {2, 1};
        case 16: return {2, 1, 2, 1};
        case 32: return {2, 2, 1, 2, 1};
        case 64: return {2, 2, 2, 2, 1};
    }
    return {};
}

// PNNW flat binary (layout in <pkg>/weights.py)
void read_pnnw(const std::string& path, int* width, int* is_fc, std::map<std::string, at::Tensor>* out, int64_t* n_params,
               double* checksum) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open weights file \"" + path + "\"");
    std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (data.size() < 24 || memcmp(data.data(), "PNNWv001", 8) != 0) {
        throw std::runtime_error("\"" + path + "\" is not a PNNW flat binary (the CPU baseline backend reads only those)");
    }
    auto u32 = [&](size_t pos) {
        if (pos + 4 > data.size()) throw std::runtime_error("truncated weights file \"" + path + "\"");
        uint32_t v;
        memcpy(&v, data.data() + pos, 4);
        return v;
    };
    auto u64 = [&](size_t pos) {
        if (pos + 8 > data.size()) throw std::runtime_error("truncated weights file \"" + path + "\"");
        uint64_t v;
        memcpy(&v, data.data() + pos, 8);
        return v;
    };
    *width = (int)u32(8);
    *is_fc = u32(12) != 0;
    const uint32_t n = u32(16);
    size_t pos = 24;
    int64_t count_all = 0;
    double sum_all = 0.;
    std::map<std::string, std::vector<float>> by_name;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ln = u32(pos);
        pos += 4;
        if (pos + ln > data.size()) throw std::runtime_error("truncated weights file \"" + path + "\"");
        std::string name(data.data() + pos, ln);
        pos += ln;
        const uint32_t rank = u32(pos);
        pos += 4;
        std::vector<int64_t> dims;
        size_t count = 1;
        for (uint32_t r = 0; r < rank; ++r) {
            dims.push_back((int64_t)u32(pos));
            count *= (size_t)dims.back();
            pos += 4;
        }
        const uint64_t off = u64(pos), nbytes = u64(pos + 8);
        pos += 16;
        if (nbytes != count * 4 || off + nbytes > data.size()) throw std::runtime_error("bad tensor table entry \"" + name + "\"");
        if (out) {
            at::Tensor t = at::empty(dims, at::kFloat);
            memcpy(t.data_ptr<float>(), data.data() + off, nbytes);
            (*out)[name] = t;
        }
        if (checksum) {
            std::vector<float> v(count);
            memcpy(v.data(), data.data() + off, nbytes);
            by_name[name] = std::move(v);
        }
        count_all += (int64_t)count;
    }
    for (const auto& kv : by_name) {                 // same definition as libpnn_cuda's pnn_inspect_net_file
        double s = 0.;
        for (size_t i = 0; i < kv.second.size(); ++i) s += (double)(i % 7 + 1) * (double)kv.second[i];
        sum_all += s;
    }
    if (n_params) *n_params = count_all;
    if (checksum) *checksum = sum_all;
}

const at::Tensor& need(const RefNet& net, const std::string& name) {
    auto it = net.t.find(name);
    if (it == net.t.end()) throw std::runtime_error("weights file lacks tensor \"" + name + "\"");
    return it->second;
}

// TensorFlow 'SAME': (pad_before, pad_after) -- oracle/nets.py same_padding
void same_padding(int64_t n, int64_t k, int64_t s, int64_t* before, int64_t* after) {
    const int64_t out = (n + s - 1) / s;
    const int64_t total = std::max<int64_t>((out - 1) * s + k - n, 0);
    *before = total / 2;
    *after = total - total / 2;
}

at::Tensor lrelu(const at::Tensor& x) { return at::maximum(x * 0.1f, x); }      // pnn/tfutils.py:192

// x NCHW; w already [Cout, Cin, k, k]
at::Tensor conv_same(const at::Tensor& x, const at::Tensor& w, const at::Tensor& b, int64_t s) {
    const int64_t k = w.size(2);
    int64_t pt, pb, pl, pr;
    same_padding(x.size(2), k, s, &pt, &pb);
    same_padding(x.size(3), k, s, &pl, &pr);
    return at::conv2d(at::constant_pad_nd(x, {pl, pr, pt, pb}, 0), w, b, {s, s});
}

// x NCHW; w already [Cin, Cout, k, k]; conv2d_transpose 'SAME' = full transposed conv cropped by the forward pad_before
at::Tensor tconv_same(const at::Tensor& x, const at::Tensor& w, const at::Tensor& b, int64_t s) {
    const int64_t k = w.size(2);
    const int64_t ho = x.size(2) * s, wo = x.size(3) * s;
    int64_t pt, pb, pl, pr;
    same_padding(ho, k, s, &pt, &pb);
    same_padding(wo, k, s, &pl, &pr);
    at::Tensor y = at::conv_transpose2d(x, w, {}, {s, s});
    y = y.slice(2, pt, pt + ho).slice(3, pl, pl + wo);
    return y + b.view({1, -1, 1, 1});
}

void prepare(RefNet& net) {
    // reference layouts -> torch layouts, once
    for (auto& kv : net.t) {
        const std::string& name = kv.first;
        if (name.find("/weights") == std::string::npos) continue;
        if (name.find("convolution_") != std::string::npos) {
            // conv [k,k,Cin,Cout] -> [Cout,Cin,k,k]; tconv [k,k,Cout,Cin] -> [Cin,Cout,k,k]: the same permutation
            kv.second = kv.second.permute({3, 2, 0, 1}).contiguous();
        }
    }
}

at::Tensor forward(const RefNet& net, const float* above_or_flat, const float* left) {
    c10::InferenceMode guard;
    const int64_t W = net.width;
    if (net.is_fc) {
        at::Tensor x = at::from_blob((void*)above_or_flat, {1, 5 * W * W}, at::kFloat);
        for (int i = 0; i < 4; ++i) {
            const std::string s = std::to_string(i);
            x = at::addmm(need(net, "fully_connected/biases_" + s), x, need(net, "fully_connected/weights_" + s));
            if (i != 3) x = lrelu(x);
        }
        return x.reshape({W, W}).contiguous();
    }
    at::Tensor outs[2];
    for (int br = 0; br < 2; ++br) {
        const std::string bname = br == 0 ? "above" : "left";
        at::Tensor x = br == 0 ? at::from_blob((void*)above_or_flat, {1, 1, W, 3 * W}, at::kFloat)
                               : at::from_blob((void*)left, {1, 1, 2 * W, W}, at::kFloat);
        for (size_t i = 0; i < net.strides.size(); ++i) {
            const std::string p = "convolutional/branch_" + bname + "/convolution_" + std::to_string(i) + "/";
            x = lrelu(conv_same(x, need(net, p + "weights"), need(net, p + "biases"), net.strides[i]));
        }
        outs[br] = x;                                      // [1, C, 4, 12] / [1, C, 8, 4]
    }
    // merger (pnn/tfutils.py:60-73): per channel [48 above row-major | 32 left row-major] x [80, 16]
    const int64_t C = outs[0].size(1);
    at::Tensor cat = at::cat({outs[0].reshape({C, 1, 48}), outs[1].reshape({C, 1, 32})}, 2);     // [C, 1, 80]
    const std::string pm = "convolutional/merger/channelwise_fully_connected_merger/";
    at::Tensor m = at::bmm(cat, need(net, pm + "weights")) + need(net, pm + "biases").unsqueeze(1);   // [C, 1, 16]
    at::Tensor x = lrelu(m).reshape({1, C, 4, 4});
    const int nb = (int)net.strides.size();
    for (int i = 0; i < nb; ++i) {
        const std::string p = "convolutional/merger/transpose_convolution_" + std::to_string(i) + "/";
        x = tconv_same(x, need(net, p + "weights"), need(net, p + "biases"), net.strides[nb - 1 - i]);
        if (i != nb - 1) x = lrelu(x);
    }
    return x.reshape({W, W}).contiguous();
}

}  // namespace

struct pnn_handle {
    std::string error;
    std::map<std::pair<int, int>, std::unique_ptr<RefNet>> nets;
    bool null_backend = false;
    // intra-op threads per kind of net (PNN_REF_THREADS_FC / PNN_REF_THREADS_CONV, else PNN_REF_THREADS, else all cores):
    // the baseline is given the setting that is fastest on the box (tools/ref_backend_latency.py sweeps it)
    int threads_fc = 0, threads_conv = 0, threads_now = 0;
};

extern "C" {

const char* pnn_version(void) { return "libpnn_ref 0.2 (libtorch-CPU stand-in for the TensorFlow-CPU build; baseline only)"; }

int pnn_create(const char* paths_file, float, int qp_selection, int, pnn_handle** out) {
    if (!out) {
        g_error = "`out` is NULL";
        return -1;
    }
    *out = nullptr;
    if (qp_selection <= 0) {
        g_error = "The quantization parameter used for selecting each prediction neural network model is not strictly positive.";
        return -1;
    }
    if (paths_file && paths_file[0]) {
        g_error = "libpnn_ref takes its nets through pnn_load_net (the link seam's load_graph)";
        return -1;
    }
    pnn_handle* h = new pnn_handle();
    const char* e = getenv("PNN_REF_NULL");
    h->null_backend = e && atoi(e) != 0;
    auto env_int = [](const char* name) {
        const char* v = getenv(name);
        return v && atoi(v) > 0 ? atoi(v) : 0;
    };
    const int all = env_int("PNN_REF_THREADS");
    h->threads_fc = env_int("PNN_REF_THREADS_FC") ? env_int("PNN_REF_THREADS_FC") : all;
    h->threads_conv = env_int("PNN_REF_THREADS_CONV") ? env_int("PNN_REF_THREADS_CONV") : all;
    *out = h;
    return 0;
}

void pnn_destroy(pnn_handle* h) { delete h; }

int pnn_create_deferred(const char* paths_file, float mean, int qp_selection, int device, pnn_handle** out) {
    return pnn_create(paths_file, mean, qp_selection, device, out);
}

int pnn_release_at_exit(pnn_handle*) { return 0; }
int pnn_warm_up(pnn_handle*) { return 0; }

const char* pnn_last_error(pnn_handle* h) { return h ? h->error.c_str() : g_error.c_str(); }

int pnn_inspect_net_file(const char* path, int* width_target, int* is_fully_connected, int64_t* n_parameters, double* checksum) {
    try {
        if (!path) throw std::runtime_error("`path` is NULL");
        int w = 0, fc = 0;
        read_pnnw(path, &w, &fc, nullptr, n_parameters, checksum);
        if (width_target) *width_target = w;
        if (is_fully_connected) *is_fully_connected = fc;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
    return 0;
}

int pnn_load_net(pnn_handle* h, const char* path) {
    if (!h) return -1;
    try {
        if (!path) throw std::runtime_error("`path` is NULL");
        std::unique_ptr<RefNet> net(new RefNet());
        int fc = 0;
        read_pnnw(path, &net->width, &fc, &net->t, nullptr, nullptr);
        net->is_fc = fc != 0;
        net->strides = strides_branch(net->width);
        if (net->strides.empty()) throw std::runtime_error("unsupported target width");
        prepare(*net);
        h->nets[{net->width, fc}] = std::move(net);
    } catch (const std::exception& e) {
        h->error = e.what();
        return -1;
    }
    return 0;
}

int pnn_register_net(pnn_handle* h, const char* path) { return pnn_load_net(h, path); }   // the baseline loads at once
int pnn_set_hm_cache(pnn_handle*, int) { return 0; }                                       // and has no memo
int pnn_hm_cache_stats(pnn_handle*, int64_t* hits, int64_t* misses) {
    if (hits) *hits = 0;
    if (misses) *misses = 0;
    return 0;
}

// baseline-only helper (tools/ref_backend_latency.py): intra-op threads of the FC / convolutional nets
int pnn_ref_set_threads(pnn_handle* h, int threads_fc, int threads_conv) {
    if (!h) return -1;
    h->threads_fc = threads_fc;
    h->threads_conv = threads_conv;
    return 0;
}

int pnn_predict_hm_context(pnn_handle* h, int width, const float* above_or_flat, const float* left, float* out) {
    if (!h) return -1;
    try {
        if (!above_or_flat || !out) throw std::runtime_error("NULL buffer");
        auto it = h->nets.find({width, 1});
        if (it == h->nets.end()) it = h->nets.find({width, 0});
        if (it == h->nets.end()) throw std::runtime_error("no PNN of width " + std::to_string(width) + " is loaded");
        const RefNet& net = *it->second;
        if (h->null_backend) {
            memset(out, 0, (size_t)width * width * sizeof(float));
            return 0;
        }
        if (!net.is_fc && !left) left = above_or_flat + 3 * width * width;
        const int want = net.is_fc ? h->threads_fc : h->threads_conv;
        if (want > 0 && want != h->threads_now) {
            at::set_num_threads(want);
            h->threads_now = want;
        }
        const at::Tensor y = forward(net, above_or_flat, left);
        memcpy(out, y.data_ptr<float>(), (size_t)width * width * sizeof(float));
    } catch (const std::exception& e) {
        h->error = e.what();
        return -1;
    }
    return 0;
}

}  // extern "C"
