// Host side of libpnn_cuda: weight loading / re-packing, network plans, workspaces, the C ABI.
#include "../../include/pnn_cuda.h"
#include "pnn_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

using namespace pnn;

namespace {

thread_local std::string g_create_error;

// PNN_TIMING=1: wall-clock trace of the start-up and tear-down steps on stderr (tuning aid)
struct Trace {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const char* what;
    static bool on() {
        static const bool v = getenv("PNN_TIMING") && atoi(getenv("PNN_TIMING")) != 0;
        return v;
    }
    explicit Trace(const char* w) : what(w) {}
    ~Trace() {
        if (on()) fprintf(stderr, "[pnn timing] %-28s %8.1f ms\n", what,
                          1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
};

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e_));       \
        }                                                                                       \
    } while (0)

// reference pnn/PredictionNeuralNetwork.py:126-132
std::vector<int> strides_branch(int w) {
    switch (w) {
        case 4: return {1, 1};
        case 8: return {2, 1};
        case 16: return {2, 1, 2, 1};
        case 32: return {2, 2, 1, 2, 1};
        case 64: return {2, 2, 2, 2, 1};
    }
    return {};
}

// TensorFlow 'SAME': pad_before of a k x k window, stride s, input n
int same_pad_before(int n, int k, int s) {
    const int out = (n + s - 1) / s;
    const int total = std::max((out - 1) * s + k - n, 0);
    return total / 2;
}

uint16_t bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf16_to_float(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void reserve(size_t n) {
        if (n <= bytes) return;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        CUDA_TRY(cudaMalloc(&p, n));
        bytes = n;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// Uploads of a net under construction go to g_upload_stream (the handle's load stream), like the kernels that re-tile them:
// a plain cudaMemcpy from pageable memory may return before its DMA has landed, and is only ordered with the legacy stream.
// The source vector may die right after the call (the copy out of pageable memory is staged before the call returns).
thread_local cudaStream_t g_upload_stream = nullptr;

template <typename T>
T* upload(const std::vector<T>& v, std::vector<std::unique_ptr<DevBuf>>& keep) {
    keep.emplace_back(new DevBuf());
    keep.back()->reserve(std::max<size_t>(v.size() * sizeof(T), 16));
    CUDA_TRY(cudaMemcpyAsync(keep.back()->p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, g_upload_stream));
    return (T*)keep.back()->p;
}

enum StepKind { STEP_GEMM, STEP_MERGER, STEP_COL2IM, STEP_CONV_FIRST };

struct Step {
    StepKind kind;
    // buffers (indices into Net::buf_elems / workspace); -1 = unused
    int in0 = -1, in1 = -1, out = -1;
    bool is_final = false;
    // GEMM
    GemmGeom g{};
    const float* d_w32 = nullptr;
    const uint8_t* d_wt = nullptr;
    const float* d_bias = nullptr;
    // conv0
    int IH = 0, IW = 0, OH = 0, OW = 0, C = 0, k = 0, stride = 0, pad = 0, KP = 0;
    float bias_scalar = 0.f;
    std::shared_ptr<ConvFirstWeights> conv_first;   // host copy: travels as a kernel parameter
    // in-loop (batch-1) scheduling: independent steps (the two branches, the phases of a transposed convolution) carry
    // different lanes and run on different streams of the captured graph; a `join` step first waits for every lane
    int lane = 0;
    bool join = false;
};

struct Net {
    int W = 0;
    bool is_fc = false;
    std::vector<Step> steps;
    std::vector<int64_t> buf_elems;   // per-sample elements of every activation buffer
    std::vector<bool> buf_fp32_only;  // context inputs of conv nets are always fp32
    int in_above = -1, in_left = -1;  // conv: context buffers; FC: in_above = flat context
    std::vector<std::unique_ptr<DevBuf>> dev;   // weights
    int64_t param_count = 0;
    // workspace
    int64_t cap = 0;
    std::vector<std::unique_ptr<DevBuf>> ws0, ws1;
    DevBuf out_raw, out_u8, out_i32, out_psnr;
    int64_t cap_limit = (int64_t)1 << 40;     // ensure_workspace_fit: the largest capacity that fitted when memory was short
    // in-loop (batch-1) path: captured launch sequence; FC nets of width <= 8: per-CTA weight images (kernel_fc_inloop.cu)
    // and the activation vectors of the layer-per-launch fall-back
    cudaGraphExec_t hm_exec = nullptr;
    int hm_exec_precision = -1;
    int hm_launches = 0;
    DevBuf hm_vec[3];
    const float* fci_images = nullptr;
    int fci_stride = 0;
    struct TileJob {
        const float* w;
        int K, N;
        uint8_t* tiles;
    };
    std::vector<TileJob> tile_jobs;           // tcgen05 weight tiles still to be made (ensure_tiles)
    // memo of in-loop results keyed by the exact staged context (direct-mapped; pnn_set_hm_cache)
    struct HmCache {
        size_t entries = 0, key_words = 0, out_words = 0;
        std::vector<uint64_t> tags;
        std::vector<int32_t> keys;
        std::vector<uint32_t> outs;
    } cache;
    // asynchronous read-back of the image-block path (stream_out) still reading out_* of this net
    cudaEvent_t computed = nullptr, read_back = nullptr;
    bool read_back_pending = false;
    void drop_hm_graph() {
        if (hm_exec) cudaGraphExecDestroy(hm_exec);
        hm_exec = nullptr;
    }
    ~Net() {
        drop_hm_graph();
        if (computed) cudaEventDestroy(computed);
        if (read_back) cudaEventDestroy(read_back);
    }
};

// `header_only`: PNNW files are read up to the end of their tensor table (names, shapes, offsets checked against the file
// size) and the tensors stay empty; frozen graphs are always parsed whole.
FlatFile read_flat(const std::string& path, bool header_only = false) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open weights file \"" + path + "\"");
    std::vector<char> data;
    size_t file_size = 0;
    {
        f.seekg(0, std::ios::end);
        file_size = (size_t)f.tellg();
        f.seekg(0, std::ios::beg);
        char magic[8] = {0};
        f.read(magic, 8);
        f.seekg(0, std::ios::beg);
        const bool is_pnnw = file_size >= 24 && memcmp(magic, "PNNWv001", 8) == 0;
        // the tensor table of a PNNW file ends before its first (256-byte aligned) tensor: 64 KB hold any of them
        const size_t want = header_only && is_pnnw ? std::min<size_t>(file_size, (size_t)64 << 10) : file_size;
        data.resize(want);
        f.read(data.data(), (std::streamsize)want);
        if ((size_t)f.gcount() != want) throw std::runtime_error("cannot read weights file \"" + path + "\"");
    }
    if (data.size() < 24 || memcmp(data.data(), "PNNWv001", 8) != 0) {
        // not our flat binary: the frozen graph the reference's HM loads (load_graph, integration_...cpp:29-69)?
        FlatFile graph;
        try {
            read_frozen_graph(data, path, &graph);
        } catch (const std::exception& e) {
            throw std::runtime_error("\"" + path + "\" is neither a PNNW flat binary nor a frozen PNN graph (" + e.what() + ")");
        }
        return graph;
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
    FlatFile out;
    out.width = (int)u32(8);
    out.is_fc = u32(12) != 0;
    const uint32_t n = u32(16);
    size_t pos = 24;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ln = u32(pos);
        pos += 4;
        if (pos + ln > data.size()) throw std::runtime_error("truncated weights file \"" + path + "\"");
        std::string name(data.data() + pos, ln);
        pos += ln;
        const uint32_t rank = u32(pos);
        pos += 4;
        std::vector<int> dims;
        size_t count = 1;
        for (uint32_t r = 0; r < rank; ++r) {
            dims.push_back((int)u32(pos));
            count *= dims.back();
            pos += 4;
        }
        const uint64_t off = u64(pos), nbytes = u64(pos + 8);
        pos += 16;
        if (nbytes != count * 4 || off + nbytes > file_size) {
            throw std::runtime_error("bad tensor table entry \"" + name + "\" in \"" + path + "\"");
        }
        std::vector<float> v;
        if (!header_only) {
            v.resize(count);
            memcpy(v.data(), data.data() + off, nbytes);
        }
        out.t[name] = std::move(v);
        out.shape[name] = dims;
    }
    return out;
}

const std::vector<float>& need(const FlatFile& ff, const std::string& name, const std::vector<int>& shape) {
    auto it = ff.t.find(name);
    if (it == ff.t.end()) throw std::runtime_error("weights file lacks tensor \"" + name + "\"");
    if (ff.shape.at(name) != shape) throw std::runtime_error("tensor \"" + name + "\" has an unexpected shape");
    return it->second;
}

int add_buf(Net& net, int64_t elems, bool fp32_only = false) {
    net.buf_elems.push_back(elems);
    net.buf_fp32_only.push_back(fp32_only);
    return (int)net.buf_elems.size() - 1;
}

void add_gemm_weights(Net& net, Step& st, const std::vector<float>& w_kn, const std::vector<float>& bias) {
    st.d_w32 = upload(w_kn, net.dev);
    // [K][N] fp32 -> pre-swizzled bf16 hi/lo tiles: made on the device (34 M weights on one host thread cost seconds of
    // start-up) by the first call that runs the tensor-core kernels of this net (ensure_tiles), never while loading: the
    // persistent kernel of the in-loop path may hold every SM at that moment
    net.dev.emplace_back(new DevBuf());
    net.dev.back()->reserve(tc_total_bytes(st.g.N, st.g.K));
    st.d_wt = (const uint8_t*)net.dev.back()->p;
    net.tile_jobs.push_back({st.d_w32, st.g.K, st.g.N, (uint8_t*)net.dev.back()->p});
    st.d_bias = upload(bias, net.dev);
}

// Names and shapes of the variables of one PNN (reference pnn/components.py:103-180 FC; :10-101, 182-261 and
// pnn/tfutils.py:8-73, 75-139, 395-462 convolutional); what build_fc / build_conv will ask for.
void check_tensor_table(const FlatFile& ff) {
    const int W = ff.width;
    auto want = [&](const std::string& name, const std::vector<int>& shape) {
        auto it = ff.shape.find(name);
        if (it == ff.shape.end()) throw std::runtime_error("weights file lacks tensor \"" + name + "\"");
        if (it->second != shape) throw std::runtime_error("tensor \"" + name + "\" has an unexpected shape");
    };
    if (W != 4 && W != 8 && W != 16 && W != 32 && W != 64) throw std::runtime_error("unsupported target width " + std::to_string(W));
    if (ff.is_fc) {
        const int dims[5] = {5 * W * W, 1200, 1200, 1200, W * W};
        for (int i = 0; i < 4; ++i) {
            want("fully_connected/weights_" + std::to_string(i), {dims[i], dims[i + 1]});
            want("fully_connected/biases_" + std::to_string(i), {dims[i + 1]});
        }
        return;
    }
    const std::vector<int> strides = strides_branch(W);
    int c = 32;
    for (const char* bname : {"above", "left"}) {
        int c_in = 1;
        c = 32;
        for (size_t i = 0; i < strides.size(); ++i) {
            const int s = strides[i], k = 2 * s + 1;
            c *= s;
            const std::string p = std::string("convolutional/branch_") + bname + "/convolution_" + std::to_string(i) + "/";
            want(p + "weights", {k, k, c_in, c});
            want(p + "biases", {c});
            c_in = c;
        }
    }
    want("convolutional/merger/channelwise_fully_connected_merger/weights", {c, 80, 16});
    want("convolutional/merger/channelwise_fully_connected_merger/biases", {c, 16});
    const int nb = (int)strides.size();
    for (int i = 0; i < nb; ++i) {
        const int s = strides[nb - 1 - i], k = 2 * s + 1;
        const int c_out = i == nb - 1 ? 1 : c / s;
        const std::string p = "convolutional/merger/transpose_convolution_" + std::to_string(i) + "/";
        want(p + "weights", {k, k, c_out, c});
        want(p + "biases", {c_out});
        c = c_out;
    }
}

// reference pnn/components.py:103-180
void build_fc(Net& net, const FlatFile& ff) {
    const int W = net.W;
    const int dims[5] = {5 * W * W, 1200, 1200, 1200, W * W};
    int cur = add_buf(net, dims[0]);
    net.in_above = cur;
    for (int i = 0; i < 4; ++i) {
        Step st;
        st.kind = STEP_GEMM;
        GemmGeom& g = st.g;
        g.P = 1; g.OW = 1; g.Cin = dims[i]; g.TH = 1; g.TW = 1; g.IH = 1; g.IW = 1;
        g.sy_o = 0; g.sy_t = 0; g.cy = 0; g.sx_o = 0; g.sx_t = 0; g.cx = 0;
        g.in_sample_stride = dims[i];
        g.N = dims[i + 1]; g.K = dims[i];
        g.OHf = 1; g.OWf = 1; g.osy = 1; g.ooy = 0; g.osx = 1; g.oox = 0;
        g.out_sample_stride = dims[i + 1];
        g.leaky = i != 3;
        const std::string sfx = std::to_string(i);
        add_gemm_weights(net, st, need(ff, "fully_connected/weights_" + sfx, {dims[i], dims[i + 1]}),
                         need(ff, "fully_connected/biases_" + sfx, {dims[i + 1]}));
        net.param_count += (int64_t)dims[i] * dims[i + 1] + dims[i + 1];
        st.in0 = cur;
        if (i == 3) {
            st.is_final = true;
        } else {
            st.out = add_buf(net, dims[i + 1]);
            cur = st.out;
        }
        net.steps.push_back(st);
    }
    if (W <= 8) {
        // in-loop batch-1 kernels (kernel_fc_inloop.cu): per-CTA packed images of the fp32 weights
        auto wt = [&](int i) { return need(ff, "fully_connected/weights_" + std::to_string(i), {dims[i], dims[i + 1]}).data(); };
        auto bs = [&](int i) { return need(ff, "fully_connected/biases_" + std::to_string(i), {dims[i + 1]}).data(); };
        net.fci_images = upload(fci_build_images(wt(0), wt(1), wt(2), wt(3), bs(0), bs(1), bs(2), bs(3), dims[0], dims[4]), net.dev);
        net.fci_stride = fci_image_floats(dims[0]);
    }
}

// GEMM over rows that are pixels of a [H, W] map with K contiguous values each (no taps, no stride)
GemmGeom pixel_gemm_geom(int H, int W, int K, int N, int leaky) {
    GemmGeom g{};
    g.P = H * W; g.OW = W; g.Cin = K; g.TH = 1; g.TW = 1; g.IH = H; g.IW = W;
    g.sy_o = 1; g.sy_t = 0; g.cy = 0; g.sx_o = 1; g.sx_t = 0; g.cx = 0;
    g.in_sample_stride = (int64_t)H * W * K;
    g.N = N; g.K = K;
    g.OHf = H; g.OWf = W; g.osy = 1; g.ooy = 0; g.osx = 1; g.oox = 0;
    g.out_sample_stride = (int64_t)H * W * N;
    g.leaky = leaky;
    return g;
}

// reference pnn/components.py:10-101, 182-261
void build_conv(Net& net, const FlatFile& ff) {
    const int W = net.W;
    const std::vector<int> strides = strides_branch(W);
    if (strides.empty()) throw std::runtime_error("unsupported width for a convolutional PNN");
    int branch_out[2] = {-1, -1};
    int C = 32;
    for (int br = 0; br < 2; ++br) {
        const std::string bname = br == 0 ? "above" : "left";
        int h = br == 0 ? W : 2 * W, w = br == 0 ? 3 * W : W;
        int cur = add_buf(net, (int64_t)h * w, true);
        (br == 0 ? net.in_above : net.in_left) = cur;
        int c_in = 1, c = 32;
        for (size_t i = 0; i < strides.size(); ++i) {
            const int s = strides[i], k = 2 * s + 1;
            c *= s;
            const int oh = h / s, ow = w / s;
            const int pad_y = same_pad_before(h, k, s), pad_x = same_pad_before(w, k, s);
            const std::string p = "convolutional/branch_" + bname + "/convolution_" + std::to_string(i) + "/";
            const std::vector<float>& wt = need(ff, p + "weights", {k, k, c_in, c});
            const std::vector<float>& bs = need(ff, p + "biases", {c});
            net.param_count += (int64_t)wt.size() + bs.size();
            Step st;
            st.in0 = cur;
            st.lane = br;                                       // the two branches are independent
            if (i == 0) {
                if (pad_x != pad_y) throw std::runtime_error("unexpected asymmetric padding");
                // first convolution (one input channel): direct FFMA kernel, weights [k*k][C] as stored
                if ((c != 32 && c != 64) || oh % 2 || !((k == 5 && s == 2) || (k == 3 && s == 1))) {
                    throw std::runtime_error("unexpected first convolution");
                }
                st.kind = STEP_CONV_FIRST;
                st.out = add_buf(net, (int64_t)oh * ow * c);
                st.IH = h; st.IW = w; st.OH = oh; st.OW = ow; st.k = k; st.stride = s; st.pad = pad_y; st.C = c;
                st.conv_first.reset(new ConvFirstWeights());
                memset(st.conv_first.get(), 0, sizeof(ConvFirstWeights));
                for (int t = 0; t < k * k; ++t)
                    for (int co = 0; co < c; ++co) st.conv_first->w[t * 64 + co] = wt[(size_t)t * c + co];   // [k,k,1,C]
                for (int co = 0; co < c; ++co) st.conv_first->b[co] = bs[co];
            } else {
                st.out = add_buf(net, (int64_t)oh * ow * c);
                st.kind = STEP_GEMM;
                GemmGeom& g = st.g;
                g.P = oh * ow; g.OW = ow; g.Cin = c_in; g.TH = k; g.TW = k; g.IH = h; g.IW = w;
                g.sy_o = s; g.sy_t = 1; g.cy = -pad_y; g.sx_o = s; g.sx_t = 1; g.cx = -pad_x;
                g.in_sample_stride = (int64_t)h * w * c_in;
                g.N = c; g.K = k * k * c_in;
                g.OHf = oh; g.OWf = ow; g.osy = 1; g.ooy = 0; g.osx = 1; g.oox = 0;
                g.out_sample_stride = (int64_t)oh * ow * c;
                g.leaky = 1;
                add_gemm_weights(net, st, wt, bs);       // TF [k,k,Cin,Cout] is already [K][N]
            }
            net.steps.push_back(st);
            cur = st.out;
            h = oh; w = ow; c_in = c;
        }
        branch_out[br] = cur;
        C = c;
        if ((br == 0 && (h != 4 || w != 12)) || (br == 1 && (h != 8 || w != 4))) {
            throw std::runtime_error("unexpected branch output map");
        }
    }
    // merger (reference pnn/tfutils.py:8-73): weights [C,80,16] -> [80][16][C], biases [C,16] -> [16][C]
    {
        const std::string p = "convolutional/merger/channelwise_fully_connected_merger/";
        const std::vector<float>& wt = need(ff, p + "weights", {C, 80, 16});
        const std::vector<float>& bs = need(ff, p + "biases", {C, 16});
        net.param_count += (int64_t)wt.size() + bs.size();
        std::vector<float> wtr((size_t)80 * 16 * C), btr((size_t)16 * C);
        for (int c = 0; c < C; ++c) {
            for (int q = 0; q < 80; ++q)
                for (int o = 0; o < 16; ++o) wtr[((size_t)q * 16 + o) * C + c] = wt[((size_t)c * 80 + q) * 16 + o];
            for (int o = 0; o < 16; ++o) btr[(size_t)o * C + c] = bs[(size_t)c * 16 + o];
        }
        Step st;
        st.kind = STEP_MERGER;
        st.join = true;
        st.in0 = branch_out[0];
        st.in1 = branch_out[1];
        st.out = add_buf(net, (int64_t)16 * C);
        st.C = C;
        st.d_w32 = upload(wtr, net.dev);
        st.d_bias = upload(btr, net.dev);
        net.steps.push_back(st);
    }
    int cur = net.steps.back().out;
    int h = 4, w = 4, c = C;
    const int nb = (int)strides.size();
    for (int i = 0; i < nb; ++i) {
        const int s = strides[nb - 1 - i], k = 2 * s + 1;
        const bool last = i == nb - 1;
        const int c_out = last ? 1 : c / s;
        const int oh = h * s, ow = w * s;
        // conv2d_transpose 'SAME' is the gradient of the forward conv on the (oh, ow) map:
        // out[y] = sum over (iy, ky) with iy*s + ky - pad == y   (reference pnn/tfutils.py:455-459)
        const int pad = same_pad_before(oh, k, s);
        const std::string p = "convolutional/merger/transpose_convolution_" + std::to_string(i) + "/";
        const std::vector<float>& wt = need(ff, p + "weights", {k, k, c_out, c});
        const std::vector<float>& bs = need(ff, p + "biases", {c_out});
        net.param_count += (int64_t)wt.size() + bs.size();
        if (last) {
            // last transposed convolution (one output channel, linear): GEMM of every input pixel with the
            // k*k taps (D[(b,iy,ix), tap] = in[b,iy,ix,:] . w[tap,:]) + col2im with the fused output epilogue
            const int NP = (k * k + 15) / 16 * 16;
            Step gm;
            gm.kind = STEP_GEMM;
            gm.join = true;
            gm.in0 = cur;
            gm.out = add_buf(net, (int64_t)h * w * NP);
            gm.g = pixel_gemm_geom(h, w, c, NP, 0);
            std::vector<float> wkn((size_t)c * NP, 0.f), zero_bias((size_t)NP, 0.f);
            for (int t = 0; t < k * k; ++t)
                for (int ci = 0; ci < c; ++ci) wkn[(size_t)ci * NP + t] = wt[(size_t)t * c + ci];   // [k,k,1,Cin]
            add_gemm_weights(net, gm, wkn, zero_bias);
            net.steps.push_back(gm);
            Step st;
            st.kind = STEP_COL2IM;
            st.in0 = gm.out;
            st.is_final = true;
            st.IH = h; st.IW = w; st.C = c; st.k = k; st.stride = s; st.pad = pad; st.KP = NP;
            st.bias_scalar = bs[0];
            net.steps.push_back(st);
        } else {
            const int out_buf = add_buf(net, (int64_t)oh * ow * c_out);
            // one GEMM per output phase (py, px): output y = s*oy + py uses taps ky = ky0 + s*j with
            // ky0 = (py + pad) % s, reading input row iy = oy + (py + pad - ky0)/s - j.
            for (int py = 0; py < s; ++py) {
                for (int px = 0; px < s; ++px) {
                    const int ky0 = (py + pad) % s, kx0 = (px + pad) % s;
                    const int th = (k - ky0 + s - 1) / s, tw = (k - kx0 + s - 1) / s;
                    Step st;
                    st.kind = STEP_GEMM;
                    st.in0 = cur;
                    st.out = out_buf;
                    st.lane = py * s + px;                      // the phases write disjoint pixels of the same map
                    st.join = st.lane == 0;
                    GemmGeom& g = st.g;
                    g.P = h * w; g.OW = w; g.Cin = c; g.TH = th; g.TW = tw; g.IH = h; g.IW = w;
                    g.sy_o = 1; g.sy_t = -1; g.cy = (py + pad - ky0) / s;
                    g.sx_o = 1; g.sx_t = -1; g.cx = (px + pad - kx0) / s;
                    g.in_sample_stride = (int64_t)h * w * c;
                    g.N = c_out; g.K = th * tw * c;
                    g.OHf = oh; g.OWf = ow; g.osy = s; g.ooy = py; g.osx = s; g.oox = px;
                    g.out_sample_stride = (int64_t)oh * ow * c_out;
                    g.leaky = 1;
                    std::vector<float> wkn((size_t)g.K * c_out);
                    for (int jy = 0; jy < th; ++jy)
                        for (int jx = 0; jx < tw; ++jx)
                            for (int ci = 0; ci < c; ++ci)
                                for (int co = 0; co < c_out; ++co) {
                                    const int ky = ky0 + s * jy, kx = kx0 + s * jx;
                                    wkn[((size_t)(jy * tw + jx) * c + ci) * c_out + co] =
                                        wt[(((size_t)ky * k + kx) * c_out + co) * c + ci];
                                }
                    add_gemm_weights(net, st, wkn, bs);
                    net.steps.push_back(st);
                }
            }
            cur = out_buf;
        }
        h = oh; w = ow; c = c_out;
    }
    if (h != W || w != W) throw std::runtime_error("unexpected merger output map");
}

}  // namespace

struct pnn_handle {
    int device = 0;
    float mean = 0.f;
    int precision = PNN_PRECISION_BF16X3;
    std::string error;
    std::map<std::pair<int, int>, std::unique_ptr<Net>> nets;   // (width, is_fc)
    // nets registered but not uploaded yet (pnn_register_net, pnn_create with a paths file): a PNNW path whose header was
    // validated, or an already parsed frozen graph
    struct Pending {
        std::string path;
        std::shared_ptr<FlatFile> graph;
    };
    std::map<std::pair<int, int>, Pending> pending;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    size_t workspace_budget = (size_t)20 << 30;   // per net; pnn_set_workspace_budget / PNN_WORKSPACE_GB override (20 GB: one chunk for the bench's conv nets)
    // host-API staging
    DevBuf d_images, d_idx, d_rows, d_cols, d_in0, d_in1;
    // image-block host path: copies run on their own streams so that the upload of call k+1 and the read-back of call k
    // overlap the kernels of the other call (pnn_predict_image_blocks_async); two input sets alternate
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    struct InputSet {
        DevBuf images, idx, rows, cols;
        cudaEvent_t uploaded = nullptr, consumed = nullptr;
        bool used = false;
    } in_set[2];
    int in_next = 0;
    // HM path
    int32_t* hm_staged = nullptr;    // page-aligned inside hm_staged_storage, pinned (cudaHostRegister) once the device is up: header + 5*64*64 ints
    std::vector<int32_t> hm_staged_storage;
    bool device_ready = false;
    // warm-up thread (pnn_warm_up): device initialisation and the upload of the registered nets run beside the caller
    std::thread bg;
    std::mutex mu;                            // nets / pending / loading_key while the warm-up thread is alive
    std::condition_variable cv;
    std::atomic<bool> bg_active{false}, bg_init_done{false}, persist_stale{false};
    bool bg_started = false;
    std::string bg_error;
    std::pair<int, int> loading_key{0, -1};   // the net the warm-up thread is uploading right now
    cudaStream_t stream_load = nullptr;
    int32_t* hm_out = nullptr;       // pinned, 64*64 ints
    int32_t* d_hm_out_mapped = nullptr;      // device alias of hm_out (mapped pinned memory)
    float* hm_out_raw = nullptr;             // pinned + mapped, 64*64 floats (raw prediction)
    float* d_hm_out_raw_mapped = nullptr;
    int32_t* d_hm_staged_mapped = nullptr;   // device alias of hm_staged
    volatile uint64_t* hm_ll = nullptr;      // pinned + mapped {value, seq} pairs written by the persistent FC kernel
    uint2* d_hm_ll_mapped = nullptr;
    volatile uint64_t* fci_req = nullptr;    // pinned + mapped request of the persistent FC kernel (header + context pairs)
    uint2* d_fci_req_mapped = nullptr;
    DevBuf d_splitk;                         // split-K partial sums of the in-loop conv calls
    DevBuf d_fci_relay, d_fci_xchg, d_fc_stamps;
    unsigned long long* fc_stamps_host = nullptr;
    unsigned fc_seq = 0;                     // 30-bit sequence number of the last request (0 is never used)
    bool persist_running = false;            // fci_persist_kernel is resident on every SM: nothing else can run
    bool persist_failed = false;             // the cooperative launch was refused once: layer-per-launch fall-back from then on
    int fc_calls_since_stop = 0;
    bool hm_fused_fc = true;
    bool hm_split_k = true;
    bool hm_cache = false;
    int64_t hm_cache_hits = 0, hm_cache_misses = 0;
    cudaStream_t lane_stream[4] = {nullptr, nullptr, nullptr, nullptr};   // [0] unused (= the calling stream)
    cudaEvent_t lane_fork = nullptr, lane_done[4] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf d_hm_staged;
    int hm_width = 0;
    struct ContextDesc {                     // arguments of the last pnn_set_context
        int width = 0, pic_stride = 0, num_intra_neighbor = 0, unit_width = 0, unit_height = 0, above_units = 0, left_units = 0;
        const int32_t* roi_origin = nullptr;
        uint8_t flags[2 * 64 + 1];
    } hm_desc;
    bool hm_desc_pending = false;            // the pixels of hm_desc are not staged yet (pnn_set_context_lazy)
    struct HmInflight {                      // pnn_predict_hm_begin: ONE in-loop request posted to the GPU, not collected yet
        bool active = false;
        bool answered = false;               // the memo held the result at posting time: nothing is running
        bool persistent = false;             // posted to the persistent FC kernel (else: the net's graph was launched)
        Net* net = nullptr;
        uint64_t tag = 0;                    // memo slot of the staged context (hm_staged stays untouched until the collect)
        size_t slot = 0;
    } hm_inflight;
    bool hm_lazy_context = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float hm_ms = 0.f;
    // per-kernel profiling
    struct ProfRec {
        std::string name;
        int64_t M, N, K;
        double flops;
        bool is_gemm;
        cudaEvent_t e0, e1;
    };
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
    std::string prof_text;
};

static inline unsigned next_seq(unsigned seq) {
    seq = (seq + 1u) & 0x3fffffffu;
    return seq ? seq : 1u;
}

static inline void req_store(volatile uint64_t* p, uint32_t payload, unsigned seq) {
    __atomic_store_n((uint64_t*)p, (uint64_t)payload | ((uint64_t)seq << 32), __ATOMIC_RELEASE);
}

// The persistent FC kernel takes every SM (one CTA of ~200 KB shared memory each): a second one, or any other kernel of this
// process, would queue behind it for ever.  So at most ONE handle of the process owns a running persistent kernel, and any
// handle that is about to use the GPU for something else first asks the owner's kernel to leave.
static std::mutex g_persist_mu;
static pnn_handle* g_persist_owner = nullptr;

static void persist_stop_locked(pnn_handle* h) {
    if (!h->persist_running) return;
    h->fc_seq = next_seq(h->fc_seq);
    req_store(h->fci_req + 0, FCI_CMD_QUIT, h->fc_seq);
    h->persist_running = false;
    h->fc_calls_since_stop = 0;
    if (g_persist_owner == h) g_persist_owner = nullptr;
}

// Asks the persistent FC kernel (of this handle, and of any other handle of the process) to exit.  Everything enqueued
// afterwards (any stream) runs once its CTAs have left the SMs; the host does not wait.  Called at the top of every entry
// point that needs the GPU for something else.
static void persist_stop(pnn_handle* h) {
    std::lock_guard<std::mutex> lock(g_persist_mu);
    if (g_persist_owner && g_persist_owner != h) persist_stop_locked(g_persist_owner);
    persist_stop_locked(h);
}

static void drop_hm_caches(pnn_handle* h) {
    for (auto& kv : h->nets) kv.second->cache = Net::HmCache();
}

extern "C" {
static void hm_drain(pnn_handle* h);   // defined with the in-loop path below
}
struct Quiesce {   // entry points that use the GPU for anything but an in-loop FC call
    explicit Quiesce(pnn_handle* h) {
        if (!h) return;
        hm_drain(h);            // a request posted by pnn_predict_hm_begin is answered (and remembered) first
        persist_stop(h);
    }
};

namespace {

cudaEvent_t take_event(pnn_handle* h) {
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    return e;
}

// RAII bracket of one launch when profiling is on
struct ProfScope {
    pnn_handle* h;
    cudaStream_t s;
    size_t idx = 0;
    bool on;
    ProfScope(pnn_handle* h_, cudaStream_t s_, const char* name, int64_t M, int64_t N, int64_t K, bool is_gemm)
        : h(h_), s(s_), on(h_->profiling) {
        if (!on) return;
        pnn_handle::ProfRec r{name, M, N, K, 2.0 * (double)M * (double)N * (double)K, is_gemm, take_event(h), take_event(h)};
        cudaEventRecord(r.e0, s);
        h->prof.push_back(r);
        idx = h->prof.size() - 1;
    }
    ~ProfScope() {
        if (on) cudaEventRecord(h->prof[idx].e1, s);
    }
};

void load_flat_file(pnn_handle* h, const FlatFile& ff, bool from_warm_up = false);

Net* find_net(pnn_handle* h, int width, int is_fc) {
    const std::pair<int, int> key{width, is_fc ? 1 : 0};
    std::unique_lock<std::mutex> lock(h->mu, std::defer_lock);
    if (h->bg_active.load()) lock.lock();
    auto it = h->nets.find(key);
    if (it == h->nets.end()) {
        if (lock.owns_lock() && h->loading_key == key) {
            // the warm-up thread is uploading exactly this net
            h->cv.wait(lock, [&] { return h->loading_key != key; });
            it = h->nets.find(key);
        }
    }
    if (it == h->nets.end()) {
        auto pend = h->pending.find(key);
        if (pend != h->pending.end()) {
            // registered earlier, needed now
            const pnn_handle::Pending p = pend->second;
            h->pending.erase(pend);
            if (lock.owns_lock()) lock.unlock();
            Trace trace("load at first use");
            FlatFile ff;
            {
                Trace t("  read file");
                if (!p.graph) ff = read_flat(p.path);
            }
            load_flat_file(h, p.graph ? *p.graph : ff);
            if (h->bg_active.load()) lock.lock();
            it = h->nets.find(key);
        }
    }
    if (it == h->nets.end()) {
        throw std::runtime_error("no " + std::string(is_fc ? "fully-connected" : "convolutional") + " PNN of width " +
                                 std::to_string(width) + " is loaded");
    }
    return it->second.get();
}

int64_t choose_capacity(pnn_handle* h, const Net& net, int64_t n) {
    int64_t per = 0;
    for (int64_t e : net.buf_elems) per += e * 6;
    per += (int64_t)net.W * net.W * 9 + 8;
    const size_t budget = h->workspace_budget;
    int64_t cap = std::max<int64_t>(1, (int64_t)(budget / (size_t)per));
    // keep rows (cap * positions) comfortably inside int32
    cap = std::min<int64_t>(cap, (int64_t)1 << 19);
    cap = std::min<int64_t>(cap, net.cap_limit);
    return std::min(cap, std::max<int64_t>(n, 1));
}

void ensure_workspace(Net& net, int64_t cap) {
    if (cap <= net.cap) return;
    net.drop_hm_graph();                      // the captured launches hold the old buffer addresses
    net.cap = 0;                              // DevBuf::reserve frees before it allocates: if an allocation below throws,
                                              // the next call must not trust the buffers that were already replaced
    if (net.ws0.empty()) {
        for (size_t i = 0; i < net.buf_elems.size(); ++i) {
            net.ws0.emplace_back(new DevBuf());
            net.ws1.emplace_back(new DevBuf());
        }
    }
    for (size_t i = 0; i < net.buf_elems.size(); ++i) {
        net.ws0[i]->reserve((size_t)cap * net.buf_elems[i] * 4);
        if (!net.buf_fp32_only[i]) net.ws1[i]->reserve((size_t)cap * net.buf_elems[i] * 2);
    }
    const size_t px = (size_t)net.W * net.W;
    net.out_raw.reserve(cap * px * 4);
    net.out_u8.reserve(cap * px);
    net.out_i32.reserve(cap * px * 4);
    net.out_psnr.reserve(cap * 8);
    net.cap = cap;
}

Act act_of(Net& net, int buf) {
    Act a;
    a.p0 = net.ws0[buf]->p;
    a.p1 = net.ws1[buf]->p;
    return a;
}

// Makes the tcgen05 weight tiles of `net` on `stream`, ahead of the first launches that read them (same stream: ordered).
// The caller has asked the persistent kernel to leave.
// The activation workspace of a net (everything ensure_workspace made): given back when device memory runs short.
void release_workspace(Net& net) {
    net.drop_hm_graph();
    for (auto& b : net.ws0) b->release();
    for (auto& b : net.ws1) b->release();
    net.out_raw.release();
    net.out_u8.release();
    net.out_i32.release();
    net.out_psnr.release();
    net.cap = 0;
}

// ensure_workspace under memory pressure (another engine or framework on the same GPU; the budget is per net and five nets
// may be loaded): when an allocation fails, first the workspaces of the OTHER nets of the handle are given back (they grow
// again on their next call), then the capacity is halved until it fits -- a batch is cut into more chunks, and a block's
// prediction does not depend on the chunk it is computed in.  Returns the capacity that was allocated.
int64_t ensure_workspace_fit(pnn_handle* h, Net& net, int64_t cap) {
    bool others_released = false;
    for (;;) {
        try {
            ensure_workspace(net, cap);
            return cap;
        } catch (const std::exception&) {
            cudaGetLastError();                                     // (an allocation failure is not sticky)
            if (!others_released) {
                others_released = true;
                bool any = false;
                {
                    std::unique_lock<std::mutex> lock(h->mu, std::defer_lock);
                    if (h->bg_active.load()) lock.lock();
                    for (auto& kv : h->nets) {
                        if (kv.second.get() != &net && kv.second->cap > 0) {
                            release_workspace(*kv.second);
                            any = true;
                        }
                    }
                }
                if (any) continue;
            }
            if (cap <= 16) throw;
            release_workspace(net);                                 // (some of its buffers already have the larger size)
            cap = (cap + 1) / 2;
            net.cap_limit = cap;                                    // nested calls of this request must not ask for more again
        }
    }
}

void ensure_tiles(pnn_handle* h, Net& net, cudaStream_t stream) {
    if (net.tile_jobs.empty() || h->precision != PNN_PRECISION_BF16X3) return;
    for (const Net::TileJob& job : net.tile_jobs) h->launches += launch_make_tc_tiles(job.w, job.K, job.N, job.tiles, stream);
    net.tile_jobs.clear();
    CUDA_TRY(cudaGetLastError());
}

// Runs every layer of `net` on `n` samples whose contexts are already in the input buffers.
// `in_loop_fp32`: one sample of the codec's in-loop path (convolutional nets).  Every layer runs in fp32 on float buffers
// whatever the precision of the batched path: the GEMM-shaped layers through launch_gemm_skinny (16 x 16 output tiles over
// the whole K: no split-K, no reduce launch), the others through the fp32 variants of their kernels.
void run_net(pnn_handle* h, Net& net, int64_t n, const FinalOut& fin, cudaStream_t main_stream, bool allow_split_k = false,
             bool in_loop_fp32 = false) {
    const bool split = h->precision == PNN_PRECISION_BF16X3 && !in_loop_fp32;
    // In-loop calls (one sample): the kernels are far too small to fill the GPU, so independent steps run concurrently on
    // lane streams (fork / join with events; inside a stream capture they become parallel branches of the graph).
    static const bool lanes_enabled = !(getenv("PNN_HM_LANES") && atoi(getenv("PNN_HM_LANES")) == 0);
    const bool lanes = (allow_split_k || in_loop_fp32) && lanes_enabled && !h->profiling;
    bool lane_active[4] = {false, false, false, false};
    auto join_lanes = [&]() {
        for (int l = 1; l < 4; ++l) {
            if (!lane_active[l]) continue;
            CUDA_TRY(cudaEventRecord(h->lane_done[l], h->lane_stream[l]));
            CUDA_TRY(cudaStreamWaitEvent(main_stream, h->lane_done[l], 0));
            lane_active[l] = false;
        }
    };
    if (lanes) CUDA_TRY(cudaEventRecord(h->lane_fork, main_stream));          // the inputs of every lane are ready here
    for (const Step& st : net.steps) {
        cudaStream_t stream = main_stream;
        const int lane = lanes ? st.lane : 0;
        if (lanes) {
            if (st.join) {
                join_lanes();
                CUDA_TRY(cudaEventRecord(h->lane_fork, main_stream));         // everything before this step
            }
            if (lane > 0) {
                stream = h->lane_stream[lane];
                if (!lane_active[lane]) {
                    CUDA_TRY(cudaStreamWaitEvent(stream, h->lane_fork, 0));
                    lane_active[lane] = true;
                }
            }
        }
        switch (st.kind) {
            case STEP_GEMM: {
                GemmLaunch L{};
                L.g = st.g;
                L.in = act_of(net, st.in0);
                if (st.is_final) {
                    L.out_mode = OUT_FINAL;
                    L.fin = fin;
                } else {
                    L.out_mode = OUT_ACT;
                    L.out = act_of(net, st.out);
                }
                L.M = (int)(n * st.g.P);
                L.w_fp32 = st.d_w32;
                L.w_tiles = st.d_wt;
                L.bias = st.d_bias;
                L.split_k = 1;
                if (allow_split_k && split && !st.is_final) {
                    // in-loop batch-1 call: spread the K blocks of the few tiles over the SMs (fixed slicing)
                    const int tiles = tc_num_nt(st.g.N) * ((L.M + TC_BM - 1) / TC_BM);
                    const int num_kb = tc_num_kb(st.g.K);
                    if (tiles <= 32 && num_kb >= 4) {
                        int kb_per = (num_kb * tiles + 147) / 148;
                        if (kb_per < 2) kb_per = 2;
                        const int slices = (num_kb + kb_per - 1) / kb_per;
                        const size_t need_bytes = (size_t)slices * L.M * st.g.N * sizeof(float);
                        const size_t region = h->d_splitk.bytes / 4;            // one region per lane
                        if (slices > 1 && need_bytes <= region) {
                            L.split_k = slices;
                            L.partial = (float*)((char*)h->d_splitk.p + (size_t)lane * region);
                        }
                    }
                }
                ProfScope ps(h, stream, split ? "gemm_tc" : "gemm_fp32", L.M, st.g.N, st.g.K, true);
                if (in_loop_fp32) h->launches += launch_gemm_skinny(L, stream);
                else h->launches += split ? launch_gemm_tc(L, stream) : launch_gemm_fp32(L, stream);
                if (L.split_k > 1) h->launches += launch_splitk_reduce(L, stream);
                break;
            }
            case STEP_MERGER: {
                MergerLaunch L{};
                L.in0 = act_of(net, st.in0);
                L.in1 = act_of(net, st.in1);
                L.out = act_of(net, st.out);
                L.w = st.d_w32;
                L.bias = st.d_bias;
                L.n = (int)n; L.C = st.C; L.split = split;
                L.in_loop = (allow_split_k || in_loop_fp32) ? 1 : 0;
                ProfScope ps(h, stream, "merger", n * st.C, 16, 80, false);
                h->launches += launch_merger(L, stream);
                break;
            }
            case STEP_CONV_FIRST: {
                ConvFirstLaunch L{};
                L.in = (const float*)net.ws0[st.in0]->p;
                L.out = act_of(net, st.out);
                L.n = (int)n; L.IH = st.IH; L.IW = st.IW; L.OH = st.OH; L.OW = st.OW; L.C = st.C; L.k = st.k;
                L.stride = st.stride; L.pad = st.pad; L.split = split;
                L.in_loop = (allow_split_k || in_loop_fp32) ? 1 : 0;
                ProfScope ps(h, stream, "conv_first", n * st.OH * st.OW, st.C, st.k * st.k, false);
                h->launches += launch_conv_first(L, *st.conv_first, stream);
                break;
            }
            case STEP_COL2IM: {
                Col2imLaunch L{};
                L.d = act_of(net, st.in0);
                L.bias = st.bias_scalar;
                L.fin = fin;
                L.n = (int)n; L.IH = st.IH; L.IW = st.IW; L.k = st.k; L.stride = st.stride; L.pad = st.pad; L.NP = st.KP;
                L.split = split;
                ProfScope ps(h, stream, "col2im", n * st.IH * st.stride * st.IW * st.stride, 1, 9, false);
                h->launches += launch_col2im(L, stream);
                break;
            }
        }
    }
    if (lanes) join_lanes();
    CUDA_TRY(cudaGetLastError());
}

void check_masks(int W, int mask_w, int mask_h) {
    // reference sets/common.py:444-447
    if (mask_w < 0 || mask_w > W || mask_w % 4) throw std::runtime_error("`mask_w` does not belong to {0, 4, ..., width}");
    if (mask_h < 0 || mask_h > W || mask_h % 4) throw std::runtime_error("`mask_h` does not belong to {0, 4, ..., width}");
}

// device-pointer core of the image-block path
void image_blocks_device(pnn_handle* h, Net& net, const uint8_t* d_images, int n_images, int H, int Wimg, const int32_t* d_idx,
                         const int32_t* d_rows, const int32_t* d_cols, int64_t n, int mask_w, int mask_h, float* d_f32,
                         uint8_t* d_u8, double* d_psnr, cudaStream_t stream) {
    const int W = net.W;
    const int64_t px = (int64_t)W * W;
    const bool split = h->precision == PNN_PRECISION_BF16X3;
    const int64_t cap = ensure_workspace_fit(h, net, choose_capacity(h, net, n));
    ensure_tiles(h, net, stream);
    for (int64_t s0 = 0; s0 < n; s0 += cap) {
        const int64_t m = std::min(cap, n - s0);
        GatherLaunch G{};
        G.images = d_images;
        G.image_index = d_idx ? d_idx + s0 : nullptr;
        G.rows = d_rows + s0;
        G.cols = d_cols + s0;
        G.n = m; G.H = H; G.Wimg = Wimg; G.W = W; G.mask_w = mask_w; G.mask_h = mask_h; G.mean = h->mean; G.n_images = n_images;
        if (net.is_fc) {
            // reference sets/common.py:467-472: flattened above then flattened left in one row
            Act flat = act_of(net, net.in_above);
            G.above = flat;
            G.left = flat;
            const int64_t shift = 3 * px;
            if (split) {
                G.left.p0 = (__nv_bfloat16*)flat.p0 + shift;
                G.left.p1 = (__nv_bfloat16*)flat.p1 + shift;
            } else {
                G.left.p0 = (float*)flat.p0 + shift;
            }
            G.pitch_above = G.pitch_left = 5 * px;
            G.split = split;
        } else {
            G.above = act_of(net, net.in_above);
            G.left = act_of(net, net.in_left);
            G.pitch_above = 3 * px;
            G.pitch_left = 2 * px;
            G.split = 0;
        }
        {
            ProfScope ps(h, stream, "gather_image", m * 5 * px, 1, 1, false);
            h->launches += launch_gather_image(G, stream);
        }
        FinalOut fin{};
        fin.raw = d_f32 ? d_f32 + s0 * px : nullptr;
        uint8_t* u8 = d_u8 ? d_u8 + s0 * px : (d_psnr ? (uint8_t*)net.out_u8.p : nullptr);
        fin.u8 = u8;
        fin.mean = h->mean;
        fin.round_mode = PNN_ROUND_HALF_EVEN;
        run_net(h, net, m, fin, stream);
        if (d_psnr) {
            ProfScope ps(h, stream, "psnr", m * px, 1, 1, false);
            h->launches += launch_psnr(d_images, G.image_index, G.rows, G.cols, m, H, Wimg, W, u8, d_psnr + s0, stream, n_images);
        }
    }
    CUDA_TRY(cudaGetLastError());
}

void batch_device(pnn_handle* h, Net& net, const float* d_a, const float* d_l, int64_t n, float* d_out, cudaStream_t stream) {
    const int W = net.W;
    const int64_t px = (int64_t)W * W;
    const bool split = h->precision == PNN_PRECISION_BF16X3;
    const int64_t cap = ensure_workspace_fit(h, net, choose_capacity(h, net, n));
    ensure_tiles(h, net, stream);
    for (int64_t s0 = 0; s0 < n; s0 += cap) {
        const int64_t m = std::min(cap, n - s0);
        if (net.is_fc) {
            h->launches += launch_convert_input(d_a + s0 * 5 * px, act_of(net, net.in_above), m * 5 * px, split, stream);
        } else {
            CUDA_TRY(cudaMemcpyAsync(net.ws0[net.in_above]->p, d_a + s0 * 3 * px, (size_t)m * 3 * px * 4,
                                     cudaMemcpyDeviceToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(net.ws0[net.in_left]->p, d_l + s0 * 2 * px, (size_t)m * 2 * px * 4,
                                     cudaMemcpyDeviceToDevice, stream));
        }
        FinalOut fin{};
        fin.raw = d_out + s0 * px;
        fin.mean = h->mean;
        run_net(h, net, m, fin, stream);
    }
}

int fail(pnn_handle* h, const std::exception& e) {
    if (h) h->error = e.what();
    else g_create_error = e.what();
    return -1;
}

void load_flat_file(pnn_handle* h, const FlatFile& ff, bool from_warm_up) {
    Trace trace(ff.is_fc ? "  upload + tile (FC net)" : "  upload + tile (conv net)");
    // the persistent kernel holds pointers into the nets it serves: the caller's thread stops it here, the warm-up thread
    // only leaves a note (it must not touch the doorbell the caller may be ringing)
    if (!from_warm_up) persist_stop(h);
    std::unique_ptr<Net> net(new Net());
    net->W = ff.width;
    net->is_fc = ff.is_fc;
    g_upload_stream = h->stream_load;
    if (ff.width != 4 && ff.width != 8 && ff.width != 16 && ff.width != 32 && ff.width != 64) {
        throw std::runtime_error("unsupported target width " + std::to_string(ff.width));
    }
    if (ff.is_fc) build_fc(*net, ff);
    else build_conv(*net, ff);
    // (not cudaDeviceSynchronize: the persistent kernel of the in-loop path may be running and never finishes by itself)
    CUDA_TRY(cudaStreamSynchronize(h->stream_load));
    CUDA_TRY(cudaGetLastError());
    {
        std::unique_lock<std::mutex> lock(h->mu, std::defer_lock);
        if (h->bg_active.load()) lock.lock();
        h->pending.erase({ff.width, ff.is_fc ? 1 : 0});
        h->nets[{ff.width, ff.is_fc ? 1 : 0}] = std::move(net);
    }
    if (from_warm_up && ff.is_fc && ff.width <= 8) h->persist_stale.store(true);
}

void load_net_impl(pnn_handle* h, const std::string& path) { load_flat_file(h, read_flat(path)); }

// Validates the file now (a PNNW header with every tensor the net needs, or a whole frozen graph) and uploads at first use.
void register_net_impl(pnn_handle* h, const std::string& path) {
    std::shared_ptr<FlatFile> ff(new FlatFile(read_flat(path, /*header_only=*/true)));
    check_tensor_table(*ff);
    pnn_handle::Pending p;
    bool has_data = true;
    for (const auto& kv : ff->t) has_data = has_data && !kv.second.empty();
    if (has_data) p.graph = ff;               // a frozen graph: parsed already, keep it
    else p.path = path;
    persist_stop(h);
    h->nets.erase({ff->width, ff->is_fc ? 1 : 0});
    h->pending[{ff->width, ff->is_fc ? 1 : 0}] = p;
}

}  // namespace

extern "C" {

const char* pnn_version(void) { return "libpnn_cuda 0.2 (sm_100a)"; }

// Everything that touches the GPU at start-up: the device check (there is no CPU fallback), the context, streams, events and
// the pinned / mapped buffers of the in-loop path.  Run by pnn_create, or by the first call that needs the device after
// pnn_create_deferred.
static void init_device(pnn_handle* h) {
    Trace trace("device initialisation");
    int count = 0;
    cudaError_t e;
    {
        Trace t("  cudaGetDeviceCount");
        e = cudaGetDeviceCount(&count);
    }
    if (e != cudaSuccess || count == 0) {
        throw std::runtime_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libpnn_cuda has no CPU fallback)");
    }
    const int device = h->device;
    if (device < 0 || device >= count) throw std::runtime_error("invalid CUDA device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        throw std::runtime_error(std::string("device \"") + prop.name + "\" is not sm_100 (libpnn_cuda is built for sm_100a only)");
    }
    {
        Trace t("  context (cudaFree(0))");
        CUDA_TRY(cudaFree(0));
    }
    h->device_ready = true;                    // from here on pnn_destroy has something to release
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream_out, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream_load, cudaStreamNonBlocking));
    for (auto& set : h->in_set) {
        CUDA_TRY(cudaEventCreateWithFlags(&set.uploaded, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&set.consumed, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&h->lane_fork, cudaEventDisableTiming));
    for (int l = 1; l < 4; ++l) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->lane_done[l], cudaEventDisableTiming));
    }
    // (the staging buffer exists already: pnn_set_context may run before the device is needed)
    CUDA_TRY(cudaHostRegister(h->hm_staged, (HM_HEADER_INTS + 5 * 64 * 64) * sizeof(int32_t), cudaHostRegisterDefault));
    h->d_splitk.reserve((size_t)64 << 20);
    CUDA_TRY(cudaHostAlloc((void**)&h->hm_out, 64 * 64 * sizeof(int32_t), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&h->d_hm_out_mapped, h->hm_out, 0));
    CUDA_TRY(cudaHostAlloc((void**)&h->hm_out_raw, 64 * 64 * sizeof(float), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&h->d_hm_out_raw_mapped, h->hm_out_raw, 0));
    h->d_hm_staged.reserve((HM_HEADER_INTS + 5 * 64 * 64) * sizeof(int32_t));
    small_kernels_init();
    CUDA_TRY(gemm_tc_init());
}

static void ensure_device(pnn_handle* h) {
    if (h->bg_started) {
        // pnn_warm_up is initialising the device on its own thread
        while (!h->bg_init_done.load()) std::this_thread::sleep_for(std::chrono::microseconds(50));
        if (!h->bg_error.empty()) throw std::runtime_error(h->bg_error);
    } else if (!h->device_ready) {
        init_device(h);
    }
    CUDA_TRY(cudaSetDevice(h->device));
}

// Body of the warm-up thread: device initialisation, then the registered nets in the order the codec first needs them.
static void warm_up_thread(pnn_handle* h) {
    try {
        init_device(h);
    } catch (const std::exception& e) {
        h->bg_error = e.what();
    }
    h->bg_init_done.store(true);
    if (h->bg_error.empty()) {
        for (;;) {
            pnn_handle::Pending p;
            std::pair<int, int> key;
            {
                std::lock_guard<std::mutex> lock(h->mu);
                if (h->pending.empty()) break;
                key = h->pending.begin()->first;        // (4,1), (8,1), (16,0), (32,0), (64,0)
                p = h->pending.begin()->second;
                h->pending.erase(h->pending.begin());
                h->loading_key = key;
            }
            try {
                Trace trace("warm-up load");
                FlatFile ff;
                if (!p.graph) ff = read_flat(p.path);
                load_flat_file(h, p.graph ? *p.graph : ff, /*from_warm_up=*/true);
            } catch (const std::exception& e) {
                // leave it to the caller's thread, which reports the error of its own attempt
                std::lock_guard<std::mutex> lock(h->mu);
                h->pending[key] = p;
                h->loading_key = {0, -1};
                h->cv.notify_all();
                break;
            }
            {
                std::lock_guard<std::mutex> lock(h->mu);
                h->loading_key = {0, -1};
            }
            h->cv.notify_all();
        }
    }
    h->bg_active.store(false);
}

static void join_warm_up(pnn_handle* h) {
    if (h->bg.joinable()) h->bg.join();
}

static int create_impl(const char* paths_file, float mean_training, int qp_selection, int device, pnn_handle** out, bool deferred) {
    if (!out) {
        g_create_error = "`out` is NULL";
        return -1;
    }
    *out = nullptr;
    Trace trace("pnn_create");
    std::unique_ptr<pnn_handle> h(new pnn_handle());
    try {
        // reference TComPrediction.cpp(substitution):129-133
        if (qp_selection <= 0) {
            throw std::runtime_error("The quantization parameter used for selecting each prediction neural network model is not strictly positive.");
        }
        h->device = device;
        h->mean = mean_training;
        if (getenv("PNN_WORKSPACE_GB") && atoi(getenv("PNN_WORKSPACE_GB")) > 0) {
            h->workspace_budget = (size_t)atoi(getenv("PNN_WORKSPACE_GB")) << 30;
        }
        h->hm_staged_storage.resize((size_t)HM_HEADER_INTS + 5 * 64 * 64 + 1024);
        h->hm_staged = (int32_t*)(((uintptr_t)h->hm_staged_storage.data() + 4095) & ~(uintptr_t)4095);
        if (!deferred) init_device(h.get());
        if (paths_file && paths_file[0]) {
            // reference hevc/hm_common/c++/source_common/tools.cpp:40-110 (parse_file_strings_three_keys):
            // `width,is_pair,0,path`; "pair" models are used when listed and qp_selection >= 32
            // (TComPrediction.cpp(substitution):156)
            std::ifstream f(paths_file);
            if (!f) throw std::runtime_error(std::string("The file at \"") + paths_file + "\" cannot be opened.");
            // maps keyed by (first key, third key) like the reference's; the lookups below use (width, 0)
            std::map<std::pair<unsigned, unsigned>, std::string> single, pair;
            std::string line;
            while (std::getline(f, line)) {
                if (line.find_first_not_of(" \t\f\v\n\r") == std::string::npos) continue;   // tools.cpp:3-6, 74-78
                // split_string with the regular expression "[,]+" (tools.cpp:127-152): a run of delimiters separates two
                // substrings; a line that begins with one has an empty first substring
                std::vector<std::string> parts;
                size_t pos = 0;
                while (pos <= line.size()) {
                    const size_t next = line.find(',', pos);
                    if (next == std::string::npos) {
                        parts.push_back(line.substr(pos));
                        break;
                    }
                    parts.push_back(line.substr(pos, next - pos));
                    pos = line.find_first_not_of(',', next);
                    if (pos == std::string::npos) break;                      // trailing delimiters add nothing
                }
                if (parts.size() < 4) throw std::runtime_error("malformed line in the paths file: \"" + line + "\"");
                unsigned long width = 0, is_pair = 0, third = 0;
                try {                                                         // std::stoul, as tools.cpp:89-93
                    width = std::stoul(parts[0]);
                    is_pair = std::stoul(parts[1]);
                    third = std::stoul(parts[2]);
                } catch (const std::exception&) {
                    throw std::runtime_error("malformed line in the paths file: \"" + line + "\"");
                }
                std::string p = parts[3];
                p.erase(0, p.find_first_not_of(" \t\f\v\n\r"));
                p.erase(p.find_last_not_of(" \t\f\v\n\r") + 1);
                (is_pair ? pair : single)[std::make_pair((unsigned)width, (unsigned)third)] = p;
            }
            const bool use_pair = !pair.empty() && qp_selection >= 32;
            const std::map<std::pair<unsigned, unsigned>, std::string>& chosen = use_pair ? pair : single;
            for (unsigned width : {4u, 8u, 16u, 32u, 64u}) {
                auto it = chosen.find(std::make_pair(width, 0u));              // TComPrediction.cpp(substitution):158-170
                if (it == chosen.end()) throw std::runtime_error("the paths file has no entry for width " + std::to_string(width));
                register_net_impl(h.get(), it->second);
            }
        }
    } catch (const std::exception& e) {
        g_create_error = e.what();
        pnn_destroy(h.release());             // streams, events and pinned buffers created so far
        return -1;
    }
    *out = h.release();
    return 0;
}

int pnn_create(const char* paths_file, float mean_training, int qp_selection, int device, pnn_handle** out) {
    return create_impl(paths_file, mean_training, qp_selection, device, out, false);
}

int pnn_create_deferred(const char* paths_file, float mean_training, int qp_selection, int device, pnn_handle** out) {
    return create_impl(paths_file, mean_training, qp_selection, device, out, true);
}

int pnn_warm_up(pnn_handle* h) {
    if (!h) return -1;
    if (h->device_ready || h->bg_started) return 0;        // nothing left to overlap
    h->bg_started = true;
    h->bg_active.store(true);
    h->bg = std::thread(warm_up_thread, h);
    return 0;
}

int pnn_release_at_exit(pnn_handle* h) {
    if (!h) return -1;
    join_warm_up(h);
    if (!h->device_ready) return 0;
    Trace trace("pnn_release_at_exit");
    cudaSetDevice(h->device);
    hm_drain(h);
    persist_stop(h);
    cudaDeviceSynchronize();
    return 0;
}

int pnn_inspect_net_file(const char* path, int* width_target, int* is_fully_connected, int64_t* n_parameters,
                         double* checksum) {
    try {
        if (!path) throw std::runtime_error("`path` is NULL");
        const FlatFile ff = read_flat(path, /*header_only=*/checksum == nullptr);
        int64_t n = 0;
        double sum = 0.;
        for (const auto& kv : ff.t) {
            double s = 0.;
            for (size_t i = 0; i < kv.second.size(); ++i) s += (double)(i % 7 + 1) * (double)kv.second[i];
            sum += s;
            int64_t count = 1;
            for (int d : ff.shape.at(kv.first)) count *= d;
            n += count;
        }
        if (width_target) *width_target = ff.width;
        if (is_fully_connected) *is_fully_connected = ff.is_fc ? 1 : 0;
        if (n_parameters) *n_parameters = n;
        if (checksum) *checksum = sum;
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return -1;
    }
}

void pnn_destroy(pnn_handle* h) {
    if (!h) return;
    join_warm_up(h);
    if (!h->device_ready) {                   // created deferred and never used, or the device initialisation failed early
        delete h;
        return;
    }
    Trace trace("pnn_destroy");
    cudaSetDevice(h->device);
    hm_drain(h);
    persist_stop(h);
    cudaDeviceSynchronize();
    h->nets.clear();
    if (h->hm_staged) cudaHostUnregister(h->hm_staged);
    if (h->hm_out) cudaFreeHost(h->hm_out);
    if (h->hm_out_raw) cudaFreeHost(h->hm_out_raw);
    if (h->hm_ll) cudaFreeHost((void*)h->hm_ll);
    if (h->fci_req) cudaFreeHost((void*)h->fci_req);
    if (h->fc_stamps_host) cudaFreeHost(h->fc_stamps_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->lane_fork) cudaEventDestroy(h->lane_fork);
    for (int l = 1; l < 4; ++l) {
        if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
        if (h->lane_done[l]) cudaEventDestroy(h->lane_done[l]);
    }
    if (h->stream_in) cudaStreamDestroy(h->stream_in);
    if (h->stream_out) cudaStreamDestroy(h->stream_out);
    if (h->stream_load) cudaStreamDestroy(h->stream_load);
    for (auto& set : h->in_set) {
        if (set.uploaded) cudaEventDestroy(set.uploaded);
        if (set.consumed) cudaEventDestroy(set.consumed);
    }
    delete h;
}

const char* pnn_last_error(pnn_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int pnn_load_net(pnn_handle* h, const char* path) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (!path) throw std::runtime_error("`flat_binary_path` is NULL");
        ensure_device(h);
        load_net_impl(h, path);
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_register_net(pnn_handle* h, const char* path) {
    if (!h) return -1;
    try {
        if (!path) throw std::runtime_error("`path` is NULL");
        register_net_impl(h, path);
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_set_precision(pnn_handle* h, int precision) {
    if (!h) return -1;
    Quiesce quiesce(h);
    if (precision != PNN_PRECISION_FP32 && precision != PNN_PRECISION_BF16X3) {
        h->error = "unknown precision";
        return -1;
    }
    h->precision = precision;
    drop_hm_caches(h);
    return 0;
}

int pnn_debug_get_activation(pnn_handle* h, int width, int is_fc, int buffer_index, int64_t n_samples, float* out,
                             int64_t* elems_per_sample) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        Net& net = *find_net(h, width, is_fc);
        if (buffer_index < 0 || buffer_index >= (int)net.buf_elems.size()) throw std::runtime_error("no such activation buffer");
        const int64_t per = net.buf_elems[buffer_index];
        if (elems_per_sample) *elems_per_sample = per;
        if (!out) return 0;
        if (n_samples > net.cap) throw std::runtime_error("more samples requested than the workspace holds");
        ensure_device(h);
        CUDA_TRY(cudaDeviceSynchronize());
        const size_t count = (size_t)n_samples * per;
        const bool split = h->precision == PNN_PRECISION_BF16X3 && !net.buf_fp32_only[buffer_index];
        if (!split) {
            CUDA_TRY(cudaMemcpy(out, net.ws0[buffer_index]->p, count * 4, cudaMemcpyDeviceToHost));
        } else {
            std::vector<uint16_t> hi(count), lo(count);
            CUDA_TRY(cudaMemcpy(hi.data(), net.ws0[buffer_index]->p, count * 2, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(lo.data(), net.ws1[buffer_index]->p, count * 2, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < count; ++i) out[i] = bf16_to_float(hi[i]) + bf16_to_float(lo[i]);
        }
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_hevc_best_mode_device(pnn_handle* h, int width, const uint8_t* d_images, int n_images, int height, int width_image,
                              const int32_t* d_idx, const int32_t* d_rows, const int32_t* d_cols, int64_t n, int mask_w,
                              int mask_h, uint8_t* d_best, double* d_psnr, uint8_t* d_pred, void* stream) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0) throw std::runtime_error("negative number of blocks");
        if (width != 4 && width != 8 && width != 16 && width != 32 && width != 64) {
            throw std::runtime_error("the width of the target patch does not belong to {4, 8, 16, 32, 64}");
        }
        if (!d_images || !d_rows || !d_cols) throw std::runtime_error("NULL buffer");
        if (n_images > 1 && !d_idx) throw std::runtime_error("`image_index` is NULL while there are several images");
        check_masks(width, mask_w, mask_h);                       // intraprediction.py:66-72
        ensure_device(h);
        ProfScope ps(h, (cudaStream_t)stream, "hevc_best_mode", n, 35, (int64_t)width * width, false);
        h->launches += launch_hevc_best_mode(d_images, d_idx, d_rows, d_cols, n, height, width_image, width, mask_w, mask_h,
                                             d_best, d_psnr, d_pred, (cudaStream_t)stream, n_images);
        CUDA_TRY(cudaGetLastError());
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_hevc_best_mode(pnn_handle* h, int width, const uint8_t* images, int n_images, int height, int width_image,
                       const int32_t* idx, const int32_t* rows, const int32_t* cols, int64_t n, int mask_w, int mask_h,
                       uint8_t* best, double* psnr, uint8_t* pred) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0) throw std::runtime_error("negative number of blocks");
        if (!images || !rows || !cols) throw std::runtime_error("NULL buffer");
        if (n_images <= 0 || height <= 0 || width_image <= 0) throw std::runtime_error("empty image set");
        if (n_images > 1 && !idx) throw std::runtime_error("`image_index` is NULL while there are several images");
        for (int64_t i = 0; i < n; ++i) {
            // the block and the first pixel of its intra pattern must lie inside the image
            if (rows[i] < 1 || cols[i] < 1 || rows[i] + width > height || cols[i] + width > width_image) {
                throw std::runtime_error("block " + std::to_string(i) + " or its intra pattern anchor lies outside the image");
            }
            if (idx && (idx[i] < 0 || idx[i] >= n_images)) throw std::runtime_error("`image_index` out of range");
        }
        if (n == 0) return 0;
        ensure_device(h);
        cudaStream_t s = h->stream;
        const size_t img_bytes = (size_t)n_images * height * width_image, px = (size_t)width * width;
        h->d_images.reserve(img_bytes);
        h->d_rows.reserve(n * 4);
        h->d_cols.reserve(n * 4);
        if (idx) h->d_idx.reserve(n * 4);
        h->d_in0.reserve((size_t)n * (px + 1));
        h->d_in1.reserve((size_t)n * 8);
        CUDA_TRY(cudaMemcpyAsync(h->d_images.p, images, img_bytes, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(h->d_rows.p, rows, n * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(h->d_cols.p, cols, n * 4, cudaMemcpyHostToDevice, s));
        if (idx) CUDA_TRY(cudaMemcpyAsync(h->d_idx.p, idx, n * 4, cudaMemcpyHostToDevice, s));
        uint8_t* d_pred = (uint8_t*)h->d_in0.p;
        uint8_t* d_best = d_pred + (size_t)n * px;
        if (pnn_hevc_best_mode_device(h, width, (const uint8_t*)h->d_images.p, n_images, height, width_image,
                                      idx ? (const int32_t*)h->d_idx.p : nullptr, (const int32_t*)h->d_rows.p,
                                      (const int32_t*)h->d_cols.p, n, mask_w, mask_h, best ? d_best : nullptr,
                                      psnr ? (double*)h->d_in1.p : nullptr, pred ? d_pred : nullptr, s) != 0) {
            return -1;
        }
        if (best) CUDA_TRY(cudaMemcpyAsync(best, d_best, n, cudaMemcpyDeviceToHost, s));
        if (psnr) CUDA_TRY(cudaMemcpyAsync(psnr, h->d_in1.p, n * 8, cudaMemcpyDeviceToHost, s));
        if (pred) CUDA_TRY(cudaMemcpyAsync(pred, d_pred, (size_t)n * px, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_win_flags_device(pnn_handle* h, const double* d_psnr, const double* d_base, int64_t n, uint8_t* d_win, void* stream) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0 || !d_psnr || !d_base || !d_win) throw std::runtime_error("bad arguments");
        ensure_device(h);
        h->launches += launch_win_flags(d_psnr, d_base, n, d_win, (cudaStream_t)stream);
        CUDA_TRY(cudaGetLastError());
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_set_hm_fused(pnn_handle* h, int enabled) {
    if (!h) return -1;
    Quiesce quiesce(h);
    h->hm_fused_fc = enabled != 0;
    h->hm_split_k = enabled != 0;            // both in-loop optimisations follow the switch (0 = plain kernels)
    for (auto& kv : h->nets) kv.second->drop_hm_graph();
    drop_hm_caches(h);
    return 0;
}

int pnn_set_profiling(pnn_handle* h, int enabled) {
    if (!h) return -1;
    Quiesce quiesce(h);
    h->profiling = enabled != 0;
    return 0;
}

const char* pnn_profile_report(pnn_handle* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches, double* other_ms) {
    if (!h) return "";
    Quiesce quiesce(h);
    double g_ms = 0., g_fl = 0., o_ms = 0.;
    int64_t g_n = 0;
    try {
        ensure_device(h);
        CUDA_TRY(cudaDeviceSynchronize());
        struct Agg { int64_t launches = 0; double ms = 0., flops = 0.; };
        std::map<std::string, Agg> agg;
        for (auto& r : h->prof) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, r.e0, r.e1));
            h->event_pool.push_back(r.e0);
            h->event_pool.push_back(r.e1);
            char key[160];
            snprintf(key, sizeof(key), "%-12s %10lld %5lld %6lld", r.name.c_str(), (long long)r.M, (long long)r.N, (long long)r.K);
            Agg& a = agg[key];
            a.launches += 1;
            a.ms += ms;
            a.flops += r.flops;
            if (r.is_gemm) { g_ms += ms; g_fl += r.flops; g_n += 1; } else { o_ms += ms; }
        }
        h->prof.clear();
        std::ostringstream os;
        os << "kernel                M     N      K  launches        ms   TFLOP/s\n";
        for (auto& kv : agg) {
            char line[256];
            snprintf(line, sizeof(line), "%s %9lld %9.3f %9.2f\n", kv.first.c_str(), (long long)kv.second.launches, kv.second.ms,
                     kv.second.ms > 0 ? kv.second.flops / (kv.second.ms * 1e-3) / 1e12 : 0.);
            os << line;
        }
        h->prof_text = os.str();
    } catch (const std::exception& e) {
        h->error = e.what();
        h->prof_text = std::string("error: ") + e.what();
    }
    if (gemm_ms) *gemm_ms = g_ms;
    if (gemm_flops) *gemm_flops = g_fl;
    if (gemm_launches) *gemm_launches = g_n;
    if (other_ms) *other_ms = o_ms;
    return h->prof_text.c_str();
}

float pnn_debug_time_gemm(pnn_handle* h, int64_t M, int N, int K, int iters, int flags) {
    if (!h) return -1.f;
    Quiesce quiesce(h);
    try {
        if (M <= 0 || N <= 0 || K <= 0 || N % 16 || K % 8 || iters <= 0) throw std::runtime_error("bad problem size");
        ensure_device(h);
        DevBuf a_hi, a_lo, o_hi, o_lo, wt, bias;
        a_hi.reserve((size_t)M * K * 2);
        a_lo.reserve((size_t)M * K * 2);
        o_hi.reserve((size_t)M * N * 2);
        o_lo.reserve((size_t)M * N * 2);
        wt.reserve(tc_total_bytes(N, K));
        bias.reserve((size_t)N * 4);
        CUDA_TRY(cudaMemset(a_hi.p, 0, (size_t)M * K * 2));
        CUDA_TRY(cudaMemset(a_lo.p, 0, (size_t)M * K * 2));
        CUDA_TRY(cudaMemset(wt.p, 0, tc_total_bytes(N, K)));
        CUDA_TRY(cudaMemset(bias.p, 0, (size_t)N * 4));
        GemmLaunch L{};
        GemmGeom& g = L.g;
        g.P = 1; g.OW = 1; g.Cin = K; g.TH = 1; g.TW = 1; g.IH = 1; g.IW = 1;
        g.in_sample_stride = K; g.N = N; g.K = K; g.OHf = 1; g.OWf = 1; g.osy = 1; g.osx = 1;
        g.out_sample_stride = N; g.leaky = 1;
        L.in.p0 = a_hi.p; L.in.p1 = a_lo.p; L.out.p0 = o_hi.p; L.out.p1 = o_lo.p;
        L.out_mode = OUT_ACT; L.M = (int)M; L.w_tiles = (const uint8_t*)wt.p; L.bias = (const float*)bias.p;
        L.debug_flags = flags;
        cudaStream_t s = h->stream;
        launch_gemm_tc(L, s);
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(cudaEventRecord(h->ev0, s));
        for (int i = 0; i < iters; ++i) launch_gemm_tc(L, s);
        CUDA_TRY(cudaEventRecord(h->ev1, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(cudaGetLastError());
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        return ms / iters;
    } catch (const std::exception& e) {
        fail(h, e);
        return -1.f;
    }
}

int64_t pnn_launch_count(pnn_handle* h) { return h ? h->launches : 0; }

float pnn_last_hm_device_ms(pnn_handle* h) { return h ? h->hm_ms : 0.f; }

// Copies the context described by h->hm_desc into the staging buffer (reference extraction_context.cpp:56-205; masks and
// mean subtraction follow on the device).
static void stage_context(pnn_handle* h) {
    const pnn_handle::ContextDesc& d = h->hm_desc;
    const int W = d.width, cw = 3 * W;
    const int32_t* roi_origin = d.roi_origin;
    const int pic_stride = d.pic_stride, above_units = d.above_units, left_units = d.left_units;
    const int unit_width = d.unit_width, unit_height = d.unit_height;
    const uint8_t* flags = d.flags;
    int32_t* above = h->hm_staged + HM_HEADER_INTS;
    int32_t* left = above + 3 * W * W;
    const int total = above_units + left_units + 1;
    uint32_t lo = 0, hi = 0;
    int left_rows;
    if (d.num_intra_neighbor == total) {
        // extraction_context.cpp:56-90: straight copy
        const int32_t* p = roi_origin - (int64_t)W * pic_stride - W;
        for (int i = 0; i < W; ++i, p += pic_stride) memcpy(above + i * cw, p, cw * sizeof(int32_t));
        p = roi_origin - W;
        for (int i = 0; i < 2 * W; ++i, p += pic_stride) memcpy(left + i * W, p, W * sizeof(int32_t));
        for (int u = 0; u < above_units; ++u) (u < 32 ? lo : hi) |= 1u << (u & 31);
        left_rows = 2 * W;
        // a unit grid that does not cover the whole portion leaves the rest to the straight copy
        if (above_units * unit_width < 2 * W) {
            for (int u = above_units; u * unit_width < 2 * W; ++u) (u < 32 ? lo : hi) |= 1u << (u & 31);
        }
    } else {
        memset(above, 0, 5 * W * W * sizeof(int32_t));
        // extraction_context.cpp:119-127: the W x W block above-left is always copied
        const int32_t* p = roi_origin - (int64_t)W * pic_stride - W;
        for (int i = 0; i < W; ++i, p += pic_stride) memcpy(above + i * cw, p, W * sizeof(int32_t));
        // extraction_context.cpp:149-166
        for (int u = 0; u < above_units; ++u) {
            if (!flags[left_units + 1 + u]) continue;
            (u < 32 ? lo : hi) |= 1u << (u & 31);
            const int32_t* q = roi_origin - (int64_t)W * pic_stride + u * unit_width;
            int32_t* dst = above + W + u * unit_width;
            for (int j = 0; j < W; ++j, q += pic_stride, dst += cw) memcpy(dst, q, unit_width * sizeof(int32_t));
        }
        // extraction_context.cpp:189-205: source and destination advance only on available units
        int n_left = 0;
        for (int u = 0; u < left_units; ++u) n_left += flags[left_units - 1 - u] ? 1 : 0;
        left_rows = n_left * unit_height;
        p = roi_origin - W;
        for (int i = 0; i < left_rows; ++i, p += pic_stride) memcpy(left + i * W, p, W * sizeof(int32_t));
    }
    h->hm_staged[0] = (int32_t)lo;
    h->hm_staged[1] = (int32_t)hi;
    h->hm_staged[2] = unit_width;
    h->hm_staged[3] = left_rows;
    h->hm_desc_pending = false;
}

int pnn_set_context(pnn_handle* h, int width, const int32_t* roi_origin, int pic_stride, const uint8_t* flags,
                    int num_intra_neighbor, int unit_width, int unit_height, int above_units, int left_units) {
    if (!h) return -1;
    try {
        // same checks and messages as reference extraction_context.cpp:17-47
        if (!roi_origin) throw std::runtime_error("`piRoiOrigin` is NULL.");
        if (!flags) throw std::runtime_error("`bNeighborFlags` is NULL.");
        if (num_intra_neighbor <= 0) throw std::runtime_error("`iNumIntraNeighbor` is not strictly positive.");
        if (width != 4 && width != 8 && width != 16 && width != 32 && width != 64) {
            throw std::runtime_error("the width of the TB does not belong to {4, 8, 16, 32, 64}");
        }
        if (unit_width <= 0 || unit_height <= 0 || above_units <= 0 || left_units <= 0 || above_units > 64 || left_units > 64 ||
            above_units * unit_width > 2 * width || left_units * unit_height > 2 * width) {
            throw std::runtime_error("inconsistent neighbouring unit description");
        }
        // extraction_context.cpp:133-138
        if (num_intra_neighbor != above_units + left_units + 1 && !flags[left_units]) {
            throw std::runtime_error("The neighbouring unit above and on the left side of the current TB is not available.");
        }
        pnn_handle::ContextDesc& d = h->hm_desc;
        d.width = width;
        d.roi_origin = roi_origin;
        d.pic_stride = pic_stride;
        d.num_intra_neighbor = num_intra_neighbor;
        d.unit_width = unit_width;
        d.unit_height = unit_height;
        d.above_units = above_units;
        d.left_units = left_units;
        memcpy(d.flags, flags, (size_t)(above_units + left_units + 1));
        h->hm_desc_pending = true;
        h->hm_width = width;
        if (!h->hm_lazy_context) {   // lazy: pnn_predict_hm copies the pixels, if it is ever called
            hm_drain(h);             // the staged context is the memo key of a request in flight
            stage_context(h);
        }
    } catch (const std::exception& e) {
        h->hm_width = 0;
        return fail(h, e);
    }
    return 0;
}

int pnn_set_context_lazy(pnn_handle* h, int enabled) {
    if (!h) return -1;
    h->hm_lazy_context = enabled != 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// In-loop (batch-1) path.
// ---------------------------------------------------------------------------------------------
static bool persist_wanted(pnn_handle* h) { return h->hm_fused_fc && !h->persist_failed && !h->profiling; }

// Launches the persistent FC kernel (asynchronous) with the FC nets of widths 4 and 8 that are loaded.
static bool persist_start(pnn_handle* h) {
    if (h->persist_running) return true;
    if (!persist_wanted(h)) return false;
    FciPersist P{};
    std::unique_lock<std::mutex> lock(h->mu, std::defer_lock);
    if (h->bg_active.load()) lock.lock();
    for (int i = 0; i < 2; ++i) {
        auto it = h->nets.find({4 << i, 1});
        if (it == h->nets.end() || !it->second->fci_images) continue;
        FciNet& n = P.net[i];
        n.images = it->second->fci_images;
        n.W = 4 << i;
        n.K0 = 5 * n.W * n.W;
        n.N3 = n.W * n.W;
        n.present = 1;
        n.stride = it->second->fci_stride;
    }
    if (lock.owns_lock()) lock.unlock();
    if (!P.net[0].present && !P.net[1].present) return false;
    if (!h->fci_req) {
        CUDA_TRY(cudaHostAlloc((void**)&h->fci_req, FCI_REQ_PAIRS * sizeof(uint2), cudaHostAllocMapped));
        memset((void*)h->fci_req, 0, FCI_REQ_PAIRS * sizeof(uint2));
        CUDA_TRY(cudaHostGetDevicePointer((void**)&h->d_fci_req_mapped, (void*)h->fci_req, 0));
        CUDA_TRY(cudaHostAlloc((void**)&h->hm_ll, 128 * sizeof(uint2), cudaHostAllocMapped));
        memset((void*)h->hm_ll, 0, 128 * sizeof(uint2));
        CUDA_TRY(cudaHostGetDevicePointer((void**)&h->d_hm_ll_mapped, (void*)h->hm_ll, 0));
        h->d_fci_relay.reserve((size_t)FCI_COPIES * FCI_REQ_PAIRS * sizeof(uint2));
        h->d_fci_xchg.reserve((size_t)3 * FCI_COPIES * FCI_HID * sizeof(uint2));
        CUDA_TRY(cudaMemsetAsync(h->d_fci_relay.p, 0, (size_t)FCI_COPIES * FCI_REQ_PAIRS * sizeof(uint2), h->stream));
        CUDA_TRY(cudaMemsetAsync(h->d_fci_xchg.p, 0, (size_t)3 * FCI_COPIES * FCI_HID * sizeof(uint2), h->stream));
    }
    P.req = h->d_fci_req_mapped;
    P.relay = (uint2*)h->d_fci_relay.p;
    P.xchg = (uint2*)h->d_fci_xchg.p;
    P.out_ll = h->d_hm_ll_mapped;
    P.mean = h->mean;
    P.round_mode = PNN_ROUND_HALF_AWAY;
    P.seq0 = next_seq(h->fc_seq);
    static const bool want_stamps = getenv("PNN_FC_STAMPS") && atoi(getenv("PNN_FC_STAMPS")) != 0;
    if (want_stamps) {
        if (!h->d_fc_stamps.p) {
            h->d_fc_stamps.reserve(16 * sizeof(unsigned long long));
            CUDA_TRY(cudaHostAlloc((void**)&h->fc_stamps_host, 16 * sizeof(unsigned long long), cudaHostAllocDefault));
        }
        P.stamps = (unsigned long long*)h->d_fc_stamps.p;
    }
    cudaError_t e;
    {
        std::lock_guard<std::mutex> owner_lock(g_persist_mu);
        // a process that exits without destroying its handle must not leave a spinning kernel to the driver's teardown
        static bool at_exit_registered = false;
        if (!at_exit_registered) {
            at_exit_registered = true;
            atexit([] {
                std::lock_guard<std::mutex> lock(g_persist_mu);
                if (g_persist_owner) persist_stop_locked(g_persist_owner);
            });
        }
        if (g_persist_owner && g_persist_owner != h) persist_stop_locked(g_persist_owner);   // its kernel leaves, ours queues behind
        e = launch_fci_persist(P, h->stream);
        if (e == cudaSuccess) g_persist_owner = h;
    }
    if (e != cudaSuccess) {
        // 148 CTAs of ~200 KB cannot be co-resident on this device / partition: serve the calls layer by layer instead
        cudaGetLastError();
        h->persist_failed = true;
        return false;
    }
    h->launches += 1;
    h->persist_running = true;
    h->fc_calls_since_stop = 0;
    return true;
}

// One in-loop FC prediction through the persistent kernel: the staged context goes out as {payload, seq} pairs, the
// outputs come back the same way (8-byte accesses are single-copy atomic on both sides).
// Posts the staged context to the persistent FC kernel; fc_collect polls the answer.
static void fc_post(pnn_handle* h, Net& net) {
    const int W = net.W, K0 = 5 * W * W;
    const int32_t* staged = h->hm_staged;
    const unsigned seq = h->fc_seq = next_seq(h->fc_seq);
    volatile uint64_t* req = h->fci_req;
    // The context travels as 10-bit codes, three per pair (0..255 = reconstruction pixel, the kernel subtracts the mean;
    // 0x100 = masked / unavailable): a 4x4 context is 27 pairs, which the gate CTA sees together with the header in a single
    // PCIe read.  A pre-processed float context (pnn_predict_hm_context) is turned back into codes when every value is
    // exactly `pixel - mean` or 0, which is what extract_context_portions produces; anything else travels as float bits.
    uint32_t codes[FCI_CTX_MAX];
    bool as_codes = true;
    if (staged[2] == 0) {
        for (int e = 0; e < K0 && as_codes; ++e) {
            float v;
            memcpy(&v, staged + HM_HEADER_INTS + e, sizeof(float));
            if (v == 0.f) {
                codes[e] = 0x100u;
            } else {
                const long p = lrintf(v + h->mean);
                as_codes = p >= 0 && p <= 255 && (float)p - h->mean == v;
                codes[e] = (uint32_t)p;
            }
        }
    } else {
        const int na = 3 * W * W;
        for (int e = 0; e < K0 && as_codes; ++e) {
            const int32_t raw = staged[HM_HEADER_INTS + e];
            bool masked = false;
            if (e < na) {
                const int cc = e % (3 * W);
                if (cc >= W) {
                    const int u = (cc - W) / staged[2];
                    masked = !(u < 32 ? ((uint32_t)staged[0] >> u) & 1u : ((uint32_t)staged[1] >> (u - 32)) & 1u);
                }
            } else {
                masked = (e - na) / W >= staged[3];
            }
            as_codes = masked || (raw >= 0 && raw <= 255);
            codes[e] = masked ? 0x100u : (uint32_t)raw;
        }
    }
    uint32_t cmd = (W == 8 ? FCI_CMD_NET : 0u);
    if (as_codes) {
        const int n_pay = (K0 + 2) / 3;
        for (int i = 0; i < n_pay; ++i) {
            uint32_t pay = codes[3 * i];
            if (3 * i + 1 < K0) pay |= codes[3 * i + 1] << 10;
            if (3 * i + 2 < K0) pay |= codes[3 * i + 2] << 20;
            req_store(req + 1 + i, pay, seq);
        }
    } else {
        // a bit depth above 8 or a hand-made context: masks and mean on the host (same fp32 operations as hm_value_from_raw)
        cmd |= FCI_CMD_FLOAT;
        const int na = 3 * W * W;
        for (int e = 0; e < K0; ++e) {
            float v;
            if (staged[2] == 0) {
                memcpy(&v, staged + HM_HEADER_INTS + e, sizeof(float));
            } else {
                v = (float)staged[HM_HEADER_INTS + e] - h->mean;
                if (e < na) {
                    const int cc = e % (3 * W);
                    if (cc >= W) {
                        const int u = (cc - W) / staged[2];
                        if (!(u < 32 ? ((uint32_t)staged[0] >> u) & 1u : ((uint32_t)staged[1] >> (u - 32)) & 1u)) v = 0.f;
                    }
                } else if ((e - na) / W >= staged[3]) {
                    v = 0.f;
                }
            }
            uint32_t bits;
            memcpy(&bits, &v, sizeof(bits));
            req_store(req + 1 + e, bits, seq);
        }
    }
    req_store(req + 0, cmd, seq);
    h->fc_calls_since_stop += 1;
}

static void fc_collect(pnn_handle* h, Net& net) {
    const int W = net.W, n_out = W * W;
    const unsigned seq = h->fc_seq;
    long long spins = 0;
    bool warned = false;
    const std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
    for (int n = 0; n < n_out; ++n) {
        for (;;) {
            const uint64_t a = h->hm_ll[n], b = h->hm_ll[64 + n];
            if ((uint32_t)(a >> 32) == seq && (uint32_t)(b >> 32) == seq) {
                const uint32_t raw_bits = (uint32_t)a;
                memcpy(h->hm_out_raw + n, &raw_bits, sizeof(float));
                h->hm_out[n] = (int32_t)(uint32_t)b;
                break;
            }
            if ((++spins & 0xfffff) == 0 && Trace::on() && !warned &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 1.) {
                warned = true;
                fprintf(stderr, "[pnn timing] waiting > 1 s for the persistent kernel: W=%d seq=%u output %d of %d, stream: %s\n", W, seq, n, n_out,
                        cudaGetErrorString(cudaStreamQuery(h->stream)));
            }
            if ((spins & 0xfffff) == 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 10.) {
                // the kernel never got the SMs (another process or handle holds them) or died: leave it a way out, fall back
                // to plain launches from now on, and report -- without blocking on the stream
                persist_stop(h);
                h->persist_failed = true;
                throw std::runtime_error(std::string("the persistent FC kernel did not answer within 10 s (") +
                                         cudaGetErrorString(cudaStreamQuery(h->stream)) + ")");
            }
        }
    }
    h->hm_ms = 0.f;
    if (h->fc_stamps_host && (seq % 2000u) == 0) {
        // tuning aid: where a call spends its time on the device (copy engine: the SMs are taken by the persistent kernel)
        CUDA_TRY(cudaMemcpyAsync(h->fc_stamps_host, h->d_fc_stamps.p, 11 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream_out));
        CUDA_TRY(cudaStreamSynchronize(h->stream_out));
        const unsigned long long* t = h->fc_stamps_host;
        const double ns_per_cycle = (double)(t[10] - t[9]) / (double)(t[8] - t[0]);
        fprintf(stderr, "fci W=%d ns since the gate saw the request: relay seen by CTA 1 %.0f | layer publish / collect:", W,
                ns_per_cycle * (double)(long long)(t[1] - t[0]));   // CTA 1 sits on another SM: its counter is only roughly aligned
        for (int i = 2; i < 9; ++i) fprintf(stderr, " %.0f", ns_per_cycle * (double)(t[i] - t[0]));
        fprintf(stderr, " | SM clock %.0f MHz\n", 1e3 / ns_per_cycle);
    }
}

// Enqueues the launches of one in-loop prediction on `s` (called once per net under stream capture).
static void enqueue_hm(pnn_handle* h, Net& net, cudaStream_t s) {
    const int W = net.W;
    const bool split = h->precision == PNN_PRECISION_BF16X3;
    const size_t staged_bytes = (size_t)(HM_HEADER_INTS + 5 * W * W) * sizeof(int32_t);
    CUDA_TRY(cudaMemcpyAsync(h->d_hm_staged.p, h->hm_staged, staged_bytes, cudaMemcpyHostToDevice, s));
    FinalOut fin{};
    // the last kernel writes the (<= 16 KB) prediction straight into mapped pinned host memory: no copy node
    fin.i32 = h->d_hm_out_mapped;
    fin.raw = h->d_hm_out_raw_mapped;
    fin.mean = h->mean;
    fin.round_mode = PNN_ROUND_HALF_AWAY;
    int launches = 0;
    if (net.is_fc && net.fci_images) {
        // FC nets of width 4 / 8, one launch per layer: the same device functions and per-CTA weight images as the
        // persistent kernel (identical bits), HM gather fused into the first layer
        for (int layer = 0; layer < 4; ++layer) {
            FciLayerLaunch L{};
            L.images = net.fci_images;
            L.K0 = 5 * W * W; L.N3 = W * W; L.W = W; L.stride = net.fci_stride; L.layer = layer;
            L.staged = (const int32_t*)h->d_hm_staged.p;
            L.x = layer > 0 ? (const float*)net.hm_vec[layer - 1].p : nullptr;
            L.y = layer < 3 ? (float*)net.hm_vec[layer].p : nullptr;
            L.mean = h->mean;
            L.fin = fin;
            launches += launch_fci_layer(L, s);
        }
    } else {
        // convolutional nets, and FC nets wider than 8 (the offline comparison nets): the batched kernels on one sample
        GatherHmLaunch G{};
        G.staged = (const int32_t*)h->d_hm_staged.p;
        G.W = W;
        G.mean = h->mean;
        if (net.is_fc) {
            Act flat = act_of(net, net.in_above);
            G.above = flat;
            G.left = flat;
            const int64_t shift = 3 * (int64_t)W * W;
            if (split) {
                G.left.p0 = (__nv_bfloat16*)flat.p0 + shift;
                G.left.p1 = (__nv_bfloat16*)flat.p1 + shift;
            } else {
                G.left.p0 = (float*)flat.p0 + shift;
            }
            G.split = split;
        } else {
            G.above = act_of(net, net.in_above);
            G.left = act_of(net, net.in_left);
            G.split = 0;
        }
        launches += launch_gather_hm(G, s);
        const int64_t before = h->launches;
        // convolutional nets.  bf16x3 (default): the tensor-core kernels on one sample, their K blocks spread over the SMs
        // (split-K, fixed slicing) -- measured faster than the fp32 batch-1 layers (CONV-16 / 32 / 64: 104 / 147 / 203 us
        // against 110 / 203 / 458 us per call); fp32 precision: launch_gemm_skinny, fp32 from end to end.
        // pnn_set_hm_fused(0): the plain batched kernels.
        const bool fp32_path = h->precision == PNN_PRECISION_FP32 && h->hm_split_k && !net.is_fc;
        run_net(h, net, 1, fin, s, /*allow_split_k=*/h->hm_split_k && !net.is_fc && !fp32_path, /*in_loop_fp32=*/fp32_path);
        launches += (int)(h->launches - before);
        h->launches = before;
    }
    net.hm_launches = launches;
}

// Replays the captured launch sequence of `net` on the staged context (captures it on first use); graph_collect waits
// for it.
static void graph_post(pnn_handle* h, Net& net) {
    const bool restart = h->persist_running && h->fc_calls_since_stop > 0;   // FC calls are interleaved with this net's
    persist_stop(h);
    ensure_workspace(net, 1);
    cudaStream_t s = h->stream;
    if (!(net.is_fc && net.fci_images)) ensure_tiles(h, net, s);   // (not under the capture below: made once, not per replay)
    if (!net.hm_exec || net.hm_exec_precision != h->precision) {
        // capture the launch sequence once; every later call replays it (the per-call availability
        // masks travel in the staged header, so no kernel parameter changes between calls)
        net.drop_hm_graph();
        if (net.is_fc) {
            for (int i = 0; i < 3; ++i) net.hm_vec[i].reserve(1280 * sizeof(float));   // no allocation under capture
        }
        const bool prof = h->profiling;
        h->profiling = false;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            enqueue_hm(h, net, s);
        } catch (...) {
            cudaStreamEndCapture(s, &graph);
            if (graph) cudaGraphDestroy(graph);
            h->profiling = prof;
            throw;
        }
        h->profiling = prof;
        CUDA_TRY(cudaStreamEndCapture(s, &graph));
        cudaError_t e = cudaGraphInstantiate(&net.hm_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            net.hm_exec = nullptr;
            throw std::runtime_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
        net.hm_exec_precision = h->precision;
    }
    CUDA_TRY(cudaEventRecord(h->ev0, s));
    CUDA_TRY(cudaGraphLaunch(net.hm_exec, s));
    CUDA_TRY(cudaEventRecord(h->ev1, s));
    // the codec alternates this net with hundreds of FC calls: bring the persistent kernel back behind the graph, while
    // the host goes on with the result
    if (restart) persist_start(h);
}

static void graph_collect(pnn_handle* h, Net& net) {
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    h->launches += net.hm_launches;
    CUDA_TRY(cudaEventElapsedTime(&h->hm_ms, h->ev0, h->ev1));
}

// ---- memo of in-loop results: identical staged context -> identical prediction (the nets are pure functions), so a
// repeated context -- the codec evaluates the same TU in the fast pass, in the RD pass and again for the final
// reconstruction -- is answered from host memory.  Direct-mapped on a 64-bit hash, verified by comparing the whole key.
static uint64_t hash_words(const int32_t* p, size_t n) {
    uint64_t hsh = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    size_t i = 0;
    for (; i + 1 < n; i += 2) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        hsh = (hsh ^ w) * 0xD6E8FEB86659FD93ull;
        hsh ^= hsh >> 32;
    }
    if (i < n) {
        hsh = (hsh ^ (uint32_t)p[i]) * 0xD6E8FEB86659FD93ull;
        hsh ^= hsh >> 32;
    }
    return hsh ? hsh : 1ull;
}

// Posts the staged context: looks it up in the memo, else hands it to the persistent FC kernel / launches the net's graph.
// Does not wait: hm_collect does, and remembers the answer.  hm_staged must stay untouched in between.
static void hm_post(pnn_handle* h, Net& net) {
    if (h->persist_stale.exchange(false)) persist_stop(h);    // the warm-up thread added an FC net: relaunch with it
    const int W = net.W;
    const size_t key_words = (size_t)HM_HEADER_INTS + 5 * W * W, out_words = (size_t)2 * W * W;
    Net::HmCache& c = net.cache;
    pnn_handle::HmInflight& f = h->hm_inflight;
    f.net = &net;
    f.answered = false;
    f.tag = 0;
    f.slot = 0;
    if (h->hm_cache) {
        if (c.entries == 0) {
            // about 16 MB per net
            c.entries = std::max<size_t>(64, ((size_t)16 << 20) / ((key_words + out_words) * 4));
            c.key_words = key_words;
            c.out_words = out_words;
            c.tags.assign(c.entries, 0);
            c.keys.resize(c.entries * key_words);
            c.outs.resize(c.entries * out_words);
        }
        f.tag = hash_words(h->hm_staged, key_words);
        f.slot = (size_t)(f.tag % c.entries);
        if (c.tags[f.slot] == f.tag && memcmp(c.keys.data() + f.slot * key_words, h->hm_staged, key_words * 4) == 0) {
            memcpy(h->hm_out_raw, c.outs.data() + f.slot * out_words, (size_t)W * W * 4);
            memcpy(h->hm_out, c.outs.data() + f.slot * out_words + W * W, (size_t)W * W * 4);
            h->hm_cache_hits += 1;
            h->hm_ms = 0.f;
            f.answered = true;
            f.active = true;
            return;
        }
        h->hm_cache_misses += 1;
    }
    f.persistent = net.is_fc && net.fci_images && persist_wanted(h) && persist_start(h);
    if (f.persistent) fc_post(h, net);
    else graph_post(h, net);
    f.active = true;
}

static void hm_collect(pnn_handle* h) {
    pnn_handle::HmInflight& f = h->hm_inflight;
    if (!f.active) return;
    f.active = false;                       // (also when the wait below throws: the request is given up)
    if (f.answered) return;
    Net& net = *f.net;
    if (f.persistent) fc_collect(h, net);
    else graph_collect(h, net);
    if (h->hm_cache && net.cache.entries) {
        Net::HmCache& c = net.cache;
        const int W = net.W;
        c.tags[f.slot] = f.tag;
        memcpy(c.keys.data() + f.slot * c.key_words, h->hm_staged, c.key_words * 4);
        memcpy(c.outs.data() + f.slot * c.out_words, h->hm_out_raw, (size_t)W * W * 4);
        memcpy(c.outs.data() + f.slot * c.out_words + W * W, h->hm_out, (size_t)W * W * 4);
    }
}

// A request in flight whose answer nobody is waiting for yet: finish it so that the GPU, the doorbell and hm_staged are
// free again; the answer stays in the memo (the codec usually asks for it later: the RD pass repeats the fast pass's TU).
static void hm_drain(pnn_handle* h) {
    if (!h->hm_inflight.active) return;
    try {
        hm_collect(h);
    } catch (const std::exception& e) {
        h->error = e.what();
    }
}

static void run_hm(pnn_handle* h, Net& net) {
    hm_post(h, net);
    hm_collect(h);
}

// the net that serves in-loop calls of this width: fully-connected if one is loaded (the reference's choice for widths 4
// and 8, TComPrediction.cpp:564-566), else convolutional -- a convolutional net may also serve widths 4 and 8
static Net& hm_net(pnn_handle* h, int width) {
    bool has_fc;
    {
        std::unique_lock<std::mutex> lock(h->mu, std::defer_lock);
        if (h->bg_active.load()) lock.lock();
        has_fc = h->nets.count({width, 1}) != 0 || h->pending.count({width, 1}) != 0 || h->loading_key == std::make_pair(width, 1);
    }
    return *find_net(h, width, has_fc ? 1 : 0);
}

int pnn_predict_hm(pnn_handle* h, int width, int32_t* dst, int dst_stride) {
    if (!h) return -1;
    try {
        if (!dst) throw std::runtime_error("`piPred` is NULL.");
        if (h->hm_width != width) throw std::runtime_error("pnn_predict_hm called without a matching pnn_set_context");
        const int W = width;
        if (h->hm_inflight.active && h->hm_inflight.net->W == width && !h->hm_desc_pending) {
            hm_collect(h);                 // pnn_predict_hm_begin posted exactly this context
        } else {
            hm_drain(h);
            ensure_device(h);
            Net& net = hm_net(h, width);
            if (h->hm_desc_pending) stage_context(h);
            run_hm(h, net);
        }
        // reference TComPrediction.cpp(substitution):626-635: row-major copy with HM's stride
        for (int i = 0; i < W; ++i) memcpy(dst + (int64_t)i * dst_stride, h->hm_out + i * W, W * sizeof(int32_t));
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_predict_hm_begin(pnn_handle* h, int width) {
    if (!h) return -1;
    try {
        if (h->hm_width != width) throw std::runtime_error("pnn_predict_hm_begin called without a matching pnn_set_context");
        hm_drain(h);
        ensure_device(h);
        Net& net = hm_net(h, width);
        if (h->hm_desc_pending) stage_context(h);
        hm_post(h, net);
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_predict_hm_context(pnn_handle* h, int width, const float* above_or_flat, const float* left, float* out) {
    if (!h) return -1;
    try {
        hm_drain(h);
        if (!above_or_flat || !out) throw std::runtime_error("NULL buffer");
        if (width != 4 && width != 8 && width != 16 && width != 32 && width != 64) {
            throw std::runtime_error("the width of the TB does not belong to {4, 8, 16, 32, 64}");
        }
        ensure_device(h);
        Net& net = hm_net(h, width);
        const int W = width;
        // a convolutional net serving widths 4 / 8 reads its two portions from the two halves of the flattened context
        // (TComPattern.cpp:352-353)
        if (!net.is_fc && !left) left = above_or_flat + 3 * W * W;
        h->hm_staged[0] = h->hm_staged[1] = h->hm_staged[3] = 0;
        h->hm_staged[2] = 0;                                  // float mode (see pnn_internal.h)
        float* px = (float*)(h->hm_staged + HM_HEADER_INTS);
        if (net.is_fc) {
            memcpy(px, above_or_flat, (size_t)5 * W * W * sizeof(float));
        } else {
            memcpy(px, above_or_flat, (size_t)3 * W * W * sizeof(float));
            memcpy(px + 3 * W * W, left, (size_t)2 * W * W * sizeof(float));
        }
        h->hm_width = 0;                                       // a staged pnn_set_context is consumed
        h->hm_desc_pending = false;
        run_hm(h, net);
        memcpy(out, h->hm_out_raw, (size_t)W * W * sizeof(float));
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_set_workspace_budget(pnn_handle* h, int64_t bytes_per_net) {
    if (!h) return -1;
    if (bytes_per_net <= 0) {
        h->error = "the workspace budget must be positive";
        return -1;
    }
    h->workspace_budget = (size_t)bytes_per_net;
    for (auto& kv : h->nets) kv.second->cap_limit = (int64_t)1 << 40;     // a limit found under memory pressure is forgotten
    return 0;
}

int pnn_set_hm_cache(pnn_handle* h, int enabled) {
    if (!h) return -1;
    hm_drain(h);
    h->hm_cache = enabled != 0;
    if (!h->hm_cache) drop_hm_caches(h);
    return 0;
}

int pnn_hm_cache_stats(pnn_handle* h, int64_t* hits, int64_t* misses) {
    if (!h) return -1;
    if (hits) *hits = h->hm_cache_hits;
    if (misses) *misses = h->hm_cache_misses;
    return 0;
}

int pnn_predict_batch_device(pnn_handle* h, int width, int is_fc, const float* d_a, const float* d_l, int64_t n,
                             float* d_out, void* stream) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0) throw std::runtime_error("negative number of predictions");
        if (!d_a || !d_out || (!is_fc && !d_l)) throw std::runtime_error("NULL buffer");
        ensure_device(h);
        batch_device(h, *find_net(h, width, is_fc), d_a, d_l, n, d_out, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_predict_batch(pnn_handle* h, int width, int is_fc, const float* a, const float* l, int64_t n, float* out) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0) throw std::runtime_error("negative number of predictions");
        if (n == 0) return 0;
        if (!a || !out || (!is_fc && !l)) throw std::runtime_error("NULL buffer");
        ensure_device(h);
        Net& net = *find_net(h, width, is_fc);
        const int64_t px = (int64_t)width * width;
        const int64_t na = is_fc ? 5 * px : 3 * px;
        // chunk so that the staging buffers stay bounded
        const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1, ((int64_t)1 << 28) / (5 * px * 4)));
        h->d_in0.reserve((size_t)chunk * na * 4);
        if (!is_fc) h->d_in1.reserve((size_t)chunk * 2 * px * 4);
        ensure_workspace_fit(h, net, choose_capacity(h, net, chunk));
        DevBuf& d_out = net.out_raw;
        d_out.reserve((size_t)chunk * px * 4);
        for (int64_t s0 = 0; s0 < n; s0 += chunk) {
            const int64_t m = std::min(chunk, n - s0);
            CUDA_TRY(cudaMemcpyAsync(h->d_in0.p, a + s0 * na, (size_t)m * na * 4, cudaMemcpyHostToDevice, h->stream));
            if (!is_fc) {
                CUDA_TRY(cudaMemcpyAsync(h->d_in1.p, l + s0 * 2 * px, (size_t)m * 2 * px * 4, cudaMemcpyHostToDevice, h->stream));
            }
            batch_device(h, net, (const float*)h->d_in0.p, (const float*)h->d_in1.p, m, (float*)d_out.p, h->stream);
            CUDA_TRY(cudaMemcpyAsync(out + s0 * px, d_out.p, (size_t)m * px * 4, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_predict_image_blocks_device(pnn_handle* h, int width, int is_fc, const uint8_t* d_images, int n_images,
                                    int height, int width_image, const int32_t* d_idx, const int32_t* d_rows,
                                    const int32_t* d_cols, int64_t n, int mask_w, int mask_h, float* d_f32,
                                    uint8_t* d_u8, double* d_psnr, void* stream) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0) throw std::runtime_error("negative number of predictions");
        if (!d_images || !d_rows || !d_cols) throw std::runtime_error("NULL buffer");
        if (n_images > 1 && !d_idx) throw std::runtime_error("`image_index` is NULL while there are several images");
        check_masks(width, mask_w, mask_h);
        ensure_device(h);
        if (n_images <= 0 || height <= 0 || width_image <= 0) throw std::runtime_error("empty image set");
        image_blocks_device(h, *find_net(h, width, is_fc), d_images, n_images, height, width_image, d_idx, d_rows, d_cols, n, mask_w,
                            mask_h, d_f32, d_u8, d_psnr, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

static int image_blocks_host(pnn_handle* h, int width, int is_fc, const uint8_t* images, int n_images, int height,
                             int width_image, const int32_t* idx, const int32_t* rows, const int32_t* cols, int64_t n,
                             int mask_w, int mask_h, float* out_f32, uint8_t* out_u8, double* out_psnr, bool wait) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        if (n < 0) throw std::runtime_error("negative number of predictions");
        if (!images || !rows || !cols) throw std::runtime_error("NULL buffer");
        if (n_images > 1 && !idx) throw std::runtime_error("`image_index` is NULL while there are several images");
        if (n_images <= 0 || height <= 0 || width_image <= 0) throw std::runtime_error("empty image set");
        check_masks(width, mask_w, mask_h);
        for (int64_t i = 0; i < n; ++i) {
            // the target block and its above-left context anchor must lie inside the image
            // (reference sets/common.py:82-91 raises for windows leaving the image; here only the
            // above-right / below-left parts may leave it, and they are masked)
            if (rows[i] < width || cols[i] < width || rows[i] + width > height || cols[i] + width > width_image) {
                throw std::runtime_error("block " + std::to_string(i) + " or its context anchor lies outside the image");
            }
            if (idx && (idx[i] < 0 || idx[i] >= n_images)) throw std::runtime_error("`image_index` out of range");
        }
        if (n == 0) return 0;
        ensure_device(h);
        Net& net = *find_net(h, width, is_fc);
        const int64_t px = (int64_t)width * width;
        cudaStream_t s = h->stream;
        const size_t img_bytes = (size_t)n_images * height * width_image;
        // upload on stream_in into the input set that the call before the previous one used
        pnn_handle::InputSet& in = h->in_set[h->in_next];
        h->in_next ^= 1;
        in.images.reserve(img_bytes);          // growing frees the old buffer, which waits for the device
        in.rows.reserve(n * 4);
        in.cols.reserve(n * 4);
        if (idx) in.idx.reserve(n * 4);
        if (in.used) CUDA_TRY(cudaStreamWaitEvent(h->stream_in, in.consumed, 0));
        CUDA_TRY(cudaMemcpyAsync(in.images.p, images, img_bytes, cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(in.rows.p, rows, n * 4, cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaMemcpyAsync(in.cols.p, cols, n * 4, cudaMemcpyHostToDevice, h->stream_in));
        if (idx) CUDA_TRY(cudaMemcpyAsync(in.idx.p, idx, n * 4, cudaMemcpyHostToDevice, h->stream_in));
        CUDA_TRY(cudaEventRecord(in.uploaded, h->stream_in));
        CUDA_TRY(cudaStreamWaitEvent(s, in.uploaded, 0));
        // outputs are produced chunk by chunk into the net's own output buffers
        const int64_t cap = ensure_workspace_fit(h, net, choose_capacity(h, net, n));
        if (!net.computed) {
            CUDA_TRY(cudaEventCreateWithFlags(&net.computed, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&net.read_back, cudaEventDisableTiming));
        }
        if (net.read_back_pending) CUDA_TRY(cudaStreamWaitEvent(s, net.read_back, 0));   // out_* still being read
        for (int64_t s0 = 0; s0 < n; s0 += cap) {
            const int64_t m = std::min(cap, n - s0);
            const bool last = s0 + cap >= n;
            image_blocks_device(h, net, (const uint8_t*)in.images.p, n_images, height, width_image,
                                idx ? (const int32_t*)in.idx.p + s0 : nullptr, (const int32_t*)in.rows.p + s0,
                                (const int32_t*)in.cols.p + s0, m, mask_w, mask_h,
                                out_f32 ? (float*)net.out_raw.p : nullptr,
                                (out_u8 || out_psnr) ? (uint8_t*)net.out_u8.p : nullptr,
                                out_psnr ? (double*)net.out_psnr.p : nullptr, s);
            // the read-back of the last chunk goes to stream_out (it overlaps the next call); earlier chunks stay in order
            cudaStream_t so = s;
            if (last) {
                CUDA_TRY(cudaEventRecord(net.computed, s));
                CUDA_TRY(cudaStreamWaitEvent(h->stream_out, net.computed, 0));
                so = h->stream_out;
            }
            if (out_f32) CUDA_TRY(cudaMemcpyAsync(out_f32 + s0 * px, net.out_raw.p, (size_t)m * px * 4, cudaMemcpyDeviceToHost, so));
            if (out_u8) CUDA_TRY(cudaMemcpyAsync(out_u8 + s0 * px, net.out_u8.p, (size_t)m * px, cudaMemcpyDeviceToHost, so));
            if (out_psnr) CUDA_TRY(cudaMemcpyAsync(out_psnr + s0, net.out_psnr.p, (size_t)m * 8, cudaMemcpyDeviceToHost, so));
        }
        CUDA_TRY(cudaEventRecord(net.read_back, h->stream_out));
        net.read_back_pending = true;
        CUDA_TRY(cudaEventRecord(in.consumed, s));
        in.used = true;
        if (wait) {
            CUDA_TRY(cudaStreamSynchronize(h->stream_out));   // ordered after the kernels of this call
            net.read_back_pending = false;
        }
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

int pnn_predict_image_blocks(pnn_handle* h, int width, int is_fc, const uint8_t* images, int n_images, int height,
                             int width_image, const int32_t* idx, const int32_t* rows, const int32_t* cols, int64_t n,
                             int mask_w, int mask_h, float* out_f32, uint8_t* out_u8, double* out_psnr) {
    return image_blocks_host(h, width, is_fc, images, n_images, height, width_image, idx, rows, cols, n, mask_w, mask_h, out_f32,
                             out_u8, out_psnr, true);
}

int pnn_predict_image_blocks_async(pnn_handle* h, int width, int is_fc, const uint8_t* images, int n_images, int height,
                                   int width_image, const int32_t* idx, const int32_t* rows, const int32_t* cols, int64_t n,
                                   int mask_w, int mask_h, float* out_f32, uint8_t* out_u8, double* out_psnr) {
    return image_blocks_host(h, width, is_fc, images, n_images, height, width_image, idx, rows, cols, n, mask_w, mask_h, out_f32,
                             out_u8, out_psnr, false);
}

int pnn_synchronize(pnn_handle* h) {
    if (!h) return -1;
    Quiesce quiesce(h);
    try {
        ensure_device(h);
        CUDA_TRY(cudaStreamSynchronize(h->stream_in));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream_out));
        for (auto& kv : h->nets) kv.second->read_back_pending = false;
    } catch (const std::exception& e) {
        return fail(h, e);
    }
    return 0;
}

}  // extern "C"
