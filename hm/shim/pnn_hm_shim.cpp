// Implementation of the link seam (see the two headers of this directory) on top of libpnn_cuda.
#include "integration_prediction_neural_network.h"
#include "interface_c_python.h"

#include "pnn_cuda.h"

#include <errno.h>    // program_invocation_short_name (glibc)

#include <chrono>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <fstream>

namespace {

pnn_handle* g_handle = NULL;
long long g_calls[5] = {0, 0, 0, 0, 0};
double g_seconds[5] = {0., 0., 0., 0., 0.};
double g_first_seconds[5] = {0., 0., 0., 0., 0.};   // the first call of a width: device initialisation (first width) and the upload of the net

void print_stats() {
    // PNN call statistics of the process: calls and wall time per block width
    const char* path = getenv("PNN_HM_STATS");
    FILE* f = path ? fopen(path, "w") : stderr;
    if (!f) f = stderr;
    long long total(0);
    double seconds(0.);
    for (int i(0); i < 5; i++) {
        // the first call of a width waits for the device initialisation / the upload of its net: counted in the total, kept out
        // of the per-call figure
        const long long later(g_calls[i] > 1 ? g_calls[i] - 1 : 0);
        fprintf(f, "pnn_calls width %d: %lld calls, %.6f s, %.2f us/call (first call %.1f ms, not in the per-call figure)\n", 4 << i,
                g_calls[i], g_seconds[i] + g_first_seconds[i], later ? 1.e6 * g_seconds[i] / later : 0., 1.e3 * g_first_seconds[i]);
        total += g_calls[i];
        seconds += g_seconds[i] + g_first_seconds[i];
    }
    fprintf(f, "pnn_calls total: %lld calls, %.6f s\n", total, seconds);
    if (g_handle) {
        int64_t hits(0), misses(0);
        pnn_hm_cache_stats(g_handle, &hits, &misses);
        fprintf(f, "pnn_cache: %lld hits, %lld misses\n", (long long)hits, (long long)misses);
    }
    if (f != stderr) fclose(f);
    if (g_handle) {
        pnn_release_at_exit(g_handle);      // the process ends here: no buffer-by-buffer teardown
        g_handle = NULL;
    }
}

pnn_handle* handle() {
    if (!g_handle) {
        // The contexts reach the library already mean-centred and the outputs leave it raw, as with
        // Session::Run, so the mean given here is not used; the QP test of TComPrediction.cpp:156 has
        // already selected the paths.
        int device(0);
        const char* env = getenv("PNN_DEVICE");
        if (env) device = atoi(env);
        if (pnn_create_deferred(NULL, 0.f, 1, device, &g_handle) != 0) {
            fprintf(stderr, "%s\n", pnn_last_error(NULL));
            return NULL;
        }
        // the codec evaluates the same unit with the same context in its fast pass, its RD pass and the final
        // reconstruction: answer repeated contexts from the library's memo (PNN_HM_CACHE=0 switches it off)
        const char* cache = getenv("PNN_HM_CACHE");
        pnn_set_hm_cache(g_handle, cache ? atoi(cache) : 1);
        atexit(print_stats);
    }
    return g_handle;
}

}  // namespace

namespace tensorflow {

std::ostream& operator<<(std::ostream& os, const Status& status) {
    return os << (status.ok() ? "OK" : status.error_message());
}

Status Session::Run(const std::vector<std::pair<string, Tensor> >& inputs, const std::vector<string>&,
                    const std::vector<string>&, std::vector<Tensor>* outputs) {
    if (!outputs) return Status("`outputs` is NULL");
    const float* above_or_flat(NULL);
    const float* left(NULL);
    if (inputs.size() == 1) {
        // flattened context (widths 4 and 8); a convolutional net loaded for such a width reads its two
        // portions from the two halves of the flattened context (handled by the library)
        above_or_flat = inputs[0].second.flat<float>().data();
    } else {
        if (inputs.size() != 2) return Status("a convolutional PNN takes two inputs");
        for (std::size_t i(0); i < 2; i++) {
            if (inputs[i].first == "node_portion_above") above_or_flat = inputs[i].second.flat<float>().data();
            else if (inputs[i].first == "node_portion_left") left = inputs[i].second.flat<float>().data();
        }
        if (!above_or_flat || !left) return Status("unknown input node names");
    }
    Tensor prediction(DT_FLOAT, TensorShape({1, width_, width_, 1}));
    const std::chrono::steady_clock::time_point t0(std::chrono::steady_clock::now());
    const int code(pnn_predict_hm_context(handle(), width_, above_or_flat, left, prediction.flat<float>().data()));
    const double dt(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    if (code != 0) return Status(pnn_last_error(g_handle));
    const int index(static_cast<int>(std::log2(static_cast<double>(width_))) - 2);
    if (g_calls[index] == 0) g_first_seconds[index] = dt;
    else g_seconds[index] += dt;
    g_calls[index] += 1;
    outputs->clear();
    outputs->push_back(prediction);
    return Status::OK();
}

}  // namespace tensorflow

// reference integration_prediction_neural_network.cpp:3-27: batch-1 input tensors of the five nets
void create_tensors_context_portion(std::vector<tensorflow::Tensor>& tensors_portion_above,
                                    std::vector<tensorflow::Tensor>& tensors_portion_left) {
    for (long long width = 16; width <= 64; width *= 2) {
        tensors_portion_above.push_back(tensorflow::Tensor(tensorflow::DT_FLOAT, tensorflow::TensorShape({1, width, 3 * width, 1})));
        tensors_portion_left.push_back(tensorflow::Tensor(tensorflow::DT_FLOAT, tensorflow::TensorShape({1, 2 * width, width, 1})));
    }
}

void create_tensors_flattened_context(std::vector<tensorflow::Tensor>& tensors_flattened_context) {
    for (long long width = 4; width <= 8; width *= 2) {
        tensors_flattened_context.push_back(tensorflow::Tensor(tensorflow::DT_FLOAT, tensorflow::TensorShape({1, 5 * width * width})));
    }
}

// reference integration_prediction_neural_network.cpp:29-54: one frozen graph -> one session.  The path may name
// the frozen graph itself (the binary GraphDef `graph_output.pbtxt`; libpnn_cuda reads its constants) or a PNNW flat
// binary; when a ".pnnw" sibling of a ".pbtxt" path exists it is preferred (faster to parse).
tensorflow::Status load_graph(const tensorflow::string& path_to_graph_output,
                              std::unique_ptr<tensorflow::Session>& unique_ptr_session) {
    std::string path(path_to_graph_output);
    const std::string suffix(".pbtxt");
    if (path.size() > suffix.size() && path.compare(path.size() - suffix.size(), suffix.size(), suffix) == 0) {
        const std::string sibling(path.substr(0, path.size() - suffix.size()) + ".pnnw");
        if (std::ifstream(sibling.c_str(), std::ios::binary).good()) path = sibling;
    }
    int width(0), is_fc(0);
    if (pnn_inspect_net_file(path.c_str(), &width, &is_fc, NULL, NULL) != 0) {
        return tensorflow::Status("Failed to load the weights at \"" + path + "\": " + pnn_last_error(NULL));
    }
    pnn_handle* h(handle());
    if (!h) return tensorflow::Status("libpnn_cuda could not be initialised");
    // validated now, uploaded the first time the codec needs this width
    if (pnn_register_net(h, path.c_str()) != 0) return tensorflow::Status(pnn_last_error(h));
    unique_ptr_session.reset(new tensorflow::Session(static_cast<int>(width), is_fc != 0));
    return tensorflow::Status::OK();
}

// reference integration_prediction_neural_network.cpp:56-69
tensorflow::Status load_graphs(const std::vector<std::string>& vector_paths_to_graphs_output,
                               std::vector<std::unique_ptr<tensorflow::Session> >& vector_unique_ptrs_session) {
    if (vector_paths_to_graphs_output.size() != vector_unique_ptrs_session.size()) {
        return tensorflow::Status("The number of paths is not equal to the number of sessions.");
    }
    for (std::size_t i(0); i < vector_paths_to_graphs_output.size(); i++) {
        const tensorflow::Status status(load_graph(vector_paths_to_graphs_output[i], vector_unique_ptrs_session[i]));
        if (!status.ok()) return status;
    }
    // An encoder will need every net, but not before it has coded its first row of coding tree units (no causal context
    // there): device initialisation and uploads start now, on a thread of the library.  A decoder may never need them.
    const char* warm = getenv("PNN_HM_WARM_UP");
    const bool is_encoder(program_invocation_short_name && strstr(program_invocation_short_name, "Encoder") != NULL);
    if (g_handle && (warm ? atoi(warm) != 0 : is_encoder)) pnn_warm_up(g_handle);
    return tensorflow::Status::OK();
}

// ---------------------------------------------------------------------------------------------------------
// the CPython names used by TComPrediction.cpp:180-236
// ---------------------------------------------------------------------------------------------------------
static int g_python_initialized = 0;
static PyObject g_callable = {false, 0.};

void Py_Initialize() { g_python_initialized = 1; }
int Py_IsInitialized() { return g_python_initialized; }
void Py_Finalize() { g_python_initialized = 0; }
void Py_DECREF(PyObject* object) {
    if (object && object != &g_callable) delete object;
}
int PyFloat_CheckExact(PyObject* object) { return object && object->is_float; }
double PyFloat_AsDouble(PyObject* object) { return object ? object->value : -1.; }
PyObject* PyErr_Occurred() { return NULL; }
void PyErr_Print() {}

int append_sys_path(const std::string&) { return 0; }
PyObject* get_callable(const std::string&, const std::string&) { return &g_callable; }

// The file holds a pickled Python float (the reference's sets/results/training_set/means/luminance/
// mean_training.pkl, protocol 2: 0x80 0x02 'G' + 8 big-endian bytes + '.'; protocol 0: 'F' + repr + '\n.')
// or a plain text float.
PyObject* load_via_pickle(PyObject*, const std::string& path_to_file) {
    std::ifstream file(path_to_file.c_str(), std::ios::binary);
    if (!file) {
        fprintf(stderr, "The file at \"%s\" cannot be opened.\n", path_to_file.c_str());
        return NULL;
    }
    std::string data((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
    std::size_t pos(0);
    if (data.size() >= 2 && static_cast<unsigned char>(data[0]) == 0x80) pos = 2;     // PROTO opcode
    PyObject* object(new PyObject);
    object->is_float = false;
    object->value = 0.;
    if (pos < data.size() && data[pos] == 'G' && pos + 9 <= data.size()) {           // BINFLOAT
        uint64_t bits(0);
        for (int i(0); i < 8; i++) bits = (bits << 8) | static_cast<unsigned char>(data[pos + 1 + i]);
        memcpy(&object->value, &bits, 8);
        object->is_float = true;
    } else {
        if (pos < data.size() && data[pos] == 'F') pos += 1;                           // FLOAT (text)
        char* end(NULL);
        const double value(strtod(data.c_str() + pos, &end));
        if (end != data.c_str() + pos) {
            object->value = value;
            object->is_float = true;
        }
    }
    return object;
}
