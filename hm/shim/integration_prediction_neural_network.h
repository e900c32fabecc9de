// Link seam for the reference's HM builds (hm_16_15_substitution / hm_16_15_switch).
//
// The reference's codec reaches TensorFlow only through "integration_prediction_neural_network.h"
// (hevc/hm_common/c++/source_common/integration_prediction_neural_network.h:1-78, included from
// TComPrediction.h:45).  This header has the same name and is put FIRST on the include path, so the
// reference's TComPrediction.{h,cpp} and TComPattern.cpp compile UNMODIFIED; it declares the small part of
// the TensorFlow C++ API those files use (Tensor, Session::Run, Status, string, LOG) and the four functions
// of the reference header, all implemented in pnn_hm_shim.cpp on top of libpnn_cuda (include/pnn_cuda.h).
#ifndef INTEGRATION_PREDICTION_NEURAL_NETWORK_H
#define INTEGRATION_PREDICTION_NEURAL_NETWORK_H

#include <cstddef>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace tensorflow {

typedef std::string string;

class Status {
public:
    Status() : ok_(true) {}
    explicit Status(const std::string& message) : ok_(false), message_(message) {}
    static Status OK() { return Status(); }
    bool ok() const { return ok_; }
    const std::string& error_message() const { return message_; }
private:
    bool ok_;
    std::string message_;
};
std::ostream& operator<<(std::ostream& os, const Status& status);

enum DataType { DT_FLOAT = 1 };

class TensorShape {
public:
    TensorShape() {}
    TensorShape(std::initializer_list<long long> dims) : dims_(dims) {}
    int dims() const { return static_cast<int>(dims_.size()); }
    long long dim_size(int i) const { return dims_.at(i); }
    long long num_elements() const {
        long long n(1);
        for (std::size_t i(0); i < dims_.size(); i++) n *= dims_[i];
        return n;
    }
private:
    std::vector<long long> dims_;
};

// float32 host tensor; copies share the buffer like tensorflow::Tensor does
class Tensor {
public:
    struct Flat {
        float* ptr;
        std::size_t count;
        float* data() const { return ptr; }
        std::size_t size() const { return count; }
        float& operator()(std::size_t i) const { return ptr[i]; }
    };
    Tensor() {}
    Tensor(DataType, const TensorShape& shape)
        : shape_(shape), buffer_(new std::vector<float>(static_cast<std::size_t>(shape.num_elements()), 0.f)) {}
    const TensorShape& shape() const { return shape_; }
    int dims() const { return shape_.dims(); }
    long long dim_size(int i) const { return shape_.dim_size(i); }
    template <typename T>
    Flat flat() const {
        Flat f;
        f.ptr = buffer_ ? buffer_->data() : NULL;
        f.count = buffer_ ? buffer_->size() : 0;
        return f;
    }
private:
    TensorShape shape_;
    std::shared_ptr<std::vector<float> > buffer_;
};

// One prediction neural network loaded into libpnn_cuda
class Session {
public:
    Session(int width_target, bool is_fully_connected) : width_(width_target), is_fc_(is_fully_connected) {}
    Status Run(const std::vector<std::pair<string, Tensor> >& inputs,
               const std::vector<string>& output_tensor_names,
               const std::vector<string>& target_node_names,
               std::vector<Tensor>* outputs);
private:
    int width_;
    bool is_fc_;
};

}  // namespace tensorflow

// `LOG(ERROR) << status` (TComPrediction.cpp:176, 582)
#define LOG(severity) std::cerr

// reference integration_prediction_neural_network.h:36-76
void create_tensors_context_portion(std::vector<tensorflow::Tensor>& tensors_portion_above,
                                    std::vector<tensorflow::Tensor>& tensors_portion_left);
void create_tensors_flattened_context(std::vector<tensorflow::Tensor>& tensors_flattened_context);
tensorflow::Status load_graph(const tensorflow::string& path_to_graph_output,
                              std::unique_ptr<tensorflow::Session>& unique_ptr_session);
tensorflow::Status load_graphs(const std::vector<std::string>& vector_paths_to_graphs_output,
                               std::vector<std::unique_ptr<tensorflow::Session> >& vector_unique_ptrs_session);

#endif
