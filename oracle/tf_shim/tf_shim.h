// Oracle build support (TEST INFRASTRUCTURE ONLY).
//
// A minimal stand-in for the handful of TensorFlow 1.x C++ framework types that
// the reference's two custom-op sources touch, so that
//   /root/reference/octbit/octbit_mat_mul_op.cc
//   /root/reference/octbit/octbit_ops_reg.cc
//   /root/reference/positional_encoding/positional_encoding_op.cc
// compile UNMODIFIED, from where they lie, into oracle/_ref/libkws_ref_ops.so
// (see oracle/Makefile).  No reference source is copied into this repository;
// this header only supplies the framework surface (Tensor accessors, OpKernel,
// OP_REQUIRES, REGISTER_*), written from the op sources' usage, not from TF.
#ifndef KWS_ORACLE_TF_SHIM_H_
#define KWS_ORACLE_TF_SHIM_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {
typedef long DenseIndex;
template <typename I>
struct IndexPair {
  I first;
  I second;
};
template <typename T, int N>
struct array {
  T v[N];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
}  // namespace Eigen

namespace tensorflow {

typedef unsigned char uint8;
typedef signed char int8;
typedef int int32;
typedef long long int64;
typedef unsigned long long uint64;
using std::string;

struct qint8 {
  int8 value;
  qint8() : value(0) {}
  qint8(const int8 v) : value(v) {}
  operator int() const { return static_cast<int>(value); }
};
struct quint8 {
  uint8 value;
  quint8() : value(0) {}
  quint8(const uint8 v) : value(v) {}
  operator int() const { return static_cast<int>(value); }
};

// ---------------------------------------------------------------- Status
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline void Append(std::ostringstream&) {}
template <typename T, typename... R>
void Append(std::ostringstream& os, const T& v, const R&... r) {
  os << v;
  Append(os, r...);
}
template <typename... A>
Status InvalidArgument(const A&... a) {
  std::ostringstream os;
  os << "InvalidArgument: ";
  Append(os, a...);
  return Status(os.str());
}
}  // namespace errors

// ---------------------------------------------------------------- shapes
class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : dims_(d) {}
  int dims() const { return static_cast<int>(dims_.size()); }
  int64 dim_size(int i) const { return dims_[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (size_t i = 0; i < dims_.size(); ++i) n *= dims_[i];
    return n;
  }
  std::string DebugString() const {
    std::ostringstream os;
    os << "[";
    for (size_t i = 0; i < dims_.size(); ++i) os << (i ? "," : "") << dims_[i];
    os << "]";
    return os.str();
  }

 private:
  std::vector<int64> dims_;
};

struct TensorShapeUtils {
  static bool IsMatrix(const TensorShape& s) { return s.dims() == 2; }
  static bool IsScalar(const TensorShape& s) { return s.dims() == 0; }
};

// ---------------------------------------------------------------- Tensor
namespace shim {
template <typename T>
struct FlatView {
  T* p;
  int64 n;
  T* data() const { return p; }
  T& operator()(int64 i) const { return p[i]; }
  int64 size() const { return n; }
};
template <typename T>
struct MatrixView {
  T* p;
  int64 rows, cols;
  T& operator()(int64 i, int64 j) const { return p[i * cols + j]; }
};
template <typename T>
struct ScalarView {
  T* p;
  T& operator()() const { return *p; }
  T& operator()(int64) const { return *p; }
};
}  // namespace shim

class Tensor {
 public:
  Tensor() : bytes_(0), data_(nullptr) {}
  Tensor(const TensorShape& shape, size_t elem_bytes) : shape_(shape), data_(nullptr) {
    bytes_ = static_cast<size_t>(shape.num_elements()) * elem_bytes;
    size_t padded = (bytes_ + 63) / 64 * 64 + 64;
    void* p = nullptr;
    if (posix_memalign(&p, 64, padded) != 0) p = nullptr;
    data_ = p;
    if (data_) std::memset(data_, 0, padded);
  }
  Tensor(const Tensor& o) : bytes_(0), data_(nullptr) { *this = o; }
  Tensor& operator=(const Tensor& o) {
    if (this == &o) return *this;
    release();
    shape_ = o.shape_;
    bytes_ = o.bytes_;
    if (o.data_) {
      size_t padded = (bytes_ + 63) / 64 * 64 + 64;
      void* p = nullptr;
      if (posix_memalign(&p, 64, padded) != 0) p = nullptr;
      data_ = p;
      std::memcpy(data_, o.data_, bytes_);
    }
    return *this;
  }
  ~Tensor() { release(); }

  const TensorShape& shape() const { return shape_; }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  int dims() const { return shape_.dims(); }
  void* raw() const { return data_; }
  size_t bytes() const { return bytes_; }

  template <typename T>
  shim::FlatView<T> flat() const {
    return shim::FlatView<T>{static_cast<T*>(data_), shape_.num_elements()};
  }
  template <typename T>
  shim::FlatView<T> vec() const {
    return flat<T>();
  }
  template <typename T>
  shim::MatrixView<T> matrix() const {
    return shim::MatrixView<T>{static_cast<T*>(data_), shape_.dim_size(0), shape_.dim_size(1)};
  }
  template <typename T>
  shim::ScalarView<T> scalar() const {
    return shim::ScalarView<T>{static_cast<T*>(data_)};
  }

 private:
  void release() {
    if (data_) free(data_);
    data_ = nullptr;
  }
  TensorShape shape_;
  size_t bytes_;
  void* data_;
};

// ---------------------------------------------------------------- kernels
class OpKernelConstruction {
 public:
  std::map<std::string, bool> bool_attrs;
  std::map<std::string, float> float_attrs;
  std::map<std::string, int> int_attrs;
  std::map<std::string, Tensor> tensor_attrs;

  Status GetAttr(const std::string& n, bool* v) const { return get(bool_attrs, n, v); }
  Status GetAttr(const std::string& n, float* v) const { return get(float_attrs, n, v); }
  Status GetAttr(const std::string& n, int* v) const { return get(int_attrs, n, v); }
  Status GetAttr(const std::string& n, Tensor* v) const { return get(tensor_attrs, n, v); }

  void CtxFailure(const Status& s) {
    if (status_.ok()) status_ = s;
  }
  void CtxFailure(const char*, int, const Status& s) { CtxFailure(s); }
  const Status& status() const { return status_; }

 private:
  template <typename M, typename T>
  static Status get(const M& m, const std::string& n, T* v) {
    typename M::const_iterator it = m.find(n);
    if (it == m.end()) return Status("NotFound: attr " + n);
    *v = it->second;
    return Status::OK();
  }
  Status status_;
};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::vector<std::string> input_names;
  std::vector<Tensor*> outputs;
  std::vector<std::string> output_names;

  ~OpKernelContext() {
    for (size_t i = 0; i < outputs.size(); ++i) delete outputs[i];
  }
  const Tensor& input(int i) const { return inputs[i]; }
  Status input(const std::string& name, const Tensor** t) const {
    for (size_t i = 0; i < input_names.size(); ++i)
      if (input_names[i] == name) {
        *t = &inputs[i];
        return Status::OK();
      }
    return Status("NotFound: input " + name);
  }
  Status allocate_output(int idx, const TensorShape& shape, Tensor** out) {
    if (static_cast<int>(outputs.size()) <= idx) outputs.resize(idx + 1, nullptr);
    delete outputs[idx];
    outputs[idx] = new Tensor(shape, 4);  // both reference ops emit float32
    *out = outputs[idx];
    return Status::OK();
  }
  Status allocate_output(const std::string& name, const TensorShape& shape, Tensor** out) {
    for (size_t i = 0; i < output_names.size(); ++i)
      if (output_names[i] == name) return allocate_output(static_cast<int>(i), shape, out);
    return Status("NotFound: output " + name);
  }
  void CtxFailure(const Status& s) {
    if (status_.ok()) status_ = s;
  }
  void CtxFailure(const char*, int, const Status& s) { CtxFailure(s); }
  const Status& status() const { return status_; }

 private:
  Status status_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)                     \
  do {                                               \
    ::tensorflow::Status _s(__VA_ARGS__);            \
    if (!_s.ok()) {                                  \
      (CTX)->CtxFailure(_s);                         \
      return;                                        \
    }                                                \
  } while (0)

#define TF_RETURN_IF_ERROR(...)                      \
  do {                                               \
    ::tensorflow::Status _status = (__VA_ARGS__);    \
    if (!_status.ok()) return _status;               \
  } while (0)

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::abort();                                                          \
    }                                                                        \
  } while (0)

#define TF_DISALLOW_COPY_AND_ASSIGN(TypeName) \
  TypeName(const TypeName&) = delete;         \
  void operator=(const TypeName&) = delete

// ---------------------------------------------------------------- registry
static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

namespace shim {
typedef std::function<OpKernel*(OpKernelConstruction*)> KernelFactory;
std::map<std::string, KernelFactory>& KernelRegistry();

struct KernelDefBuilder {
  std::string name;
  std::string device;
  explicit KernelDefBuilder(const char* n) : name(n) {}
  KernelDefBuilder& Device(const char* d) {
    device = d;
    return *this;
  }
};
struct KernelRegistrar {
  KernelRegistrar(const KernelDefBuilder& b, KernelFactory f) { KernelRegistry()[b.name] = f; }
};
}  // namespace shim

inline shim::KernelDefBuilder Name(const char* n) { return shim::KernelDefBuilder(n); }

#define KWS_SHIM_CAT2(a, b) a##b
#define KWS_SHIM_CAT(a, b) KWS_SHIM_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(builder, ...)                                     \
  static ::tensorflow::shim::KernelRegistrar KWS_SHIM_CAT(_kws_kreg_, __COUNTER__)( \
      (builder), [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { \
        return new __VA_ARGS__(c);                                                \
      })

// ---------------------------------------------------------------- op defs
namespace shape_inference {
struct DimensionHandle {
  int64 v;
  DimensionHandle() : v(-1) {}
  DimensionHandle(int64 x) : v(x) {}
};
struct ShapeHandle {
  std::vector<int64> dims;
  bool known_rank;
  ShapeHandle() : known_rank(false) {}
};
class InferenceContext {
 public:
  static const int64 kUnknownDim = -1;
  std::vector<ShapeHandle> inputs_;
  std::vector<ShapeHandle> outputs_;
  std::map<std::string, int> int_attrs;
  ShapeHandle input(int i) const { return inputs_[i]; }
  Status WithRank(const ShapeHandle& s, int rank, ShapeHandle* out) {
    if (s.known_rank && static_cast<int>(s.dims.size()) != rank)
      return errors::InvalidArgument("Shape must be rank ", rank);
    *out = s;
    return Status::OK();
  }
  Status GetAttr(const std::string& n, int32* v) const {
    std::map<std::string, int>::const_iterator it = int_attrs.find(n);
    if (it == int_attrs.end()) return Status("NotFound: attr " + n);
    *v = it->second;
    return Status::OK();
  }
  ShapeHandle Matrix(DimensionHandle a, DimensionHandle b) {
    ShapeHandle s;
    s.known_rank = true;
    s.dims.push_back(a.v);
    s.dims.push_back(b.v);
    return s;
  }
  void set_output(int i, const ShapeHandle& s) {
    if (static_cast<int>(outputs_.size()) <= i) outputs_.resize(i + 1);
    outputs_[i] = s;
  }
};
}  // namespace shape_inference

namespace shim {
struct OpDef {
  std::string name;
  std::vector<std::string> inputs, outputs, attrs;
  std::function<Status(shape_inference::InferenceContext*)> shape_fn;
};
std::map<std::string, OpDef>& OpRegistry();

class OpDefBuilder {
 public:
  explicit OpDefBuilder(const char* name) { def_.name = name; publish(); }
  OpDefBuilder& Input(const std::string& s) { def_.inputs.push_back(s); return publish(); }
  OpDefBuilder& Output(const std::string& s) { def_.outputs.push_back(s); return publish(); }
  OpDefBuilder& Attr(const std::string& s) { def_.attrs.push_back(s); return publish(); }
  OpDefBuilder& Doc(const std::string&) { return *this; }
  OpDefBuilder& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)> f) {
    def_.shape_fn = f;
    return publish();
  }

 private:
  OpDefBuilder& publish() {
    OpRegistry()[def_.name] = def_;
    return *this;
  }
  OpDef def_;
};
}  // namespace shim

#define REGISTER_OP(name) \
  static ::tensorflow::shim::OpDefBuilder KWS_SHIM_CAT(_kws_opreg_, __COUNTER__) = ::tensorflow::shim::OpDefBuilder(name)

}  // namespace tensorflow

#endif  // KWS_ORACLE_TF_SHIM_H_
