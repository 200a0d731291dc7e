// Oracle build support (TEST INFRASTRUCTURE ONLY).
// C entry points that run the reference's own, unmodified op kernels
// (compiled from /root/reference into the same .so, see oracle/Makefile)
// through the shimmed OpKernel interface.
#include "tf_shim.h"

namespace tensorflow {
namespace shim {
std::map<std::string, KernelFactory>& KernelRegistry() {
  static std::map<std::string, KernelFactory> r;
  return r;
}
std::map<std::string, OpDef>& OpRegistry() {
  static std::map<std::string, OpDef> r;
  return r;
}
}  // namespace shim
}  // namespace tensorflow

using namespace tensorflow;

static void set_err(char* err, int errlen, const std::string& m) {
  if (err && errlen > 0) {
    std::snprintf(err, static_cast<size_t>(errlen), "%s", m.c_str());
  }
}

extern "C" {

// OctbitMatMul (octbit/octbit_mat_mul_op.cc:34-190).  Attributes are passed as
// given so that the constructor's own checks (:41-46) fire.  Returns 0 on
// success, 1 on constructor failure, 2 on Compute failure, 3 if unregistered.
int ref_octbit_matmul(const float* x, const signed char* w, const float* bias, float scale,
                      int transpose_a, int transpose_b, int A, int B, int K, int Kw,
                      float* out, char* err, int errlen) {
  std::map<std::string, shim::KernelFactory>::iterator it =
      shim::KernelRegistry().find("OctbitMatMul");
  if (it == shim::KernelRegistry().end()) {
    set_err(err, errlen, "OctbitMatMul kernel not registered");
    return 3;
  }
  OpKernelConstruction cons;
  cons.bool_attrs["transpose_a"] = transpose_a != 0;
  cons.bool_attrs["transpose_b"] = transpose_b != 0;
  cons.float_attrs["scale"] = scale;
  Tensor bias_t(TensorShape({static_cast<int64>(B)}), sizeof(float));
  std::memcpy(bias_t.raw(), bias, sizeof(float) * static_cast<size_t>(B));
  cons.tensor_attrs["bias"] = bias_t;
  OpKernel* k = it->second(&cons);
  if (!cons.status().ok()) {
    set_err(err, errlen, cons.status().error_message());
    delete k;
    return 1;
  }
  OpKernelContext ctx;
  ctx.inputs.push_back(Tensor(TensorShape({static_cast<int64>(A), static_cast<int64>(K)}), sizeof(float)));
  ctx.inputs.push_back(Tensor(TensorShape({static_cast<int64>(B), static_cast<int64>(Kw)}), 1));
  ctx.input_names.push_back("input_a");
  ctx.input_names.push_back("input_b");
  ctx.output_names.push_back("output");
  std::memcpy(ctx.inputs[0].raw(), x, sizeof(float) * static_cast<size_t>(A) * K);
  std::memcpy(ctx.inputs[1].raw(), w, static_cast<size_t>(B) * Kw);
  k->Compute(&ctx);
  int rc = 0;
  if (!ctx.status().ok()) {
    set_err(err, errlen, ctx.status().error_message());
    rc = 2;
  } else {
    std::memcpy(out, ctx.outputs[0]->raw(), sizeof(float) * static_cast<size_t>(A) * B);
  }
  delete k;
  return rc;
}

// PositionalEncoding (positional_encoding/positional_encoding_op.cc:26-55).
// `out` must be pre-filled by the caller: for odd encoding_size the reference
// never writes the last column (:45), which the caller can observe.
int ref_positional_encoding(int max_position, int encoding_size, float* out, char* err, int errlen) {
  std::map<std::string, shim::KernelFactory>::iterator it =
      shim::KernelRegistry().find("PositionalEncoding");
  if (it == shim::KernelRegistry().end()) {
    set_err(err, errlen, "PositionalEncoding kernel not registered");
    return 3;
  }
  OpKernelConstruction cons;
  cons.int_attrs["encoding_size"] = encoding_size;
  OpKernel* k = it->second(&cons);
  if (!cons.status().ok()) {
    set_err(err, errlen, cons.status().error_message());
    delete k;
    return 1;
  }
  OpKernelContext ctx;
  ctx.inputs.push_back(Tensor(TensorShape(std::vector<int64>()), sizeof(int)));
  ctx.input_names.push_back("max_position");
  ctx.output_names.push_back("positional_encoding");
  *static_cast<int*>(ctx.inputs[0].raw()) = max_position;
  // pre-seed the output allocation with the caller's fill so untouched cells show
  k->Compute(&ctx);
  int rc = 0;
  if (!ctx.status().ok()) {
    set_err(err, errlen, ctx.status().error_message());
    rc = 2;
  } else {
    const float* src = static_cast<const float*>(ctx.outputs[0]->raw());
    const size_t n = static_cast<size_t>(max_position) * encoding_size;
    const int half = encoding_size / 2;
    for (size_t i = 0; i < n; ++i) {
      const int col = static_cast<int>(i % static_cast<size_t>(encoding_size));
      if (col < 2 * half) out[i] = src[i];   // columns the op wrote; the rest keep the caller's fill
    }
  }
  delete k;
  return rc;
}

// Number of ops / kernels the reference sources registered (sanity for tests).
int ref_registered(void) {
  return static_cast<int>(shim::KernelRegistry().size()) * 100 +
         static_cast<int>(shim::OpRegistry().size());
}

}  // extern "C"
