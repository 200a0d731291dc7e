// K6 -- positional encoding table fill.
// Reference: PositionalEncodingOp::Compute, positional_encoding/positional_encoding_op.cc:32-50
//   pe[p, 2i]   = sin(p / pow(10000.0, 2.0*i/size))      (double math, float store)
//   pe[p, 2i+1] = cos(same)                              i < size/2
// For odd `size` the last column is never written (:45), so it is left alone here.
//
// The size/2 frequency denominators are plan-time constants: they are produced
// once per (device, size) with the host libm `pow` -- the very function the
// reference calls -- and cached on the device, so that the table can only differ
// from the reference through the last-bit behaviour of double sin/cos.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace kws {

__global__ void __launch_bounds__(256)
posenc_kernel(const double* __restrict__ denom, int max_position, int size, int half,
              float* __restrict__ out) {
  const long total = static_cast<long>(max_position) * half;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx / half);
    const int i = static_cast<int>(idx - static_cast<long>(p) * half);
    const double angle = static_cast<double>(p) / denom[i];
    double s, c;
    sincos(angle, &s, &c);
    float* dst = out + static_cast<long>(p) * size + 2 * i;
    if ((size & 1) == 0) {
      *reinterpret_cast<float2*>(dst) = make_float2(static_cast<float>(s), static_cast<float>(c));
    } else {
      dst[0] = static_cast<float>(s);
      dst[1] = static_cast<float>(c);
    }
  }
}

static std::mutex g_pe_mu;
static std::map<std::pair<int, int>, double*> g_pe_denoms;

static int get_denoms(int size, double** out) {
  int dev = 0;
  KWS_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_pe_mu);
  auto key = std::make_pair(dev, size);
  auto it = g_pe_denoms.find(key);
  if (it != g_pe_denoms.end()) {
    *out = it->second;
    return KWS_OK;
  }
  const int half = size / 2;
  std::vector<double> host(half > 0 ? half : 1);
  for (int i = 0; i < half; ++i) host[i] = std::pow(10000.0, 2.0 * i / size);  // op.cc:46
  double* dptr = nullptr;
  KWS_CUDA_OK(cudaMalloc(&dptr, sizeof(double) * host.size()));
  KWS_CUDA_OK(cudaMemcpy(dptr, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
  g_pe_denoms[key] = dptr;
  *out = dptr;
  return KWS_OK;
}

}  // namespace kws

extern "C" int kws_positional_encoding(int32_t max_position, int32_t encoding_size, float* out,
                                       void* stream) {
  using namespace kws;
  clear_error();
  KWS_REQUIRE(encoding_size >= 1, "encoding_size must be >= 1 (Attr \"encoding_size: int >= 1\")");
  KWS_REQUIRE(max_position >= 0, "max_position must be >= 0");
  const int half = encoding_size / 2;
  if (max_position == 0 || half == 0) return KWS_OK;
  KWS_REQUIRE(out != nullptr, "out is NULL");
  double* denom = nullptr;
  int rc = get_denoms(encoding_size, &denom);
  if (rc != KWS_OK) return rc;
  const long total = static_cast<long>(max_position) * half;
  const int threads = 256;
  long blocks = ceil_div(total, threads);
  const long cap = static_cast<long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  posenc_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      denom, max_position, encoding_size, half, out);
  KWS_LAUNCH_OK("posenc_kernel");
  return KWS_OK;
}
