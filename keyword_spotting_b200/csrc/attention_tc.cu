// Dense layers of the attention_ctc forward pass on the 5th-generation tensor cores (models/attention_ctc.py:61-70
// feed_forward, :28-58 the q/k/v projections, :92-98 input_linear_trans).
//
//   Y[r, n] = act( sum_k A[r, k] W[k, n] + bias[n] ) (+ pe[r % Tp, n])          A fp32 [rows, K], Y fp32 [rows, N]
//
// One CTA = a [128 rows, 128 columns] output tile, three CTAs per SM.  K is walked in chunks of 64: the CTA's 256 threads
// read the A chunk from HBM (coalesced 16-byte loads), split every value into two fp16 terms (x = hi + lo, exact to
// 2^-22) and store both in the tcgen05 K-major canonical layout in shared memory; the weight chunk arrives pre-split and
// pre-packed (hi | lo, packed once at model creation) as one linear 32 KB copy.  One elected thread issues
// A_hi W_hi + A_lo W_hi + A_hi W_lo  (12 tcgen05.mma kind::f16 per chunk, SS form, fp32 accumulation in 128 TMEM
// columns): fp32-grade results (the dropped lo x lo term is 2^-22 relative), which is what the 1e-4 parity of this model's
// tests needs -- a single fp16 rounding of the layer-normed activations would cost 5e-4.  Epilogue: TMEM -> registers,
// bias / relu, through a shared-memory transpose to 64-byte row pieces (+ the positional encoding), fp32 to HBM.
// The chunk loop is not pipelined inside a CTA; the other two CTAs of the SM fill its load and store phases.
#include <vector>

#include "common.cuh"
#include "tc05.cuh"

namespace kws {

constexpr int kAlTile = 128;                     // rows and columns of an output tile
constexpr int kAlKc = 64;                        // K elements per chunk
constexpr int kAlThreads = 256;
constexpr int kAlLboA = 144;                     // A operand: 128 + 16 bytes between K-adjacent core matrices (bank spread of the 8-byte stores)
constexpr int kAlSboA = (kAlKc / 8) * kAlLboA;   // 1152
constexpr int kAlPartA = (kAlTile / 8) * kAlSboA;   // 18432 bytes per fp16 part of the A chunk
constexpr int kAlSboW = (kAlKc / 8) * 128;       // 1024: the weights are packed densely (linear copy)
constexpr int kAlPartW = (kAlTile / 8) * kAlSboW;   // 16384
constexpr int kAlSmem = 2 * kAlPartA + 2 * kAlPartW;

template <bool kRelu, bool kAddPe>
__global__ void __launch_bounds__(kAlThreads, 3)
att_linear_tc_kernel(const float* __restrict__ A, const unsigned char* __restrict__ wpack, const float* __restrict__ bias,
                     const float* __restrict__ pe, int Tp, long rows, int K, int nchunks, int N, float* __restrict__ Y) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                               // [hi | lo] x kAlPartA
  unsigned char* sW = smem + 2 * kAlPartA;                // [hi | lo] x kAlPartW
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long r0 = blockIdx.x * static_cast<long>(kAlTile);
  const int nt = blockIdx.y;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::idesc_f16(128, 128);
  const uint64_t adesc = tc::smem_desc(tc::smem_u32(sA), kAlLboA, kAlSboA);
  const uint64_t wdesc = tc::smem_desc(tc::smem_u32(sW), 128, kAlSboW);

  for (int kc = 0; kc < nchunks; ++kc) {
    if (kc > 0) {                                          // the previous chunk's MMAs have read both buffers
      tc::mbar_wait(&bar, (kc - 1) & 1);
      tc::fence_after_sync();
    }
    {
      const uint4* wsrc = reinterpret_cast<const uint4*>(wpack + (static_cast<size_t>(nt) * nchunks + kc) * (2 * kAlPartW));
#pragma unroll
      for (int i = tid; i < 2 * kAlPartW / 16; i += kAlThreads) reinterpret_cast<uint4*>(sW)[i] = __ldg(wsrc + i);
    }
#pragma unroll
    for (int i = tid; i < kAlTile * (kAlKc / 4); i += kAlThreads) {
      const int row = i >> 4, q = i & 15;                  // 16 consecutive threads read the 256 contiguous bytes of a row
      const int k = kc * kAlKc + 4 * q;
      const long gr = r0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows) {
        const float* src = A + gr * K + k;
        if (k + 3 < K) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (k < K) v.x = __ldg(src);
          if (k + 1 < K) v.y = __ldg(src + 1);
          if (k + 2 < K) v.z = __ldg(src + 2);
        }
      }
      const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
      const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
      const uint32_t off = (row >> 3) * kAlSboA + (q >> 1) * kAlLboA + (row & 7) * 16 + (q & 1) * 8;
      *reinterpret_cast<uint2*>(sA + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint2*>(sA + kAlPartA + off) =
          make_uint2(tc::pack_half2(v.x - b01.x, v.y - b01.y), tc::pack_half2(v.z - b23.x, v.w - b23.y));
    }
    tc::fence_proxy_async();                               // generic-proxy stores -> MMA operand fetch
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int k16 = 0; k16 < kAlKc / 16; ++k16) {
          const uint64_t a_hi = adesc + ((2 * k16 * kAlLboA) >> 4), a_lo = a_hi + (kAlPartA >> 4);
          const uint64_t w_hi = wdesc + ((2 * k16 * 128) >> 4), w_lo = w_hi + (kAlPartW >> 4);
          tc::mma_ss(tmem, a_hi, w_hi, idesc, kc > 0 || k16 > 0);
          tc::mma_ss(tmem, a_lo, w_hi, idesc, true);
          tc::mma_ss(tmem, a_hi, w_lo, idesc, true);
        }
        tc::commit(&bar);
      }
    }
  }
  tc::mbar_wait(&bar, (nchunks - 1) & 1);
  tc::fence_after_sync();

  // ---- epilogue: warp = (TMEM lane quarter, half of the tile's columns); the A buffers are free and serve as staging
  const int q4 = warp & 3, chalf = warp >> 2;
  const uint32_t lane_sel = static_cast<uint32_t>(32 * q4) << 16;
  float* stage = reinterpret_cast<float*>(sA) + warp * (32 * 20);          // [32 rows][16 cols], row stride 80 bytes
  const long rbase = r0 + 32 * q4;
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    const int c0 = 64 * chalf + 16 * ch;                                   // column of the tile
    const int n0 = nt * kAlTile + c0;
    uint32_t v[16];
    tc::ld16(tmem + lane_sel + c0, v);
    tc::wait_ld();
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + i));
      float o0 = __uint_as_float(v[i]) + b4.x, o1 = __uint_as_float(v[i + 1]) + b4.y;
      float o2 = __uint_as_float(v[i + 2]) + b4.z, o3 = __uint_as_float(v[i + 3]) + b4.w;
      if (kRelu) {
        o0 = fmaxf(o0, 0.0f); o1 = fmaxf(o1, 0.0f); o2 = fmaxf(o2, 0.0f); o3 = fmaxf(o3, 0.0f);
      }
      *reinterpret_cast<float4*>(stage + lane * 20 + i) = make_float4(o0, o1, o2, o3);
    }
    __syncwarp();
#pragma unroll
    for (int itr = 0; itr < 4; ++itr) {                                    // 4 lanes cover the 64 contiguous bytes of a row
      const int rr = itr * 8 + (lane >> 2), cc = 4 * (lane & 3);
      float4 o4 = *reinterpret_cast<const float4*>(stage + rr * 20 + cc);
      const long grr = rbase + rr;
      if (grr < rows) {
        if (kAddPe) {
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(pe + static_cast<long>(grr % Tp) * N + n0 + cc));
          o4.x += p4.x; o4.y += p4.y; o4.z += p4.z; o4.w += p4.w;
        }
        *reinterpret_cast<float4*>(Y + grr * N + n0 + cc) = o4;
      }
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

bool att_linear_tc_supported(int K, int N) { return N % kAlTile == 0 && K % 4 == 0 && K >= 4; }

// W [K, N] row-major fp32 (host) -> per (column tile, K chunk): [hi | lo] fp16 in the canonical K-major layout of the
// B operand ([128 rows = output columns, 64 k]), zero beyond K.  Returns the device buffer.
int att_pack_linear(const float* W, int K, int N, unsigned char** out, int* nchunks_out) {
  const int nchunks = (K + kAlKc - 1) / kAlKc;
  const int ntiles = N / kAlTile;
  std::vector<unsigned char> host(static_cast<size_t>(ntiles) * nchunks * 2 * kAlPartW, 0);
  for (int nt = 0; nt < ntiles; ++nt)
    for (int kc = 0; kc < nchunks; ++kc) {
      unsigned char* base = host.data() + (static_cast<size_t>(nt) * nchunks + kc) * (2 * kAlPartW);
      for (int n = 0; n < kAlTile; ++n)
        for (int kk = 0; kk < kAlKc; ++kk) {
          const int k = kc * kAlKc + kk;
          if (k >= K) continue;
          const float w = W[static_cast<size_t>(k) * N + nt * kAlTile + n];
          const __half hi = __float2half_rn(w);
          const __half lo = __float2half_rn(w - __half2float(hi));
          const size_t off = static_cast<size_t>(n / 8) * kAlSboW + static_cast<size_t>(kk / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2;
          *reinterpret_cast<__half*>(base + off) = hi;
          *reinterpret_cast<__half*>(base + kAlPartW + off) = lo;
        }
    }
  *out = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(out), host.size());
  if (e != cudaSuccess) return fail(KWS_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", host.size(), cudaGetErrorString(e));
  KWS_CUDA_OK(cudaMemcpy(*out, host.data(), host.size(), cudaMemcpyHostToDevice));
  *nchunks_out = nchunks;
  return KWS_OK;
}

int launch_att_linear_tc(const float* A, const unsigned char* wpack, int nchunks, const float* bias, const float* pe, int Tp,
                         long rows, int K, int N, bool relu, bool add_pe, float* Y, cudaStream_t st) {
  if (rows <= 0) return KWS_OK;
  dim3 grid(static_cast<unsigned>(ceil_div(rows, kAlTile)), static_cast<unsigned>(N / kAlTile));
  auto launch = [&](auto kernel) -> int {
    KWS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAlSmem));
    kernel<<<grid, kAlThreads, kAlSmem, st>>>(A, wpack, bias, pe, Tp, rows, K, nchunks, N, Y);
    return KWS_OK;
  };
  int rc;
  if (relu) rc = add_pe ? launch(att_linear_tc_kernel<true, true>) : launch(att_linear_tc_kernel<true, false>);
  else rc = add_pe ? launch(att_linear_tc_kernel<false, true>) : launch(att_linear_tc_kernel<false, false>);
  if (rc != KWS_OK) return rc;
  KWS_LAUNCH_OK("att_linear_tc_kernel");
  return KWS_OK;
}

}  // namespace kws
