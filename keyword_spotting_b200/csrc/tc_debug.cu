// Hardware self-test of the tcgen05 conventions in tc05.cuh: one 128 x N x K product with the A operand
// either in TMEM (TS form) or in shared memory (SS form), B in shared memory (K-major, no swizzle), fp32
// accumulators read back with tcgen05.ld.  Exported as kws_debug_tc_gemm so the layout assumptions are pinned
// by a GPU test (tests/test_gpu_tc.py) independently of the recurrent kernel that builds on them.
#include "common.cuh"
#include "tc05.cuh"

namespace kws {

__global__ void __launch_bounds__(128, 1)
tc_gemm_test_kernel(const float* __restrict__ A, const __half* __restrict__ Bpacked, float* __restrict__ D, int N,
                    int K, int ss_mode, __half* __restrict__ Atile) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar_tx;
  __shared__ uint32_t tmem_base_smem;
  unsigned char* sB = smem;                                   // [N, K] fp16 canonical
  unsigned char* sA = smem + static_cast<size_t>(N) * K * 2;  // [128, K] fp16 canonical (SS form only)
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 512);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init(&bar_tx, 1);
    tc::mbar_fence_init();
  }
  const int b16 = N * K * 2 / 16;
  for (int i = tid; i < b16; i += blockDim.x)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bpacked)[i];
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t a_col = 256;
  const float* arow = A + static_cast<size_t>(tid) * K;
  if (ss_mode == 3) {
    // A takes the production route of the recurrent kernel: written to global memory in the row-tiled layout
    // (tc::tile_offset), then ONE bulk copy into shared memory, completion on an mbarrier
    for (int k8 = 0; k8 < K / 8; ++k8) {
      uint4 v;
      v.x = tc::pack_half2(arow[8 * k8 + 0], arow[8 * k8 + 1]);
      v.y = tc::pack_half2(arow[8 * k8 + 2], arow[8 * k8 + 3]);
      v.z = tc::pack_half2(arow[8 * k8 + 4], arow[8 * k8 + 5]);
      v.w = tc::pack_half2(arow[8 * k8 + 6], arow[8 * k8 + 7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(Atile) + tc::tile_offset(tid, 8 * k8)) = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(K / 8) * tc::kTileChunkBytes;
      tc::mbar_arrive_expect_tx(&bar_tx, bytes);
      tc::bulk_g2s(sA, Atile, bytes, &bar_tx);
    }
    tc::mbar_wait(&bar_tx, 0);
  } else if (ss_mode) {
    for (int k8 = 0; k8 < K / 8; ++k8) {
      uint4 v;
      v.x = tc::pack_half2(arow[8 * k8 + 0], arow[8 * k8 + 1]);
      v.y = tc::pack_half2(arow[8 * k8 + 2], arow[8 * k8 + 3]);
      v.z = tc::pack_half2(arow[8 * k8 + 4], arow[8 * k8 + 5]);
      v.w = tc::pack_half2(arow[8 * k8 + 6], arow[8 * k8 + 7]);
      *reinterpret_cast<uint4*>(sA + tc::canon_offset(tid, 8 * k8, K)) = v;
    }
  } else {
    for (int k = 0; k < K; k += 32) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = tc::pack_half2(arow[k + 2 * i], arow[k + 2 * i + 1]);
      tc::st16(tmem + lane_base + a_col + k / 2, v);
    }
    tc::wait_st();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = ss_mode == 2 ? tc::idesc_f16_bmn(128, N) : tc::idesc_f16(128, N);   // 2: B is MN-major
    const uint32_t sbo = static_cast<uint32_t>(K / 8) * 128;
    for (int k16 = 0; k16 < K / 16; ++k16) {
      const uint64_t bdesc = tc::smem_desc(tc::smem_u32(sB) + k16 * 256, 128, sbo);
      if (ss_mode == 3) {
        const uint64_t adesc = tc::smem_desc(tc::smem_u32(sA) + k16 * 2 * tc::kTileChunkBytes, tc::kTileChunkBytes, 128);
        tc::mma_ss(tmem, adesc, bdesc, idesc, k16 > 0);
      } else if (ss_mode) {
        const uint64_t adesc = tc::smem_desc(tc::smem_u32(sA) + k16 * 256, 128, sbo);
        tc::mma_ss(tmem, adesc, bdesc, idesc, k16 > 0);
      } else {
        tc::mma_ts(tmem, tmem + a_col + k16 * 8, bdesc, idesc, k16 > 0);
      }
    }
    tc::commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tc::ld32(tmem + lane_base + c, v);
    tc::wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c + i < N) D[static_cast<size_t>(tid) * N + c + i] = __uint_as_float(v[i]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace kws

// ss_mode: 0 = A in TMEM, 1 = A in shared memory, 2 = A in shared memory and B in the MN-major layout, 3 = A in the
// row-tiled layout in global memory, brought to shared memory by one bulk copy (the recurrent kernel's x operand).
// A [128, K] fp32 (rounded to fp16 on the device), B [N, K] fp32 (rounded and packed on the host side of
// this call) -> D [128, N] fp32 = A * B^T with fp32 accumulation.  N % 16 == 0, N <= 256, K % 32 == 0, K <= 256.
extern "C" int kws_debug_tc_gemm(const float* A, const float* B_host, float* D, int N, int K, int ss_mode,
                                 void* stream) {
  using namespace kws;
  clear_error();
  KWS_REQUIRE(A && B_host && D, "NULL pointer");
  KWS_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256, "N must be a multiple of 16 in [16, 256]");
  KWS_REQUIRE(K % 32 == 0 && K >= 32 && K <= 256, "K must be a multiple of 32 in [32, 256]");
  std::string packed(static_cast<size_t>(N) * K * 2, '\0');
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const __half h = __float2half_rn(B_host[static_cast<size_t>(n) * K + k]);
      *reinterpret_cast<__half*>(&packed[ss_mode == 2 ? tc::canon_offset_mn(n, k, K) : tc::canon_offset(n, k, K)]) = h;
    }
  __half* dB = nullptr;
  __half* dA = nullptr;
  KWS_CUDA_OK(cudaMalloc(&dB, packed.size()));
  if (ss_mode == 3) {
    const cudaError_t ea = cudaMalloc(&dA, static_cast<size_t>(128) * K * 2);
    if (ea != cudaSuccess) {
      cudaFree(dB);
      return fail(KWS_ERR_ALLOC, "tc_gemm_test_kernel: %s", cudaGetErrorString(ea));
    }
  }
  KWS_CUDA_OK(cudaMemcpy(dB, packed.data(), packed.size(), cudaMemcpyHostToDevice));
  const size_t smem = static_cast<size_t>(N) * K * 2 + static_cast<size_t>(128) * K * 2;
  KWS_CUDA_OK(cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  tc_gemm_test_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, dB, D, N, K, ss_mode, dA);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  cudaFree(dB);
  cudaFree(dA);
  if (e != cudaSuccess) return fail(KWS_ERR_CUDA, "tc_gemm_test_kernel failed: %s", cudaGetErrorString(e));
  return KWS_OK;
}
