// K5 on the 5th-generation tensor cores: OctbitMatMul with the activation quantiser fused in.
//
// Same arithmetic as octbit.cu (octbit/octbit_mat_mul_op.cc:90-181), for the model shapes (K <= 256, B <= 256):
//   out[r, n] = ( sum_k sat16-paired( q[r,k] * w[n,k] ) - bias[n] ) * scale,   q = u8 quantisation of x
// Per 128-row tile, warp-specialised and double-buffered:
//   warps 0-7  read x fp32 (each row once, 512-byte coalesced requests), quantise it exactly as the reference
//              (IEEE x / bscale, round half away, +127) and store the u8 tile straight into the tcgen05 K-major
//              canonical layout in shared memory -- the u8 copy of x never exists in HBM;
//   warp  8    issues tcgen05.mma kind::i8 (u8 x s8 -> s32, M128 x N x K32) against the weight matrix that stays
//              resident in shared memory, accumulators in TMEM (two buffers of 256 columns);
//   warps 9-16 read the exact int32 sums with tcgen05.ld, add the sparse saturation correction
//              sum(sat16(p) - p) over the only pairs that can overflow int16 (q taken from the same smem tile),
//              apply the reference's fp32 epilogue and write the rows.
// The integer sums are exact, so the result is bit-identical to the reference (K <= 512 keeps every fp32 lane sum
// an integer below 2^24, see octbit.cu).  HBM traffic per call: x once here (+ once in the min/max pass), out once.
#include "octbit_common.cuh"
#include "tc05.cuh"

namespace kws {

constexpr int kOtTile = 128;
constexpr int kOtProducerWarps = 8;
constexpr int kOtEpiWarps = 4;            // one per TMEM lane quarter (8 = two per quarter, each half of the columns)
constexpr int kOtThreads = 32 * (kOtProducerWarps + 1 + kOtEpiWarps);
constexpr int kOtLboA = 144;             // bytes between K-adjacent core matrices of the A tile (128 + 16: the
                                         // producers' 32-bit stores of a warp then hit 32 distinct banks)
constexpr int kOtLboW = 128;

struct OctbitTcParams {
  const float* x;
  const signed char* w;                  // [B, K] row-major
  const float* bias;
  float scale_attr;
  long A;
  int B, K, Npad;
  const OctbitHeader* hdr;
  const int* cand_count;                 // [B]
  const unsigned short* cand;            // [B, K/2]
  float* out;
};

// instruction descriptor for kind::i8: u8 A, s8 B (both K-major), s32 accumulate (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4)                                 // c_format = S32
         | (0u << 7) | (1u << 10)                  // a_format = unsigned 8 bit, b_format = signed 8 bit
         | (static_cast<uint32_t>(N >> 3) << 17)   // n_dim
         | (static_cast<uint32_t>(M >> 4) << 24);  // m_dim
}
__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}

__device__ __forceinline__ void ot_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// quant_one() without the IEEE division on the common path: t = x * (1/bscale) is within 2 ulp of x / bscale, so
// round-half-away of t equals that of the true quotient unless t + copysign(.5) lands within a few ulp of an integer;
// only those elements (~0.01 %) take the exact __fdiv_rn path.  Bit-identical to quant_one() by construction.
__device__ __forceinline__ unsigned quant_fast(float x, const QuantParams& p, float inv) {
  const float t = x * inv;
  const float u = t + copysignf(0.5f, t);
  const int n = __float2int_rz(u);
  const float d = fabsf(u - static_cast<float>(n));                     // in [0, 1): distance to the integer below |u|
  const float eps = fmaf(fabsf(t), 4.8e-7f, 1e-6f);                     // 4 ulp of t
  if (d < eps || d > 1.0f - eps || !(fabsf(t) < 1024.0f)) return quant_one(x, p);   // near tie, huge, inf/nan
  return static_cast<unsigned>(n + static_cast<int>(p.offset)) & 0xffu;
}

enum { kOtFull0 = 0, kOtFull1, kOtEmpty0, kOtEmpty1, kOtDone0, kOtDone1, kOtBars };

__global__ void __launch_bounds__(kOtThreads, 1)
octbit_tc_kernel(const OctbitTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int K = p.K, kc = K / 16;                              // 16-byte K chunks per row
  const uint32_t sboA = static_cast<uint32_t>(kc) * kOtLboA;    // bytes between 8-row groups
  const uint32_t sboW = static_cast<uint32_t>(kc) * kOtLboW;
  unsigned char* sW = smem;                                     // [Npad/8][kc][8][16]
  unsigned char* sA0 = sW + static_cast<size_t>(p.Npad / 8) * sboW;
  const size_t a_bytes = static_cast<size_t>(kOtTile / 8) * sboA;
  unsigned char* sA1 = sA0 + a_bytes;
  float* sBias = reinterpret_cast<float*>(sA1 + a_bytes);       // [Npad]
  int* sCnt = reinterpret_cast<int*>(sBias + p.Npad);           // [Npad]
  int* sAny = sCnt + p.Npad;                                    // [16] per 16-column chunk
  float* sStage = reinterpret_cast<float*>(sAny + 16);          // [8 warps][32][20] epilogue transpose
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kOtEpiWarps * 32 * 20);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kOtBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kMmaWarp = kOtProducerWarps;
  if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bars[kOtFull0], 32 * kOtProducerWarps);
    tc::mbar_init(&bars[kOtFull1], 32 * kOtProducerWarps);
    tc::mbar_init(&bars[kOtEmpty0], 32 * kOtEpiWarps);
    tc::mbar_init(&bars[kOtEmpty1], 32 * kOtEpiWarps);
    tc::mbar_init(&bars[kOtDone0], 1);
    tc::mbar_init(&bars[kOtDone1], 1);
    tc::mbar_fence_init();
  }
  // weights -> canonical layout (rows >= B are zero), bias and candidate counts
  for (int i = tid; i < p.Npad * kc; i += kOtThreads) {
    const int n = i / kc, c = i - n * kc;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < p.B) v = __ldg(reinterpret_cast<const uint4*>(p.w + static_cast<long>(n) * K + 16 * c));
    *reinterpret_cast<uint4*>(sW + (n >> 3) * sboW + c * kOtLboW + (n & 7) * 16) = v;
  }
  for (int i = tid; i < p.Npad; i += kOtThreads) {
    sBias[i] = i < p.B ? p.bias[i] : 0.0f;
    sCnt[i] = i < p.B ? p.cand_count[i] : 0;
  }
  if (tid < 16) {
    int a = 0;
    for (int i = 0; i < 16; ++i) {
      const int n = 16 * tid + i;
      if (n < p.B) a |= p.cand_count[n];
    }
    sAny[tid] = a;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const long ntiles = (p.A + kOtTile - 1) / kOtTile;
  const QuantParams qp = quant_params(p.hdr);

  if (warp < kOtProducerWarps) {
    const float inv = qp.bscale != 0.0f ? __frcp_rn(qp.bscale) : 0.0f;
    // ================================================= producers: x fp32 -> u8 canonical tile
    uint32_t it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&bars[kOtEmpty0 + buf], ((it >> 1) & 1) ^ 1);          // the epilogue has released this buffer
      unsigned char* sA = buf ? sA1 : sA0;
      constexpr int kRows = kOtTile / kOtProducerWarps;                 // rows of the tile per producer warp
      const long r0 = tile * kOtTile + warp * kRows;
      // lane l covers floats 4l..4l+3 of every 128-float segment of a row: chunk 8*seg + l/4, byte 4*(l%4)
      auto load_batch = [&](int rr, float4 (&v)[4][2]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long r = r0 + rr + u;
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            v[u][sg] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (128 * sg + 4 * lane < K && r < p.A)
              v[u][sg] = __ldg(reinterpret_cast<const float4*>(p.x + r * K + 128 * sg) + lane);
          }
        }
      };
      auto quant_batch = [&](int rr, const float4 (&v)[4][2]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = warp * kRows + rr + u;
          const long r = r0 + rr + u;
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            if (128 * sg + 4 * lane < K) {
              unsigned q = 0u;
              if (r < p.A)
                q = quant_fast(v[u][sg].x, qp, inv) | (quant_fast(v[u][sg].y, qp, inv) << 8) |
                    (quant_fast(v[u][sg].z, qp, inv) << 16) | (quant_fast(v[u][sg].w, qp, inv) << 24);
              *reinterpret_cast<unsigned*>(sA + (m >> 3) * sboA + (8 * sg + (lane >> 2)) * kOtLboA + (m & 7) * 16 +
                                           4 * (lane & 3)) = q;
            }
          }
        }
      };
      // two batches of 4 rows in flight: the loads of one batch overlap the quantisation of the other
      static_assert(kRows % 8 == 0, "producer pipeline works on pairs of 4-row batches");
      float4 va[4][2], vb[4][2];
      load_batch(0, va);
#pragma unroll 1
      for (int rr = 0; rr < kRows; rr += 8) {
        load_batch(rr + 4, vb);
        quant_batch(rr, va);
        if (rr + 8 < kRows) load_batch(rr + 8, va);
        quant_batch(rr + 4, vb);
      }
      tc::fence_proxy_async();                                            // generic-proxy stores -> MMA operand fetch
      ot_arrive(&bars[kOtFull0 + buf]);
    }
  } else if (warp == kMmaWarp) {
    // ================================================= MMA issuer
    const uint32_t idesc = idesc_i8(128, p.Npad);
    const uint64_t wdesc = tc::smem_desc(tc::smem_u32(sW), kOtLboW, sboW);
    const uint64_t adesc0 = tc::smem_desc(tc::smem_u32(sA0), kOtLboA, sboA);
    const uint64_t adesc1 = tc::smem_desc(tc::smem_u32(sA1), kOtLboA, sboA);
    uint32_t it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&bars[kOtFull0 + buf], (it >> 1) & 1);
      tc::fence_after_sync();
      if (lane == 0) {
        const uint64_t ad = buf ? adesc1 : adesc0;
#pragma unroll 1
        for (int k32 = 0; k32 < K / 32; ++k32)                            // two K-adjacent core matrices per MMA
          mma_i8_ss(tmem + 256 * buf, ad + ((2 * k32 * kOtLboA) >> 4), wdesc + ((2 * k32 * kOtLboW) >> 4), idesc, k32 > 0);
        tc::commit(&bars[kOtDone0 + buf]);
      }
      __syncwarp();
    }
  } else {
    // ================================================= epilogue: warp = (TMEM lane quarter, half of the columns)
    const int q4 = warp & 3;                                              // TMEM lane quarter this warp may read
    const int chalf = (warp - kMmaWarp - 1) >> 2;
    const bool split_cols = kOtEpiWarps == 8;
    const int row = 32 * q4 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(32 * q4) << 16;
    const float scale = __fmul_rn(p.scale_attr, qp.bscale);
    const int half = K / 2;
    const bool vec_ok = (p.B & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    const int nchunks = p.Npad / 16;
    const int c_begin = !split_cols || chalf == 0 ? 0 : (nchunks + 1) / 2;
    const int c_end = !split_cols ? nchunks : (chalf == 0 ? (nchunks + 1) / 2 : nchunks);
    float* stage = sStage + (warp - kMmaWarp - 1) * (32 * 20);            // [32 rows][16 cols], row stride 80 B
    uint32_t it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&bars[kOtDone0 + buf], (it >> 1) & 1);
      tc::fence_after_sync();
      const unsigned char* sA = buf ? sA1 : sA0;
      const unsigned char* qrow = sA + (row >> 3) * sboA + (row & 7) * 16;  // + (k/16)*LBO + k%16
      const long rbase = tile * kOtTile + 32 * q4;
      for (int ch = c_begin; ch < c_end; ++ch) {
        const int c0 = 16 * ch;
        uint32_t v[16];
        tc::ld16(tmem + lane_sel + 256 * buf + c0, v);
        tc::wait_ld();
        if (sAny[ch]) {                                                   // rare: columns with pairs that can saturate int16
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = c0 + i;
            const int cnt = sCnt[n];
            int delta = 0;
            for (int c = 0; c < cnt; ++c) {
              const int k = 2 * p.cand[static_cast<long>(n) * half + c];
              const signed char* wrow = p.w + static_cast<long>(n) * K;
              const int pq = static_cast<int>(qrow[(k >> 4) * kOtLboA + (k & 15)]) * wrow[k] +
                             static_cast<int>(qrow[(k >> 4) * kOtLboA + (k & 15) + 1]) * wrow[k + 1];
              delta += sat16(pq) - pq;
            }
            v[i] += static_cast<uint32_t>(delta);
          }
        }
        // the four fp32 lane sums of the reference are exact integers below 2^24 (K <= 512), so their lane-ordered
        // sum is float(total); then -bias (signed mode) and *scale as separate roundings (:172-179)
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + i);
          float o0 = static_cast<float>(static_cast<int>(v[i])), o1 = static_cast<float>(static_cast<int>(v[i + 1]));
          float o2 = static_cast<float>(static_cast<int>(v[i + 2])), o3 = static_cast<float>(static_cast<int>(v[i + 3]));
          if (qp.is_signed) {
            o0 = __fsub_rn(o0, b4.x); o1 = __fsub_rn(o1, b4.y); o2 = __fsub_rn(o2, b4.z); o3 = __fsub_rn(o3, b4.w);
          }
          *reinterpret_cast<float4*>(stage + lane * 20 + i) =
              make_float4(__fmul_rn(o0, scale), __fmul_rn(o1, scale), __fmul_rn(o2, scale), __fmul_rn(o3, scale));
        }
        __syncwarp();
        // write the 32 x 16 block: 4 lanes cover the 64 contiguous bytes of a row
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const int rr = itr * 8 + (lane >> 2), cc = 4 * (lane & 3);
          const float4 o4 = *reinterpret_cast<const float4*>(stage + rr * 20 + cc);
          const long gr = rbase + rr;
          const int n = c0 + cc;
          if (gr < p.A) {
            float* dst = p.out + gr * p.B + n;
            if (vec_ok && n + 4 <= p.B) {
              *reinterpret_cast<float4*>(dst) = o4;
            } else {
              if (n < p.B) dst[0] = o4.x;
              if (n + 1 < p.B) dst[1] = o4.y;
              if (n + 2 < p.B) dst[2] = o4.z;
              if (n + 3 < p.B) dst[3] = o4.w;
            }
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();                                            // TMEM reads done before the buffer is reused
      ot_arrive(&bars[kOtEmpty0 + buf]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc::tmem_dealloc(tmem, 512);
}

static size_t octbit_tc_smem(int Npad, int K) {
  const size_t kc = K / 16;
  return static_cast<size_t>(Npad / 8) * kc * kOtLboW + 2 * static_cast<size_t>(kOtTile / 8) * kc * kOtLboA +
         sizeof(float) * Npad + sizeof(int) * Npad + sizeof(int) * 16 + sizeof(float) * kOtEpiWarps * 32 * 20 +
         sizeof(uint64_t) * kOtBars + 16;
}

bool octbit_tc_supported(int64_t A, int64_t B, int64_t K) {
  // narrow outputs (the 128 -> 6 FC) stay on the mma.sync path: with B < 64 this kernel is bound by its
  // quantising producers and the separate quantise + GEMM is faster (measured 1.6 vs 2.0 ms at 3.9 M rows)
  return A >= 1 && B >= 64 && B <= 256 && K >= 64 && K <= 256 && K % 64 == 0;
}

int launch_octbit_tc(const float* x, const int8_t* w, const float* bias, float scale_attr, int64_t A, int64_t B, int64_t K,
                     const OctbitHeader* hdr, const int* cand_count, const unsigned short* cand, float* out,
                     cudaStream_t st) {
  OctbitTcParams p;
  p.x = x;
  p.w = reinterpret_cast<const signed char*>(w);
  p.bias = bias;
  p.scale_attr = scale_attr;
  p.A = A;
  p.B = static_cast<int>(B);
  p.K = static_cast<int>(K);
  p.Npad = static_cast<int>((B + 15) / 16 * 16);
  p.hdr = hdr;
  p.cand_count = cand_count;
  p.cand = cand;
  p.out = out;
  const size_t smem = octbit_tc_smem(p.Npad, p.K);
  KWS_CUDA_OK(cudaFuncSetAttribute(octbit_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long ntiles = ceil_div(A, kOtTile);
  const long blocks = ntiles < sm_count() ? ntiles : sm_count();
  octbit_tc_kernel<<<static_cast<unsigned>(blocks), kOtThreads, smem, st>>>(p);
  KWS_LAUNCH_OK("octbit_tc_kernel");
  return KWS_OK;
}

}  // namespace kws
