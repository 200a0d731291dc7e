// K5 on the 5th-generation tensor cores: OctbitMatMul with the activation quantiser fused in.
//
// Same arithmetic as octbit.cu (octbit/octbit_mat_mul_op.cc:90-181), for the model shapes (K <= 256, B <= 256):
//   out[r, n] = ( sum_k sat16-paired( q[r,k] * w[n,k] ) - bias[n] ) * scale,   q = u8 quantisation of x
// Per 128-row tile, warp-specialised and double-buffered:
//   warp  9    (one lane) streams x fp32 from HBM with TMA: cp.async.bulk.tensor.2d boxes of [16 rows, K] into a ring
//              of three shared-memory slabs (48 KB in flight per SM, rows past A zero-filled by the tensor map);
//   warps 0-7  quantise the slabs exactly as the reference (IEEE x / bscale, round half away, +127: a reciprocal
//              multiply + round-to-nearest with an exact-division fallback for the ~0.01 % of values near a tie) and
//              store the u8 tile straight into the tcgen05 K-major canonical layout -- the u8 copy of x never exists in HBM;
//   warp  8    issues tcgen05.mma kind::i8 (u8 x s8 -> s32, M128 x N x K32) against the weight matrix that stays
//              resident in shared memory, accumulators in TMEM (two buffers of 256 columns);
//   warps 10-17 read the exact int32 sums with tcgen05.ld, add the sparse saturation correction
//              sum(sat16(p) - p) over the only pairs that can overflow int16 (q taken from the same smem tile),
//              apply the reference's fp32 epilogue and write 64 contiguous bytes (two full sectors) per row and load.
// The integer sums are exact, so the result is bit-identical to the reference (K <= 512 keeps every fp32 lane sum
// an integer below 2^24, see octbit.cu).  HBM traffic per call: x once here (+ once in the min/max pass), out once.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "octbit_common.cuh"
#include "tc05.cuh"

namespace kws {

constexpr int kOtTile = 128;
constexpr int kOtProducerWarps = 8;       // quantisers: shared-memory fp32 slab -> u8 canonical tile
constexpr int kOtMmaWarp = 8;
constexpr int kOtTmaWarp = 9;             // one elected lane issues cp.async.bulk.tensor (UTMALDG) into the fp32 ring
constexpr int kOtEpiWarp0 = 10;
constexpr int kOtEpiWarps = 8;            // two per TMEM lane quarter, each half of the columns
constexpr int kOtThreads = 32 * (kOtEpiWarp0 + kOtEpiWarps);
constexpr int kOtChunkRows = 16;          // rows of x per TMA box: [16, K] fp32 = 16 KB at K = 256
constexpr int kOtStages = 3;              // ring of TMA boxes: 48 KB in flight per SM (bandwidth x latency of HBM ~ 44 KB)
constexpr int kOtLboA = 144;             // bytes between K-adjacent core matrices of the A tile (128 + 16: the
                                         // producers' 32-bit stores of a warp then hit 32 distinct banks)
constexpr int kOtLboW = 128;
constexpr int kOtCandCap = 2048;         // saturation candidates kept in shared memory (pair index, w0, w1 packed); more -> global lists

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

// Four activations quantised as quant_one() does, without the IEEE division on the common path: t = x * (1/bscale) is
// within 2 ulp (3e-5 for |t| < 256) of the true fp32 quotient Q = x / bscale, so round-to-nearest of t equals
// round-half-away of Q whenever |t - n| <= 0.5 - 1.2e-4; everything else (near ties, huge values, inf / nan: ~0.01 %)
// takes the exact __fdiv_rn path.  Bit-identical to quant_one() by construction.
__device__ __forceinline__ unsigned quant4_fast(float4 v, const QuantParams& p, float inv, int off) {
  const float t0 = v.x * inv, t1 = v.y * inv, t2 = v.z * inv, t3 = v.w * inv;
  const int n0 = __float2int_rn(t0), n1 = __float2int_rn(t1), n2 = __float2int_rn(t2), n3 = __float2int_rn(t3);
  const float d0 = fabsf(t0 - static_cast<float>(n0)), d1 = fabsf(t1 - static_cast<float>(n1));
  const float d2 = fabsf(t2 - static_cast<float>(n2)), d3 = fabsf(t3 - static_cast<float>(n3));
  constexpr float kSafe = 0.49988f;
  const bool safe = (d0 <= kSafe) & (d1 <= kSafe) & (d2 <= kSafe) & (d3 <= kSafe);          // NaN compares false
  if (!safe) return quant_one(v.x, p) | (quant_one(v.y, p) << 8) | (quant_one(v.z, p) << 16) | (quant_one(v.w, p) << 24);
  return (static_cast<unsigned>(n0 + off) & 0xffu) | ((static_cast<unsigned>(n1 + off) & 0xffu) << 8) |
         ((static_cast<unsigned>(n2 + off) & 0xffu) << 16) | (static_cast<unsigned>(n3 + off) << 24);
}

enum { kOtRingFull0 = 0, kOtRingEmpty0 = kOtStages, kOtFull0 = 2 * kOtStages, kOtFull1, kOtEmpty0, kOtEmpty1, kOtDone0, kOtDone1, kOtBars };

__global__ void __launch_bounds__(kOtThreads, 1)
octbit_tc_kernel(const __grid_constant__ CUtensorMap tmap, const OctbitTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int K = p.K, kc = K / 16;                              // 16-byte K chunks per row
  const uint32_t sboA = static_cast<uint32_t>(kc) * kOtLboA;    // bytes between 8-row groups
  const uint32_t sboW = static_cast<uint32_t>(kc) * kOtLboW;
  const uint32_t slab_bytes = static_cast<uint32_t>(kOtChunkRows) * K * 4;
  unsigned char* sRing = smem;                                  // [kOtStages][16 rows][K] fp32 (TMA destination, 128-byte aligned)
  unsigned char* sW = sRing + kOtStages * slab_bytes;           // [Npad/8][kc][8][16]
  unsigned char* sA0 = sW + static_cast<size_t>(p.Npad / 8) * sboW;
  const size_t a_bytes = static_cast<size_t>(kOtTile / 8) * sboA;
  unsigned char* sA1 = sA0 + a_bytes;
  float* sBias = reinterpret_cast<float*>(sA1 + a_bytes);       // [Npad]
  int* sCnt = reinterpret_cast<int*>(sBias + p.Npad);           // [Npad]
  int* sAny = sCnt + p.Npad;                                    // [16] per 16-column chunk
  int* sOff = sAny + 16;                                        // [Npad + 1] start of each column's candidates in sCand
  int* sRowMax = sOff + p.Npad + 2;                             // [2 bufs][128] largest u8 code of each row of the tile; then [1] Wmax
  float* sStage = reinterpret_cast<float*>(sRowMax + 2 * kOtTile + 2);    // [8 warps][32][20] epilogue transpose
  unsigned* sCand = reinterpret_cast<unsigned*>(sStage + kOtEpiWarps * 32 * 20);   // [kOtCandCap] pair index | w0 << 8 | w1 << 16 (+2: keeps the mbarriers 8-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCand + kOtCandCap);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kOtBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kOtMmaWarp) tc::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < kOtStages; ++s) {
      tc::mbar_init(&bars[kOtRingFull0 + s], 1);                          // the TMA lane's arrive.expect_tx + the bytes
      tc::mbar_init(&bars[kOtRingEmpty0 + s], 32 * kOtProducerWarps);
    }
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
  if (tid < 16) sAny[tid] = 0;              // per 16-column chunk: largest |w0| + |w1| of its candidate pairs (0 = none)
  __syncthreads();
  // the saturation candidates of every column, with their two weights, packed into shared memory: the epilogue's
  // correction loop then never touches global memory (a dependent LDG chain per candidate made it the slowest role)
  if (tid == 0) {
    int acc = 0;
    for (int n = 0; n < p.Npad; ++n) {
      sOff[n] = acc;
      acc += sCnt[n];
    }
    sOff[p.Npad] = acc;
  }
  __syncthreads();
  // A pair can only saturate int16 if q_max(row) * (|w0| + |w1|) >= 32768: a row skips the correction of a 16-column
  // chunk whenever its largest code times the chunk's heaviest pair stays below that (all but the few rows holding
  // values near the tensor's extreme, or the chunk holding the weight of largest magnitude).
  for (int n = tid; n < p.B; n += kOtThreads) {
    const int cnt = sCnt[n];
    const signed char* wrow = p.w + static_cast<long>(n) * K;
    int wm = 0;
    for (int c = 0; c < cnt; ++c) {
      const int kp = p.cand[static_cast<long>(n) * (K / 2) + c];
      wm = max(wm, abs(static_cast<int>(wrow[2 * kp])) + abs(static_cast<int>(wrow[2 * kp + 1])));
    }
    if (wm) atomicMax(&sAny[n >> 4], wm);
  }
  __syncthreads();
  const bool cand_in_smem = sOff[p.Npad] <= kOtCandCap;
  if (cand_in_smem) {
    for (int n = tid; n < p.B; n += kOtThreads) {
      const int cnt = sCnt[n], o = sOff[n];
      const signed char* wrow = p.w + static_cast<long>(n) * K;
      for (int c = 0; c < cnt; ++c) {
        const unsigned kp = p.cand[static_cast<long>(n) * (K / 2) + c];
        sCand[o + c] = kp | ((static_cast<unsigned>(wrow[2 * kp]) & 0xffu) << 8) | ((static_cast<unsigned>(wrow[2 * kp + 1]) & 0xffu) << 16);
      }
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const long ntiles = (p.A + kOtTile - 1) / kOtTile;
  const QuantParams qp = quant_params(p.hdr);
  constexpr int kChunks = kOtTile / kOtChunkRows;                         // 8 TMA boxes per tile

  if (warp == kOtTmaWarp) {
    // ================================================= TMA: x [A, K] fp32 -> ring of [16, K] slabs
    if (tc::elect_one()) {
      uint32_t n = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int c = 0; c < kChunks; ++c, ++n) {
          const uint32_t stage = n % kOtStages;
          if (n >= kOtStages) tc::mbar_wait(&bars[kOtRingEmpty0 + stage], ((n / kOtStages) & 1) ^ 1);
          const uint32_t bar = tc::smem_u32(&bars[kOtRingFull0 + stage]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slab_bytes) : "memory");
          const int row0 = static_cast<int>(tile * kOtTile + c * kOtChunkRows);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
              ::"r"(tc::smem_u32(sRing + stage * slab_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(row0), "r"(bar)
              : "memory");
        }
    }
  } else if (warp < kOtProducerWarps) {
    // ================================================= quantisers: fp32 slab -> u8 canonical tile
    const float inv = qp.bscale != 0.0f ? __frcp_rn(qp.bscale) : 0.0f;
    const int off = qp.bscale != 0.0f ? static_cast<int>(qp.offset) : 0;     // bscale == 0 (x all zero): q = 0, as quant_one()
    uint32_t it = 0, n = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&bars[kOtEmpty0 + buf], ((it >> 1) & 1) ^ 1);          // the epilogue has released this buffer
      unsigned char* sA = buf ? sA1 : sA0;
      for (int c = 0; c < kChunks; ++c, ++n) {
        const uint32_t stage = n % kOtStages;
        tc::mbar_wait(&bars[kOtRingFull0 + stage], (n / kOtStages) & 1);
        const float* slab = reinterpret_cast<const float*>(sRing + stage * slab_bytes);
        // warp w quantises rows 2w, 2w+1 of the slab; lane l covers floats 4l..4l+3 of every 128-float segment:
        // chunk 8*seg + l/4, byte 4*(l%4) of the row in the canonical tile
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int rs = 2 * warp + u;                                     // row of the slab
          const int m = c * kOtChunkRows + rs;                             // row of the tile
          unsigned qmax4 = 0u;
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            if (128 * sg + 4 * lane < K) {
              const float4 v = *reinterpret_cast<const float4*>(slab + rs * K + 128 * sg + 4 * lane);
              const unsigned q = quant4_fast(v, qp, inv, off);
              qmax4 = __vmaxu4(qmax4, q);
              *reinterpret_cast<unsigned*>(sA + (m >> 3) * sboA + (8 * sg + (lane >> 2)) * kOtLboA + (m & 7) * 16 +
                                           4 * (lane & 3)) = q;
            }
          }
          const unsigned m2 = max(max(qmax4 & 0xffu, (qmax4 >> 8) & 0xffu), max((qmax4 >> 16) & 0xffu, qmax4 >> 24));
          const unsigned rmax = __reduce_max_sync(0xffffffffu, m2);
          if (lane == 0) sRowMax[buf * kOtTile + m] = static_cast<int>(rmax);
        }
        ot_arrive(&bars[kOtRingEmpty0 + stage]);                            // this thread's reads of the slab are done
      }
      tc::fence_proxy_async();                                            // generic-proxy stores -> MMA operand fetch
      ot_arrive(&bars[kOtFull0 + buf]);
    }
  } else if (warp == kOtMmaWarp) {
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
      if (tc::elect_one()) {                                              // one elect.sync region: straight UTCIMMA issue
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
    const int chalf = (warp - kOtEpiWarp0) >> 2;
    const int row = 32 * q4 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(32 * q4) << 16;
    const float scale = __fmul_rn(p.scale_attr, qp.bscale);
    const int half = K / 2;
    const bool vec_ok = (p.B & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    const int nchunks = p.Npad / 16;
    const int c_begin = chalf == 0 ? 0 : (nchunks + 1) / 2;
    const int c_end = chalf == 0 ? (nchunks + 1) / 2 : nchunks;
    uint32_t it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&bars[kOtDone0 + buf], (it >> 1) & 1);
      tc::fence_after_sync();
      const unsigned char* sA = buf ? sA1 : sA0;
      const unsigned char* qrow = sA + (row >> 3) * sboA + (row & 7) * 16;  // + (k/16)*LBO + k%16
      const int row_qmax = sRowMax[buf * kOtTile + row];
      float* stage = sStage + (warp - kOtEpiWarp0) * (32 * 20);            // [32 rows][16 cols], row stride 80 B
      const long rbase = tile * kOtTile + 32 * q4;
      for (int ch = c_begin; ch < c_end; ++ch) {
        const int c0 = 16 * ch;
        uint32_t v[16];
        tc::ld16(tmem + lane_sel + 256 * buf + c0, v);
        tc::wait_ld();
        if (row_qmax * sAny[ch] >= 32768) {                               // rare: a near-extreme code x a chunk with heavy pairs
          // kept OUT of the unrolled code (one compact loop, its 16 corrections parked in this lane's row of the staging
          // tile): unrolled sixteen times it set the register allocation of the whole role and spilled the store
          // loop's addresses -- a spill reload is an L2 round trip next to 200 KB of shared memory
          int* dpark = reinterpret_cast<int*>(stage + lane * 20);
#pragma unroll 1
          for (int i = 0; i < 16; ++i) {
            const int n = c0 + i;
            const int cnt = sCnt[n];
            int delta = 0;
            if (cand_in_smem) {
              const unsigned* e = sCand + sOff[n];
              for (int c = 0; c < cnt; ++c) {
                const unsigned pk = e[c];
                const int k = 2 * static_cast<int>(pk & 0xffu);
                const int w0 = static_cast<signed char>((pk >> 8) & 0xffu), w1 = static_cast<signed char>((pk >> 16) & 0xffu);
                const int pq = static_cast<int>(qrow[(k >> 4) * kOtLboA + (k & 15)]) * w0 +
                               static_cast<int>(qrow[(k >> 4) * kOtLboA + (k & 15) + 1]) * w1;
                delta += sat16(pq) - pq;
              }
            } else {
              for (int c = 0; c < cnt; ++c) {
                const int k = 2 * p.cand[static_cast<long>(n) * half + c];
                const signed char* wrow = p.w + static_cast<long>(n) * K;
                const int pq = static_cast<int>(qrow[(k >> 4) * kOtLboA + (k & 15)]) * wrow[k] +
                               static_cast<int>(qrow[(k >> 4) * kOtLboA + (k & 15) + 1]) * wrow[k + 1];
                delta += sat16(pq) - pq;
              }
            }
            dpark[i] = delta;
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int4 d4 = *reinterpret_cast<const int4*>(dpark + i);
            v[i] += static_cast<uint32_t>(d4.x);
            v[i + 1] += static_cast<uint32_t>(d4.y);
            v[i + 2] += static_cast<uint32_t>(d4.z);
            v[i + 3] += static_cast<uint32_t>(d4.w);
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
        // write the 32 x 16 block: 4 lanes cover the 64 contiguous bytes of a row (scattered 16-byte stores, one row per
        // lane, were measured 4x slower: 32 cache lines per store instruction)
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const int rr = itr * 8 + (lane >> 2), cc = 4 * (lane & 3);
          const float4 o4 = *reinterpret_cast<const float4*>(stage + rr * 20 + cc);
          const long grr = rbase + rr;
          const int n = c0 + cc;
          if (grr < p.A) {
            float* dst = p.out + grr * p.B + n;
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
  if (warp == kOtMmaWarp) tc::tmem_dealloc(tmem, 512);
}

static size_t octbit_tc_smem(int Npad, int K) {
  const size_t kc = K / 16;
  return static_cast<size_t>(kOtStages) * kOtChunkRows * K * 4 + static_cast<size_t>(Npad / 8) * kc * kOtLboW +
         2 * static_cast<size_t>(kOtTile / 8) * kc * kOtLboA + sizeof(float) * Npad + sizeof(int) * Npad + sizeof(int) * 16 +
         sizeof(int) * (Npad + 2) + sizeof(int) * (2 * kOtTile + 2) + sizeof(float) * kOtEpiWarps * 32 * 20 + sizeof(unsigned) * kOtCandCap + sizeof(uint64_t) * kOtBars + 16;
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
  // x [A, K] fp32 as a 2-D tensor map, box = [16 rows, K]: rows past A are zero-filled by the hardware
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    KWS_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(KWS_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  CUtensorMap tmap;
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(A)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * sizeof(float)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(K), static_cast<cuuint32_t>(kOtChunkRows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(KWS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  const size_t smem = octbit_tc_smem(p.Npad, p.K);
  KWS_CUDA_OK(cudaFuncSetAttribute(octbit_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long ntiles = ceil_div(A, kOtTile);
  const long blocks = ntiles < sm_count() ? ntiles : sm_count();
  octbit_tc_kernel<<<static_cast<unsigned>(blocks), kOtThreads, smem, st>>>(tmap, p);
  KWS_LAUNCH_OK("octbit_tc_kernel");
  return KWS_OK;
}

}  // namespace kws
