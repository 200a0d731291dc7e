// K2+K3 on the 5th-generation tensor cores: persistent recurrent GRU (TF GRUCell semantics) + FC + softmax.
//
// Same maths as gru.cu (models/rnn_ctc.py:156-165,202-284; reset gate applied BEFORE the candidate matmul):
//   [r,u] = sigmoid([x,h] Wg + bg),  c = tanh([x, r*h] Wc + bc),  h' = u*h + (1-u)*c,  softmax(h' Wfc + bfc)
//
// Mapping to sm_100a:
//   * one CTA owns 128 streams for a whole layer -- the streams are the M dimension of tcgen05.mma
//     (M=128, cta_group::1); the TMEM lane of a stream is owned by the same thread for the whole chunk, so the
//     fp32 master state h never leaves registers between time steps;
//   * all weights of the layer ([384, K] fp16, 135-196 KB) are resident in shared memory for the whole
//     kernel, in the K-major SWIZZLE_NONE canonical layout, packed on the host at model creation;
//   * the A operand [x_t | h_{t-1}] (and [x_t | r*h]) lives in TENSOR MEMORY (TS-form MMA): each thread writes
//     its own stream's row with tcgen05.st, so activations never touch shared memory and no proxy fence is
//     needed.  TMEM map (512 columns): D_r 0..127, D_u 128..255, D_cand 256..383, A_x, A_h;
//   * the input projection is not a separate GEMM: x_t W[0:in] is accumulated into the same TMEM tile as
//     h W[in:], and it is issued a phase early so it never sits on the recurrent critical path;
//   * warp-specialised: warps 0-15 run the gate algebra from TMEM (warp w: the 32 streams of TMEM lane quarter w%4,
//     hidden units [32*(w/4), +32) -- four warps per scheduler hide the MUFU / TMEM-load latencies); warp 16 only
//     issues MMAs.  Everything is handed over with mbarriers (tcgen05.commit one way,
//     512-thread arrivals the other) -- no CTA-wide barrier inside the time loop.  Per step:
//         MMA warp                                   epilogue threads
//         r-gate x-part             <- A_x ready     ...
//         r-gate h-part -> commit R <- A_h ready
//         u-gate x+h    -> commit U                  wait R: r = sigmoid(D_r), keep fp16(r*h) in registers
//         cand  x-part                               wait U: A_h <- r*h, arrive;  D_u <- u = sigmoid(D_u) in place
//         cand  h-part  -> commit C <- A_rh ready    wait C: A_x <- x_{t+1}, arrive (next r-gate x-part overlaps)
//                                                    c = tanh(D_c), h' = c + u (h - c), FC partials, A_h <- h', arrive
//     so the dependent chain of a step is  r-MMA(h) -> sigmoid -> cand-MMA(h) -> tanh,  with the u gate, all
//     x-part MMAs, the softmax and the global loads/stores running beside it;
//   * gate algebra in fp32 with ex2.approx / rcp.approx and packed f32x2 arithmetic (FFMA2/FADD2/FMUL2); activations
//     are evaluated four at a time sharing one reciprocal (5 MUFU per 4 activations, 3.75 per hidden unit and step --
//     the MUFU pipe is what this kernel is bound by), biases pre-scaled by log2(e).
// Operands are fp16 (weights rounded once on the host, activations rounded when written to TMEM), accumulation
// and all state fp32.  The layer-0 input projection is the exception: mel features are unbounded (tens for loud
// audio) and a single fp16 rounding of x and W_x alone costs up to 3e-3 on the carried state, so that product is
// issued as three fp16 MMAs  x_hi*W_hi + x_lo*W_hi + x_hi*W_lo  (x = x_hi + x_lo, W = W_hi + W_lo; the
// dropped lo*lo term is 2^-22 relative), i.e. at fp32-grade accuracy for +6 small MMAs per step.  With it the
// deviation from the fp32 graph stays below 1e-3 (contract) on probabilities and state; see DESIGN.md.
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tc05.cuh"

namespace kws {

constexpr int kTcTile = 128;
constexpr int kTcEpiWarps = 16;           // warps 0-15: gate algebra.  Warp w owns TMEM lanes 32*(w%4).. (its 32 streams)
constexpr int kTcEpiThreads = 32 * kTcEpiWarps;   //   and the 32 hidden units [32*(w/4), +32) of those streams
constexpr int kTcThreads = kTcEpiThreads + 128;   // + warpgroup 4: warp 16 issues the MMAs (warps 17-19 only donate registers)
constexpr int kTcUnits = 32;              // hidden units per gate thread


struct GruTcParams {
  int kx;                      // x width padded to a multiple of 16
  int kxw;                     // x columns of the packed weights: 2*kx when the x product is split (layer 0), else kx
  int in_dim;                  // true x width
  long S;
  int n;                       // frames of the whole sequence: the time stride of x, the hand-off, probs and logits
  int t0, nt;                  // this launch runs frames [t0, t0 + nt) (the layers of long sequences are pipelined in time chunks)
  const float* x_f32;          // layer 0: mel fp32, [S, n, in_dim] row-major or stream-tiled (x_tiled, see common.cuh)
  int x_tiled;
  const __half* x_f16;         // layer > 0: fp16, stream-tiled [tile][t][16 chunks][128 streams][8 units]
  __half* y_f16;               // non-last layers: same tiled layout
  const __half* wpack;         // [384, kxw+128] fp16, canonical layout: [Wx_hi | Wx_lo (split only) | Wh]
  const float* bias;           // [384] = gates (r | u) | candidate
  const float* h_in;           // [S, 128]
  float* h_out;                // [S, 128]
  const int* seq_len;
  const unsigned char* zero_state;
  const float* fc_w;           // [128, C]
  const float* fc_b;           // [C]
  int C;
  float* probs;                // [S, n, C]
  float* logits;               // [S, n, C] or null
  int timeline;                // record g_tc_timeline (debug)
  // FC weights [128][8] + bias [8] by value (copied to shared memory by the last layer's CTAs)
  float fcw[kHidden * kTcMaxClasses];
  float fcb[kTcMaxClasses];
};

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(a + b) with nb = -log2(e) * b:   1 / (1 + 2^(-log2e*(a+b)))         FFMA, EX2, FADD, RCP
__device__ __forceinline__ float sigmoid_pre(float a, float nb) {
  return rcp_approx(1.0f + ex2_approx(fmaf(a, -kLog2e, nb)));
}
// tanh(a + b) with pb = 2*log2(e) * b:     1 - 2 / (1 + 2^(2*log2e*(a+b)))    FFMA, EX2, FADD, RCP, FFMA
__device__ __forceinline__ float tanh_pre(float a, float pb) {
  return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(fmaf(a, 2.0f * kLog2e, pb))), 1.0f);
}

// Two activations for three MUFU operations: 1/d0 and 1/d1 from ONE reciprocal of d0*d1.  The exponents are
// clamped to 63 so that the product stays finite (d <= 2^63 + 1; sigmoid below 1e-19 / tanh above 1 - 2e-19 are
// returned exactly as the clamp leaves them -- far below fp32 resolution of the results).
__device__ __forceinline__ void sigmoid2_pre(float a0, float nb0, float a1, float nb1, float& s0, float& s1) {
  const float d0 = 1.0f + ex2_approx(fminf(fmaf(a0, -kLog2e, nb0), 63.0f));
  const float d1 = 1.0f + ex2_approx(fminf(fmaf(a1, -kLog2e, nb1), 63.0f));
  const float r = rcp_approx(d0 * d1);
  s0 = r * d1;
  s1 = r * d0;
}
__device__ __forceinline__ void tanh2_pre(float a0, float pb0, float a1, float pb1, float& c0, float& c1) {
  const float d0 = 1.0f + ex2_approx(fminf(fmaf(a0, 2.0f * kLog2e, pb0), 63.0f));
  const float d1 = 1.0f + ex2_approx(fminf(fmaf(a1, 2.0f * kLog2e, pb1), 63.0f));
  const float r = rcp_approx(d0 * d1);
  c0 = fmaf(-2.0f, r * d1, 1.0f);
  c1 = fmaf(-2.0f, r * d0, 1.0f);
}

// Four activations for FIVE MUFU operations and packed f32x2 arithmetic: with d = 1 + 2^e, the four reciprocals
// come from ONE reciprocal of d0*d1*d2*d3:  P = (d0 d2, d1 d3),  r = 1/(P.x P.y),  Rinv = r (P.y, P.x) = (1/(d0 d2),
// 1/(d1 d3)),  (1/d0, 1/d1) = Rinv (d2, d3),  (1/d2, 1/d3) = Rinv (d0, d1).  The exponents are clamped to 31 so that the
// product stays finite (<= 2^124); a sigmoid below 5e-10 / a tanh above 1 - 1e-9 is returned as the clamp leaves it.
// Inputs: a = accumulator values, b = bias pre-scaled by `mul` (the activation is act(a + bias)), mul = -log2(e) for
// the sigmoid and 2 log2(e) for tanh.  Output: 1 / (1 + 2^(mul * (a + bias))) for the four inputs.
__device__ __forceinline__ void recip4_1p_exp2(float2 a01, float2 a23, float2 b01, float2 b23, float mul,
                                               float2& s01, float2& s23) {
  const float2 m2 = make_float2(mul, mul);
  float2 e01 = __ffma2_rn(a01, m2, b01), e23 = __ffma2_rn(a23, m2, b23);
  e01.x = ex2_approx(fminf(e01.x, 31.0f));
  e01.y = ex2_approx(fminf(e01.y, 31.0f));
  e23.x = ex2_approx(fminf(e23.x, 31.0f));
  e23.y = ex2_approx(fminf(e23.y, 31.0f));
  const float2 one = make_float2(1.0f, 1.0f);
  const float2 d01 = __fadd2_rn(e01, one), d23 = __fadd2_rn(e23, one);
  const float2 pp = __fmul2_rn(d01, d23);
  const float r = rcp_approx(pp.x * pp.y);
  const float2 rinv = __fmul2_rn(make_float2(r, r), make_float2(pp.y, pp.x));
  s01 = __fmul2_rn(rinv, d23);
  s23 = __fmul2_rn(rinv, d01);
}
__device__ __forceinline__ void sigmoid4_pre(float2 a01, float2 a23, float2 nb01, float2 nb23, float2& s01, float2& s23) {
  recip4_1p_exp2(a01, a23, nb01, nb23, -kLog2e, s01, s23);
}
__device__ __forceinline__ void tanh4_pre(float2 a01, float2 a23, float2 pb01, float2 pb23, float2& c01, float2& c23) {
  float2 s01, s23;
  recip4_1p_exp2(a01, a23, pb01, pb23, 2.0f * kLog2e, s01, s23);
  const float2 m2 = make_float2(-2.0f, -2.0f), one = make_float2(1.0f, 1.0f);
  c01 = __ffma2_rn(s01, m2, one);
  c23 = __ffma2_rn(s23, m2, one);
}
__device__ __forceinline__ float2 u2f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// this thread's TMEM stores are complete and ordered before the arrival
__device__ __forceinline__ void tmem_publish(uint64_t* bar) {
  tc::wait_st();
  tc::fence_before_sync();
  mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_acquire(uint64_t* bar, uint32_t parity) {
  tc::mbar_wait(bar, parity);
  tc::fence_after_sync();
}

// Optional phase timeline of the first tile of CTA 0 (clock64 at each hand-over), for kws_debug_tc_timeline.
__device__ long long g_tc_timeline[64 * 8];
static int g_tc_timeline_on = 0;

enum { kBarAX = 0, kBarAH, kBarARH, kBarR, kBarU, kBarC, kNumBars };

// kFirst: layer 0 -- x is fp32 (mel) and its projection uses the 3-term split.
// kFirst: layer 0 -- x is fp32 (mel) and its projection uses the 3-term split.
template <bool kLast, bool kFirst>
__global__ void __launch_bounds__(kTcThreads, 1)
gru_tc_kernel(const GruTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int ktot = p.kxw + kHidden;                                                  // K extent of the packed weights
  unsigned char* sW = smem;                                                         // [384, ktot] fp16
  float* sBias = reinterpret_cast<float*>(smem + static_cast<size_t>(384) * ktot * 2);   // [384] pre-scaled
  float* sXch = sBias + 384;                                               // [3][128][8] FC partials of unit blocks 1..3
  float* sFc = sXch + 3 * kTcTile * kTcMaxClasses;                         // [128][8] FC weights (last layer)
  float* sProb = sFc + kHidden * kTcMaxClasses;                            // [4 steps][6][128] parked probabilities (last layer)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sProb + (kLast ? 4 * 6 * kTcTile : 0));   // [kNumBars]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kMmaWarp = kTcEpiWarps;

  if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bars[kBarAX], kTcEpiThreads);
    tc::mbar_init(&bars[kBarAH], kTcEpiThreads);
    tc::mbar_init(&bars[kBarARH], kTcEpiThreads);
    tc::mbar_init(&bars[kBarR], 1);
    tc::mbar_init(&bars[kBarU], 1);
    tc::mbar_init(&bars[kBarC], 1);
    tc::mbar_fence_init();
  }
  {
    const int n16 = 384 * ktot * 2 / 16;
    const uint4* src = reinterpret_cast<const uint4*>(p.wpack);
    for (int i = tid; i < n16; i += kTcThreads) reinterpret_cast<uint4*>(sW)[i] = __ldg(src + i);
    for (int i = tid; i < 384; i += kTcThreads)
      sBias[i] = p.bias[i] * (i < 2 * kHidden ? -kLog2e : 2.0f * kLog2e);
    if (kLast)
      for (int i = tid; i < kHidden * kTcMaxClasses; i += kTcThreads) sFc[i] = p.fcw[i];
  }
  tc::fence_proxy_async();            // weights written with generic stores, read by the MMA (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const uint32_t tmem = *tmem_slot;
  const uint32_t colDr = 0, colDu = 128, colDc = 256, colAx = 384;
  const uint32_t colAxl = colAx + p.kx / 2;                       // x_lo (split only)
  const bool split = kFirst && p.kxw != p.kx;                     // 3-term x product (needs 2*kx/2 + 64 <= 128 columns)
  const uint32_t colAh = colAx + p.kxw / 2;
  const long ntiles = (p.S + kTcTile - 1) / kTcTile;

  if (warp >= kMmaWarp) {
    // =========================================================== MMA issuer (one elected lane of warp 16)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == kMmaWarp) {
      const uint32_t sbo = static_cast<uint32_t>(ktot / 8) * 128;
      const uint32_t idesc128 = tc::idesc_f16(128, 128);
      const int nx = p.kx / 16, nxw = p.kxw / 16;
      const bool lead = lane == 0;
      // All operand addresses are computed warp-uniformly (they live in uniform registers); only the MMA itself is
      // predicated on the elected lane.  B descriptor of K-chunk k16 for the weight rows starting at row0 =
      // base + ((row0/8)*sbo + 256*k16)/16 added to the 14-bit start-address field (shared memory < 256 KB: no carry).
      const uint64_t wbase = tc::smem_desc(tc::smem_u32(sW), 128, sbo);
      const uint32_t row_step = (16 * sbo) >> 4;                             // 128 output rows
      // x-part of a product into D columns `dcol` (overwrites D), weight rows starting at 128*rblk
      auto issue_x = [&](uint32_t dcol, int rblk) {
        const uint64_t wrow = wbase + rblk * row_step;
#pragma unroll 1
        for (int k16 = 0; k16 < nx; ++k16)                                   // x_hi * Wx_hi
          if (lead) tc::mma_ts(tmem + dcol, tmem + colAx + 8 * k16, wrow + 16 * k16, idesc128, k16 > 0);
        if (split) {
#pragma unroll 1
          for (int k16 = 0; k16 < nx; ++k16)                                 // x_lo * Wx_hi
            if (lead) tc::mma_ts(tmem + dcol, tmem + colAxl + 8 * k16, wrow + 16 * k16, idesc128, true);
#pragma unroll 1
          for (int k16 = 0; k16 < nx; ++k16)                                 // x_hi * Wx_lo
            if (lead) tc::mma_ts(tmem + dcol, tmem + colAx + 8 * k16, wrow + 16 * (nx + k16), idesc128, true);
        }
      };
      auto issue_h = [&](uint32_t dcol, int rblk) {                          // += A_h * Wh[128*rblk .. +127]
        const uint64_t wrow = wbase + rblk * row_step + 16 * nxw;
#pragma unroll 1
        for (int k16 = 0; k16 < kHidden / 16; ++k16)
          if (lead) tc::mma_ts(tmem + dcol, tmem + colAh + 8 * k16, wrow + 16 * k16, idesc128, true);
      };
      uint32_t it = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int t = 0; t < p.nt; ++t, ++it) {
          const uint32_t par = it & 1;
          mbar_acquire(&bars[kBarAX], par);
          issue_x(colDr, 0);                                                 // r gate, x-part (D_r is free since the last r epilogue)
          mbar_acquire(&bars[kBarAH], par);
          issue_h(colDr, 0);
          if (lead) tc::commit(&bars[kBarR]);
          issue_x(colDu, 1);                                                 // u gate: D_u held the activated u until now
          issue_h(colDu, 1);
          if (lead) tc::commit(&bars[kBarU]);
          issue_x(colDc, 2);                                                 // candidate, x-part
          mbar_acquire(&bars[kBarARH], par);
          issue_h(colDc, 2);
          if (lead) tc::commit(&bars[kBarC]);
        }
      }
    }
  } else {
    // =========================================================== gate algebra (512 threads: one stream x 32 units each)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");      // launched at 96: 128 x (96-32) released = 512 x (112-96) acquired
    const int quarter = warp & 3, ublk = warp >> 2;
    const int row = 32 * quarter + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(32 * quarter) << 16;
    const int u0 = kTcUnits * ublk;                                 // first hidden unit of this thread
    const uint32_t my_ah = tmem + lane_sel + colAh + 16 * ublk;     // 32 units = 16 columns
    const int xq = p.kx / 16;                                       // layer 0: float4 chunks of x per thread
    const uint32_t my_ax = tmem + lane_sel + colAx + (p.kx / 8) * ublk;
    const float* bR = sBias + u0;
    const float* bU = sBias + kHidden + u0;
    const float* bC = sBias + 2 * kHidden + u0;
    // last layer, the model's 6 classes, 16-byte aligned rows: probabilities leave four steps at a time (fc_finish)
    const bool probs_batched = kLast && p.C == 6 && ((p.n | p.t0) & 1) == 0 && (reinterpret_cast<uintptr_t>(p.probs) & 15) == 0;
    uint32_t it = 0;

    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long s = tile * kTcTile + row;
      const bool ok = s < p.S;
      const long sr = ok ? s : 0;
      const int t_end = p.t0 + p.nt;
      const int len = ok ? (p.seq_len ? p.seq_len[s] : p.n) : 0;
      const bool all_live = __all_sync(0xffffffffu, len >= t_end);   // warp-uniform: no select in the update
      float h[kTcUnits];
      {
        const bool zero = !ok || (p.zero_state && p.zero_state[s]);
        const float4* src = reinterpret_cast<const float4*>(p.h_in + sr * kHidden + u0);
#pragma unroll
        for (int i = 0; i < kTcUnits / 4; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!zero) v = src[i];
          h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
        }
      }
      // layer 0: this thread's first 16-byte chunk of step 0, the chunk strides per step and per chunk, its valid chunks
      const int x_k0 = (p.kx / 4) * ublk;
      int nqv = (p.in_dim - x_k0 + 3) >> 2;
      nqv = nqv < 0 ? 0 : (nqv > xq ? xq : nqv);
      if (!p.x_tiled && !ok) nqv = 0;                              // row-major: rows past S do not exist
      const bool x_vec = p.x_tiled || ((p.in_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x_f32) & 15) == 0);
      const long x_q4 = p.in_dim >> 2;
      const long x_step = p.x_tiled ? x_q4 * kTcTile : x_q4;
      const int x_qs = p.x_tiled ? kTcTile : 1;
      const float4* x_base = reinterpret_cast<const float4*>(p.x_f32) +
                             (p.x_tiled ? (tile * p.n * x_q4 + (x_k0 >> 2)) * kTcTile + row : sr * p.n * x_q4 + (x_k0 >> 2));
      // x_t of this thread: K elements [ (kx/4)*ublk, +kx/4 ) of the row, as packed fp16 pairs (hi, and lo when split)
      // layer 0 keeps the raw fp32 values (xf) and splits them into fp16 hi/lo only when they are written to TMEM, so
      // the global loads issued under the u gate are not waited for until the candidate phase
      uint32_t xr[kFirst ? 1 : 16];
      float4 xf[kFirst ? 8 : 1];
      auto load_x = [&](int t) {
        if (!kFirst) {
          // chunk q of stream `row` sits at ((tile*n + t)*16 + q)*128 + row: a warp reads 512 contiguous bytes
          const uint4* src = reinterpret_cast<const uint4*>(p.x_f16) + ((tile * p.n + t) * 16 + 4 * ublk) * kTcTile + row;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(src + q * kTcTile);
            xr[kFirst ? 0 : 4 * q] = v.x; xr[kFirst ? 0 : 4 * q + 1] = v.y;
            xr[kFirst ? 0 : 4 * q + 2] = v.z; xr[kFirst ? 0 : 4 * q + 3] = v.w;
          }
        } else if (x_vec) {
          // 16-byte chunks: stream-tiled mel (chunk c of stream `row` at ((tile*n + t)*Q + c)*128 + row, Q = in_dim/4) or
          // row-major rows whose chunks are consecutive.  Base address, strides and this thread's number of chunks are
          // fixed per tile: per step one multiply and up to 8 predicated loads
          const float4* src = x_base + static_cast<long>(t) * x_step;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < nqv) v = __ldg(src + q * x_qs);
            xf[kFirst ? q : 0] = v;
          }
        } else {
          // row-major [S, n, in_dim] with in_dim % 4 != 0 or an unaligned base: element-wise with bounds (rare, slow)
          const float* src = p.x_f32 + (sr * p.n + t) * static_cast<long>(p.in_dim) + x_k0;
#pragma unroll 1
          for (int q = 0; q < 8; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = x_k0 + 4 * q;
            if (q < nqv) {
              v.x = __ldg(src + 4 * q);
              if (k + 1 < p.in_dim) v.y = __ldg(src + 4 * q + 1);
              if (k + 2 < p.in_dim) v.z = __ldg(src + 4 * q + 2);
              if (k + 3 < p.in_dim) v.w = __ldg(src + 4 * q + 3);
            }
            // (no dynamic register indexing: the loop is not unrolled)
            if (q == 0) xf[0] = v;
            if (kFirst) {
              if (q == 1) xf[kFirst ? 1 : 0] = v;
              if (q == 2) xf[kFirst ? 2 : 0] = v;
              if (q == 3) xf[kFirst ? 3 : 0] = v;
              if (q == 4) xf[kFirst ? 4 : 0] = v;
              if (q == 5) xf[kFirst ? 5 : 0] = v;
              if (q == 6) xf[kFirst ? 6 : 0] = v;
              if (q == 7) xf[kFirst ? 7 : 0] = v;
            }
          }
        }
      };
      auto store_x = [&]() {
        if (!kFirst) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t v[4] = {xr[kFirst ? 0 : 4 * q], xr[kFirst ? 0 : 4 * q + 1], xr[kFirst ? 0 : 4 * q + 2],
                                   xr[kFirst ? 0 : 4 * q + 3]};
            tc::st4(my_ax + 4 * q, v);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < xq) {
              const float4 v = xf[kFirst ? q : 0];
              const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
              tc::st2(my_ax + 2 * q, *reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
              if (split) {
                const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
                tc::st2(my_ax + p.kx / 2 + 2 * q, tc::pack_half2(v.x - b0.x, v.y - b0.y), tc::pack_half2(v.z - b1.x, v.w - b1.y));
              }
            }
        }
      };
      auto store_h = [&]() {                                       // A_h <- fp16(h)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = tc::pack_half2(h[16 * c + 2 * i], h[16 * c + 2 * i + 1]);
          tc::st8(my_ah + 8 * c, v);
        }
      };
      // FC + softmax of a step are computed right after its h' has been published, i.e. while the next step's
      // r-gate MMAs run: 32 units per thread, partial sums of unit blocks 1..3 handed to block 0 through smem.
      auto fc_finish = [&](int t_done, bool emit) {
        const bool tl2 = p.timeline && blockIdx.x == 0 && tid == 0 && tile == blockIdx.x && t_done < 32;
        // class pairs (2c, 2c+1) are accumulated with packed FFMA2: h[j] broadcast x a weight pair; the [128][8] weights
        // sit in shared memory and every lane reads the same row (broadcast, one wavefront per load)
        float2 part2[kTcMaxClasses / 2];
#pragma unroll
        for (int c = 0; c < kTcMaxClasses / 2; ++c) part2[c] = make_float2(0.0f, 0.0f);
        if (emit) {                                                 // dynamic_rnn: zero output past the length
          const float4* wrow = reinterpret_cast<const float4*>(sFc + u0 * kTcMaxClasses);
          if (p.C <= 6) {                                           // the model's 6 classes: 3 pairs
#pragma unroll
            for (int j = 0; j < kTcUnits; ++j) {
              const float4 wa = wrow[2 * j];
              const float2 wb = *reinterpret_cast<const float2*>(wrow + 2 * j + 1);
              const float2 hh = make_float2(h[j], h[j]);
              part2[0] = __ffma2_rn(hh, make_float2(wa.x, wa.y), part2[0]);
              part2[1] = __ffma2_rn(hh, make_float2(wa.z, wa.w), part2[1]);
              part2[2] = __ffma2_rn(hh, wb, part2[2]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < kTcUnits; ++j) {
              const float4 wa = wrow[2 * j], wb = wrow[2 * j + 1];
              const float2 hh = make_float2(h[j], h[j]);
              part2[0] = __ffma2_rn(hh, make_float2(wa.x, wa.y), part2[0]);
              part2[1] = __ffma2_rn(hh, make_float2(wa.z, wa.w), part2[1]);
              part2[2] = __ffma2_rn(hh, make_float2(wb.x, wb.y), part2[2]);
              part2[3] = __ffma2_rn(hh, make_float2(wb.z, wb.w), part2[3]);
            }
          }
        }
        const float part[kTcMaxClasses] = {part2[0].x, part2[0].y, part2[1].x, part2[1].y,
                                           part2[2].x, part2[2].y, part2[3].x, part2[3].y};
        if (tl2) g_tc_timeline[256 + t_done * 4 + 0] = clock64();
        if (ublk > 0) {
          float* dst = sXch + ((ublk - 1) * kTcTile + row) * 8;
          *reinterpret_cast<float4*>(dst) = make_float4(part[0], part[1], part[2], part[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(part[4], part[5], part[6], part[7]);
          asm volatile("bar.arrive 1, 512;" ::: "memory");
        } else {
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (tl2) g_tc_timeline[256 + t_done * 4 + 1] = clock64();
          if (ok) {
            float lg[8];
            float mx = -INFINITY;
            {
              // the three other unit blocks' partial sums: 16-byte reads (a row is 32 bytes: conflict-free)
              float o[3][8];
#pragma unroll
              for (int b = 0; b < 3; ++b) {
                const float4 lo4 = *reinterpret_cast<const float4*>(sXch + (b * kTcTile + row) * 8);
                const float4 hi4 = *reinterpret_cast<const float4*>(sXch + (b * kTcTile + row) * 8 + 4);
                o[b][0] = lo4.x; o[b][1] = lo4.y; o[b][2] = lo4.z; o[b][3] = lo4.w;
                o[b][4] = hi4.x; o[b][5] = hi4.y; o[b][6] = hi4.z; o[b][7] = hi4.w;
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                lg[c] = part[c] + o[0][c] + o[1][c] + o[2][c] + p.fcb[c];
                if (c < p.C) mx = fmaxf(mx, lg[c]);
              }
            }
            float e[8], sum = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              e[c] = c < p.C ? ex2_approx((lg[c] - mx) * kLog2e) : 0.0f;    // 2 ulp: far inside the 1e-3 contract
              sum += e[c];
            }
            const float inv = rcp_approx(sum);                              // sum in [1, C]
            if (probs_batched) {
              // Probabilities of four consecutive steps are parked in shared memory (a column per stream, touched by
              // this thread only) and leave as 96 contiguous bytes per stream: 3 full sectors instead of 4 x 24
              // scattered bytes.
              const int kq = (t_done - p.t0) & 3;
#pragma unroll
              for (int c = 0; c < 6; ++c) sProb[(kq * 6 + c) * kTcTile + row] = e[c] * inv;
              if (kq == 3 || t_done == t_end - 1) {
                float4* dst = reinterpret_cast<float4*>(p.probs + (s * p.n + (t_done - kq)) * 6);
                const int nq4 = ((kq + 1) * 6) >> 2;                 // whole float4s: 6 (4 steps), 4, 3, 1
#pragma unroll
                for (int i = 0; i < 6; ++i)
                  if (i < nq4)
                    dst[i] = make_float4(sProb[(4 * i) * kTcTile + row], sProb[(4 * i + 1) * kTcTile + row],
                                         sProb[(4 * i + 2) * kTcTile + row], sProb[(4 * i + 3) * kTcTile + row]);
                float* tail = reinterpret_cast<float*>(dst) + 4 * nq4;   // (kq+1)*6 mod 4 = 2 left over when kq is 0 or 2
                if (((kq + 1) * 6) & 3) {
                  tail[0] = sProb[(4 * nq4) * kTcTile + row];
                  tail[1] = sProb[(4 * nq4 + 1) * kTcTile + row];
                }
              }
            } else {
              float* pr = p.probs + (s * p.n + t_done) * p.C;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                if (c < p.C) pr[c] = e[c] * inv;
            }
            if (p.logits) {
              float* lo = p.logits + (s * p.n + t_done) * p.C;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                if (c < p.C) lo[c] = lg[c];
            }
          }
          if (tl2) g_tc_timeline[256 + t_done * 4 + 2] = clock64();
        }
      };

      // ---- prologue: A_x <- x_0, A_h <- fp16(h).  All MMAs of the previous tile have completed (its last
      // commit was waited for by every thread), so both operand regions are free.
      load_x(p.t0);
      store_x();
      store_h();
      tc::wait_st();
      tc::fence_before_sync();
      mbar_arrive(&bars[kBarAX]);
      mbar_arrive(&bars[kBarAH]);

      for (int t = p.t0; t < t_end; ++t, ++it) {
        const uint32_t par = it & 1;
        const bool more = t + 1 < t_end;
        const bool tl = p.timeline && blockIdx.x == 0 && tid == 0 && tile == blockIdx.x && t < 64;
        // ---- r gate: keep fp16(r*h) in registers until the u-gate MMAs have finished reading A_h
        uint32_t rh[kTcUnits / 2];
        if (tl) g_tc_timeline[t * 8 + 0] = clock64();
        mbar_acquire(&bars[kBarR], par);
        if (tl) g_tc_timeline[t * 8 + 1] = clock64();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[16];
          tc::ld16(tmem + lane_sel + colDr + u0 + 16 * c, v);
          tc::wait_ld();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int j = 16 * c + i;
            const float4 nb = *reinterpret_cast<const float4*>(bR + j);
            float2 r01, r23;
            sigmoid4_pre(u2f2(v[i], v[i + 1]), u2f2(v[i + 2], v[i + 3]), make_float2(nb.x, nb.y), make_float2(nb.z, nb.w), r01, r23);
            const float2 rh01 = __fmul2_rn(r01, make_float2(h[j], h[j + 1]));
            const float2 rh23 = __fmul2_rn(r23, make_float2(h[j + 2], h[j + 3]));
            rh[j / 2] = tc::pack_half2(rh01.x, rh01.y);
            rh[j / 2 + 1] = tc::pack_half2(rh23.x, rh23.y);
          }
        }
        if (tl) g_tc_timeline[t * 8 + 2] = clock64();
        mbar_acquire(&bars[kBarU], par);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t v[8] = {rh[8 * c], rh[8 * c + 1], rh[8 * c + 2], rh[8 * c + 3],
                                 rh[8 * c + 4], rh[8 * c + 5], rh[8 * c + 6], rh[8 * c + 7]};
          tc::st8(my_ah + 8 * c, v);
        }
        tmem_publish(&bars[kBarARH]);
        if (tl) g_tc_timeline[t * 8 + 3] = clock64();
        if (more) load_x(t + 1);                                   // coalesced global loads in flight under the u gate
        // ---- u gate (beside the candidate MMAs): activated in place, D_u keeps u until the update reads it
        // (holding u in registers instead was measured slower: it pushes the gate threads past their 112 registers)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[16];
          tc::ld16(tmem + lane_sel + colDu + u0 + 16 * c, v);
          tc::wait_ld();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 nb = *reinterpret_cast<const float4*>(bU + 16 * c + i);
            float2 u01, u23;
            sigmoid4_pre(u2f2(v[i], v[i + 1]), u2f2(v[i + 2], v[i + 3]), make_float2(nb.x, nb.y), make_float2(nb.z, nb.w), u01, u23);
            v[i] = __float_as_uint(u01.x);
            v[i + 1] = __float_as_uint(u01.y);
            v[i + 2] = __float_as_uint(u23.x);
            v[i + 3] = __float_as_uint(u23.y);
          }
          tc::st16(tmem + lane_sel + colDu + u0 + 16 * c, v);
        }
        tc::wait_st();
        // ---- candidate and state update: only what the next step's MMAs wait for
        if (tl) g_tc_timeline[t * 8 + 4] = clock64();
        mbar_acquire(&bars[kBarC], par);
        if (tl) g_tc_timeline[t * 8 + 5] = clock64();
        if (more) {                                                // next step's x first: its r-gate MMAs run under this phase
          store_x();
          tmem_publish(&bars[kBarAX]);
        }
        const bool live = t < len;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t vc[16], vu[16];
          tc::ld16(tmem + lane_sel + colDc + u0 + 16 * c, vc);
          tc::ld16(tmem + lane_sel + colDu + u0 + 16 * c, vu);
          tc::wait_ld();
          uint32_t packed[8];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int j = 16 * c + i;
            const float4 pb = *reinterpret_cast<const float4*>(bC + j);
            float2 c01, c23;
            tanh4_pre(u2f2(vc[i], vc[i + 1]), u2f2(vc[i + 2], vc[i + 3]), make_float2(pb.x, pb.y), make_float2(pb.z, pb.w), c01, c23);
            const float2 neg = make_float2(-1.0f, -1.0f);
            const float2 h01 = make_float2(h[j], h[j + 1]), h23 = make_float2(h[j + 2], h[j + 3]);
            float2 n01 = __ffma2_rn(u2f2(vu[i], vu[i + 1]), __ffma2_rn(c01, neg, h01), c01);          // u*h + (1-u)*c
            float2 n23 = __ffma2_rn(u2f2(vu[i + 2], vu[i + 3]), __ffma2_rn(c23, neg, h23), c23);
            if (!all_live) {                                       // dynamic_rnn: state carried past the length
              n01 = live ? n01 : h01;
              n23 = live ? n23 : h23;
            }
            h[j] = n01.x;
            h[j + 1] = n01.y;
            h[j + 2] = n23.x;
            h[j + 3] = n23.y;
            packed[i / 2] = tc::pack_half2(n01.x, n01.y);
            packed[i / 2 + 1] = tc::pack_half2(n23.x, n23.y);
          }
          if (more) tc::st8(my_ah + 8 * c, packed);
        }
        if (more) tmem_publish(&bars[kBarAH]);                     // releases the next step's r/u MMAs
        if (tl) g_tc_timeline[t * 8 + 6] = clock64();
        // ---- outputs of this step, in the shadow of the next step's r-gate MMAs.  dynamic_rnn: zero output past the length
        const bool emit = all_live || live;
        if (!kLast) {
          // tiled hand-off (rows past S are written too: the scratch is padded to whole tiles)
          uint4* dst = reinterpret_cast<uint4*>(p.y_f16) + ((tile * p.n + t) * 16 + 4 * ublk) * kTcTile + row;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = emit ? tc::pack_half2(h[8 * q], h[8 * q + 1]) : 0u;
            v.y = emit ? tc::pack_half2(h[8 * q + 2], h[8 * q + 3]) : 0u;
            v.z = emit ? tc::pack_half2(h[8 * q + 4], h[8 * q + 5]) : 0u;
            v.w = emit ? tc::pack_half2(h[8 * q + 6], h[8 * q + 7]) : 0u;
            dst[q * kTcTile] = v;
          }
        } else {
          fc_finish(t, emit);
        }
      }
      if (p.timeline && blockIdx.x == 0 && tid == 0 && tile == blockIdx.x) g_tc_timeline[7] = clock64();
      if (ok) {
        float4* dst = reinterpret_cast<float4*>(p.h_out + s * kHidden + u0);
#pragma unroll
        for (int i = 0; i < kTcUnits / 4; ++i) dst[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc::tmem_dealloc(tmem, 512);
}


static size_t gru_tc_smem_bytes(int ktot, bool last) {
  return static_cast<size_t>(384) * ktot * 2 + sizeof(float) * (384 + 3 * kTcTile * 8 + kHidden * kTcMaxClasses) +
         (last ? sizeof(float) * 4 * 6 * kTcTile : 0) + kNumBars * sizeof(uint64_t) + 16;
}

// Pack one layer's TF kernels into the fp16 canonical [384, kxw+128] B operand: [Wx_hi | Wx_lo (split) | Wh].
void pack_tc_weights(const float* gates_kernel, const float* cand_kernel, int in_dim, bool split, std::vector<__half>* out,
                     int* kx_out, int* kxw_out) {
  const int kx = (in_dim + 15) / 16 * 16;
  const int kxw = split ? 2 * kx : kx;
  const int ktot = kxw + kHidden;
  out->assign(static_cast<size_t>(384) * ktot, __float2half_rn(0.0f));
  unsigned char* base = reinterpret_cast<unsigned char*>(out->data());
  for (int n = 0; n < 384; ++n)
    for (int k = 0; k < ktot; ++k) {
      int src_row;
      bool lo = false;
      if (k < kxw) {
        const int kk = k % kx;
        lo = k >= kx;
        if (kk >= in_dim) continue;
        src_row = kk;
      } else {
        src_row = in_dim + (k - kxw);
      }
      const float v = n < 2 * kHidden ? gates_kernel[static_cast<size_t>(src_row) * 2 * kHidden + n]
                                      : cand_kernel[static_cast<size_t>(src_row) * kHidden + (n - 2 * kHidden)];
      const __half hi = __float2half_rn(v);
      const __half val = lo ? __float2half_rn(v - __half2float(hi)) : hi;
      *reinterpret_cast<__half*>(base + tc::canon_offset(n, k, ktot)) = val;
    }
  *kx_out = kx;
  *kxw_out = kxw;
}

int launch_gru_tc(kws_model* m, const GruArgs& a, cudaStream_t st) {
  if (a.S <= 0) return KWS_OK;
  if (m->cfg.num_classes > kTcMaxClasses)
    return fail(KWS_ERR_UNSUPPORTED, "the tensor-core recurrent kernel keeps %d classes, the model has %d", kTcMaxClasses,
                m->cfg.num_classes);
  const int L = m->cfg.num_layers;
  if (a.n <= 0) {
    if (a.state_out != a.state_in)
      KWS_CUDA_OK(cudaMemcpyAsync(a.state_out, a.state_in, sizeof(float) * L * a.S * kHidden, cudaMemcpyDeviceToDevice, st));
    return KWS_OK;
  }
  if (L > 1 && !a.seq_scratch) {
    const int rc = kws_model_reserve(m, a.S, a.n);
    if (rc != KWS_OK) return rc;
  }
  const long ntiles = ceil_div(a.S, kTcTile);
  const size_t per_buf = static_cast<size_t>(ntiles) * kTcTile * a.n * kHidden;     // halves (whole tiles)
  __half* seq = reinterpret_cast<__half*>(a.seq_scratch ? a.seq_scratch : m->scratch_seq);
  // frames [t0, t0 + nt) of layer l on stream `cs`; `head` = the first chunk of the sequence (incoming state, VAD reset)
  auto launch_layer = [&](int l, int t0, int nt, bool head, cudaStream_t cs) -> int {
    const bool last = l == L - 1;
    GruTcParams p;
    p.kx = m->layer[l].tc_kx;
    p.kxw = m->layer[l].tc_kxw;
    p.in_dim = m->layer[l].in_dim;
    p.S = a.S;
    p.n = a.n;
    p.t0 = t0;
    p.nt = nt;
    p.x_f32 = l == 0 ? a.x : nullptr;
    p.x_tiled = l == 0 && a.x_tiled ? 1 : 0;
    p.x_f16 = l == 0 ? nullptr : seq + ((l - 1) & 1) * per_buf;
    p.y_f16 = last ? nullptr : seq + (l & 1) * per_buf;
    p.wpack = static_cast<const __half*>(m->layer[l].tc_wpack);
    p.bias = m->layer[l].tc_bias;
    p.h_out = a.state_out + static_cast<long>(l) * a.S * kHidden;
    p.h_in = head ? a.state_in + static_cast<long>(l) * a.S * kHidden : p.h_out;
    p.seq_len = a.seq_len;
    p.zero_state = head ? a.zero_state : nullptr;
    p.fc_w = m->fc_w;
    p.fc_b = m->fc_b;
    p.C = m->cfg.num_classes;
    p.probs = a.probs;
    p.logits = a.logits;
    p.timeline = (g_tc_timeline_on == 1 || g_tc_timeline_on == 2 + l) ? 1 : 0;      // 1: every layer (the last one stays), 2 + l: layer l only
    for (int j = 0; j < kHidden; ++j)
      for (int c = 0; c < kTcMaxClasses; ++c) p.fcw[j * kTcMaxClasses + c] = c < m->cfg.num_classes ? m->fc_w_host[j * m->cfg.num_classes + c] : 0.0f;
    for (int c = 0; c < kTcMaxClasses; ++c) p.fcb[c] = c < m->cfg.num_classes ? m->fc_b_host[c] : 0.0f;
    const size_t smem = gru_tc_smem_bytes(p.kxw + kHidden, last);
    const long blocks = ntiles < sm_count() ? ntiles : sm_count();
    const bool first = l == 0;
    auto launch = [&](auto kernel) -> int {
      KWS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kernel<<<static_cast<unsigned>(blocks), kTcThreads, smem, cs>>>(p);
      return KWS_OK;
    };
    int rc;
    if (last) rc = first ? launch(gru_tc_kernel<true, true>) : launch(gru_tc_kernel<true, false>);
    else rc = first ? launch(gru_tc_kernel<false, true>) : launch(gru_tc_kernel<false, false>);
    if (rc != KWS_OK) return rc;
    KWS_LAUNCH_OK("gru_tc_kernel");
    return KWS_OK;
  };

  // Small batches of long sequences (BASELINE config 2: 4096 utterances x 298 frames = 32 tiles on 148 SMs): the two
  // layers are pipelined in time chunks on two streams -- layer 1 runs chunk c while layer 0 runs chunk c + 1 -- so
  // twice as many SMs work.  The recurrence itself is sequential; large batches fill the GPU without this.
  constexpr int kChunk = 64;                                   // multiple of 4: the batched probability stores stay aligned
  const bool pipelined = L == 2 && 2 * ntiles <= sm_count() && a.n >= 2 * kChunk && !g_tc_timeline_on;
  if (!pipelined) {
    for (int l = 0; l < L; ++l) {
      const int rc = launch_layer(l, 0, a.n, true, st);
      if (rc != KWS_OK) return rc;
    }
    return KWS_OK;
  }
  if (!m->aux_stream) KWS_CUDA_OK(cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
  const int nchunks = a.n / kChunk;                            // the last chunk takes the remainder
  while (static_cast<int>(m->aux_events.size()) < nchunks + 1) {
    cudaEvent_t e;
    KWS_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    m->aux_events.push_back(e);
  }
  for (int c = 0; c < nchunks; ++c) {
    const int t0 = c * kChunk;
    const int nt = c == nchunks - 1 ? a.n - t0 : kChunk;
    int rc = launch_layer(0, t0, nt, c == 0, st);
    if (rc != KWS_OK) return rc;
    KWS_CUDA_OK(cudaEventRecord(m->aux_events[c], st));
    KWS_CUDA_OK(cudaStreamWaitEvent(m->aux_stream, m->aux_events[c], 0));
    rc = launch_layer(1, t0, nt, c == 0, m->aux_stream);
    if (rc != KWS_OK) return rc;
  }
  KWS_CUDA_OK(cudaEventRecord(m->aux_events[nchunks], m->aux_stream));
  KWS_CUDA_OK(cudaStreamWaitEvent(st, m->aux_events[nchunks], 0));
  return KWS_OK;
}

}  // namespace kws

// Debug: phase timeline (SM clock ticks) of the first tile of CTA 0 of the LAST tensor-core GRU launch:
// per step t, [t*8 + k] = clock at {0 wait R, 1 got R, 2 r done, 3 r*h published, 4 u done, 5 got C, 6 h' published};
// slot [7] = end of the tile.
extern "C" int kws_debug_tc_timeline(int enable, long long* host_out, int count) {
  using namespace kws;
  clear_error();
  g_tc_timeline_on = enable;
  if (host_out && count > 0) {
    if (count > 64 * 8) count = 64 * 8;
    KWS_CUDA_OK(cudaMemcpyFromSymbol(host_out, g_tc_timeline, sizeof(long long) * count));
  }
  return KWS_OK;
}
