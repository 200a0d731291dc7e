// K2+K3 on the 5th-generation tensor cores: persistent recurrent GRU (TF GRUCell semantics) + FC + softmax.
//
// Same maths as gru.cu (models/rnn_ctc.py:156-165,202-284; reset gate applied BEFORE the candidate matmul):
//   [r,u] = sigmoid([x,h] Wg + bg),  c = tanh([x, r*h] Wc + bc),  h' = u*h + (1-u)*c,  softmax(h' Wfc + bfc)
//
// Mapping to sm_100a:
//   * one CTA owns 128 streams for a whole layer -- the streams are the M dimension of tcgen05.mma
//     (M=128, cta_group::1); the TMEM lane of a stream is owned by the same thread for the whole chunk, so the
//     fp32 master state h never leaves registers between time steps;
//   * all weights of the layer ([384, in+128] fp16, 135-196 KB) are resident in shared memory for the whole
//     kernel, in the K-major SWIZZLE_NONE canonical layout, packed on the host at model creation;
//   * the A operand [x_t | h_{t-1}] (and [x_t | r*h]) lives in TENSOR MEMORY (TS-form MMA): each thread writes
//     its own stream's row with tcgen05.st, so activations never touch shared memory and no proxy fence is
//     needed.  TMEM map (512 columns): D_gates 0..255 (r | u), D_cand 256..383, A_x, A_h;
//   * the input projection is not a separate GEMM: x_t W[0:in] is accumulated into the same TMEM tile as
//     h W[in:], and the candidate's x-part MMAs are issued right behind the gate MMAs so they run under
//     the gate epilogue;
//   * per step: 16+8+8 MMAs issued by one thread, two tcgen05.commit -> mbarrier hand-offs, two CTA barriers;
//     256 threads = two warpgroups that split the 128 hidden units; gate algebra in fp32 with ex2/rcp.
// Operands are fp16 (weights rounded once on the host, activations rounded when written to TMEM), accumulation
// and all state fp32.  The layer-0 input projection is the exception: mel features are unbounded (tens for loud
// audio) and a single fp16 rounding of x and W_x alone costs up to 3e-3 on the carried state, so that product is
// issued as three fp16 MMAs  x_hi*W_hi + x_lo*W_hi + x_hi*W_lo  (x = x_hi + x_lo, W = W_hi + W_lo; the
// dropped lo*lo term is 2^-22 relative), i.e. at fp32-grade accuracy for +6 small MMAs per step.  With it the
// deviation from the fp32 graph stays below 1e-3 (contract) on probabilities and state; see DESIGN.md.
#include <vector>

#include "common.cuh"
#include "tc05.cuh"

namespace kws {

constexpr int kTcTile = 128;
constexpr int kTcThreads = 256;
constexpr int kTcUnits = kHidden / 2;     // hidden units per thread (warpgroup split)

struct GruTcParams {
  int kx;                      // x width padded to a multiple of 16
  int kxw;                     // x columns of the packed weights: 2*kx when the x product is split (layer 0), else kx
  int in_dim;                  // true x width
  long S;
  int n;
  const float* x_f32;          // layer 0: [S, n, in_dim] fp32 (mel)
  const __half* x_f16;         // layer > 0: [S, n, 128] fp16
  __half* y_f16;               // non-last layers: [S, n, 128] fp16
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
};

__device__ __forceinline__ float fast_sigmoid(float x) {
  return __fdividef(1.0f, 1.0f + __expf(-x));          // ex2.approx + rcp.approx: ~1e-7 absolute
}
__device__ __forceinline__ float fast_tanh(float x) {
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}

// kFirst: layer 0 -- x is fp32 (mel) and its projection uses the 3-term split.
template <bool kLast, bool kFirst>
__global__ void __launch_bounds__(kTcThreads, 1)
gru_tc_kernel(const GruTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int ktot = p.kxw + kHidden;                                                  // K extent of the packed weights
  unsigned char* sW = smem;                                                         // [384, ktot] fp16
  float* sBias = reinterpret_cast<float*>(smem + static_cast<size_t>(384) * ktot * 2);   // [384]
  float* sFcw = sBias + 384;                                                        // [128][8]
  float* sFcb = sFcw + kHidden * 8;                                                 // [8]
  float* sXch = sFcb + 8;                                                           // [128][8] FC partials of warpgroup 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXch + kTcTile * 8);                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, row = tid & 127;

  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_fence_init();
  }
  {
    const int n16 = 384 * ktot * 2 / 16;
    const uint4* src = reinterpret_cast<const uint4*>(p.wpack);
    for (int i = tid; i < n16; i += kTcThreads) reinterpret_cast<uint4*>(sW)[i] = __ldg(src + i);
    for (int i = tid; i < 384; i += kTcThreads) sBias[i] = p.bias[i];
    if (kLast) {
      for (int i = tid; i < kHidden * 8; i += kTcThreads) {
        const int j = i >> 3, c = i & 7;
        sFcw[i] = c < p.C ? p.fc_w[j * p.C + c] : 0.0f;
      }
      if (tid < 8) sFcb[tid] = tid < p.C ? p.fc_b[tid] : 0.0f;
    }
  }
  tc::fence_proxy_async();            // weights written with generic stores, read by the MMA (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t colDg = 0, colDc = 256, colAx = 384;
  const uint32_t colAxl = colAx + p.kx / 2;                       // x_lo (kFirst only)
  const bool split = kFirst && p.kxw != p.kx;                     // 3-term x product (needs 2*kx/2 + 64 <= 128 columns)
  const uint32_t colAh = colAx + p.kxw / 2;
  const uint32_t my_ah = tmem + lane_sel + colAh + 32 * wg;      // this thread's 64 units = 32 columns
  const int xq = p.kx / 16;                                       // st4 groups of x per warpgroup
  const uint32_t my_ax = tmem + lane_sel + colAx + (p.kx / 4) * wg;
  const uint32_t sbo = static_cast<uint32_t>(ktot / 8) * 128;
  const uint32_t sW_addr = tc::smem_u32(sW);
  const uint32_t idesc_g = tc::idesc_f16(128, 256), idesc_c = tc::idesc_f16(128, 128);
  const long ntiles = (p.S + kTcTile - 1) / kTcTile;
  uint32_t phase = 0;                                             // parity of both mbarriers (one completion each per step)
  const int u0 = 64 * wg;                                         // first hidden unit of this thread

  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long s = tile * kTcTile + row;
    const bool ok = s < p.S;
    const int len = ok ? (p.seq_len ? p.seq_len[s] : p.n) : 0;
    float h[kTcUnits];
    {
      const bool zero = !ok || (p.zero_state && p.zero_state[s]);
      const float4* src = reinterpret_cast<const float4*>(p.h_in + (ok ? s : 0) * kHidden + u0);
#pragma unroll
      for (int i = 0; i < kTcUnits / 4; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!zero) v = src[i];
        h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
      }
    }
    // ---- A_h <- fp16(h);  A_x <- x_0
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = tc::pack_half2(h[32 * c + 2 * i], h[32 * c + 2 * i + 1]);
      tc::st16(my_ah + 16 * c, v);
    }
    uint32_t xr[32];                                             // x_t of this thread's half of the row, packed fp16 pairs
    uint32_t xl[kFirst ? 32 : 1];                                // residuals x - fp16(x) (kFirst only)
    auto load_x = [&](int t) {
      if (!kFirst) {
        const uint4* src = reinterpret_cast<const uint4*>(p.x_f16 + ((ok ? s : 0) * p.n + t) * static_cast<long>(kHidden) + u0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (ok) v = __ldg(src + q);
          xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
        }
      } else {
        const float* src = p.x_f32 + ((ok ? s : 0) * p.n + t) * static_cast<long>(p.in_dim);
        const int k0 = (p.kx / 2) * wg;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < xq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = k0 + 8 * q + 2 * i;
              const float a = (ok && k < p.in_dim) ? __ldg(src + k) : 0.0f;
              const float b = (ok && k + 1 < p.in_dim) ? __ldg(src + k + 1) : 0.0f;
              const __half2 hi = __floats2half2_rn(a, b);
              const float2 back = __half22float2(hi);
              xr[4 * q + i] = *reinterpret_cast<const uint32_t*>(&hi);
              xl[kFirst ? 4 * q + i : 0] = tc::pack_half2(a - back.x, b - back.y);
            }
          }
        }
      }
    };
    auto store_x = [&]() {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < xq) {
          const uint32_t v[4] = {xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]};
          tc::st4(my_ax + 4 * q, v);
          if (split) {
            const uint32_t w[4] = {xl[kFirst ? 4 * q : 0], xl[kFirst ? 4 * q + 1 : 0], xl[kFirst ? 4 * q + 2 : 0],
                                   xl[kFirst ? 4 * q + 3 : 0]};
            tc::st4(my_ax + p.kx / 2 + 4 * q, w);
          }
        }
    };
    if (p.n > 0) {
      load_x(0);
      store_x();
    }

    float fc_part[8];
    bool fc_pending = false;
    int fc_t = 0;
    auto fc_finish = [&](int t_done) {                          // warpgroup 0: combine, softmax, write
      if (wg == 0 && ok) {
        float lg[8];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          lg[c] = fc_part[c] + sXch[row * 8 + c] + sFcb[c];
          if (c < p.C) mx = fmaxf(mx, lg[c]);
        }
        float e[8], sum = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          e[c] = c < p.C ? expf(lg[c] - mx) : 0.0f;
          sum += e[c];
        }
        float* pr = p.probs + (s * p.n + t_done) * p.C;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < p.C) pr[c] = e[c] / sum;
        if (p.logits) {
          float* lo = p.logits + (s * p.n + t_done) * p.C;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < p.C) lo[c] = lg[c];
        }
      }
    };

    for (int t = 0; t < p.n; ++t) {
      // ---- A operand complete -> issue gate MMAs (+ the candidate's x part behind them)
      tc::wait_st();
      tc::fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        // B descriptor of K-chunk k16 of the packed weights; rows 0..255 = gates, 256..383 = candidate
        auto wdesc = [&](int k16, bool cand) {
          return tc::smem_desc(sW_addr + (cand ? 32 * sbo : 0) + 256 * k16, 128, sbo);
        };
        const int nx = p.kx / 16, nxw = p.kxw / 16;
        bool acc = false;
        for (int k16 = 0; k16 < nx; ++k16, acc = true)                       // x_hi * Wx_hi
          tc::mma_ts(tmem + colDg, tmem + colAx + 8 * k16, wdesc(k16, false), idesc_g, acc);
        if (split) {
          for (int k16 = 0; k16 < nx; ++k16)                                 // x_lo * Wx_hi
            tc::mma_ts(tmem + colDg, tmem + colAxl + 8 * k16, wdesc(k16, false), idesc_g, true);
          for (int k16 = 0; k16 < nx; ++k16)                                 // x_hi * Wx_lo
            tc::mma_ts(tmem + colDg, tmem + colAx + 8 * k16, wdesc(nx + k16, false), idesc_g, true);
        }
        for (int k16 = 0; k16 < kHidden / 16; ++k16, acc = true)             // h * Wh
          tc::mma_ts(tmem + colDg, tmem + colAh + 8 * k16, wdesc(nxw + k16, false), idesc_g, acc);
        tc::commit(&bars[0]);
        acc = false;
        for (int k16 = 0; k16 < nx; ++k16, acc = true)
          tc::mma_ts(tmem + colDc, tmem + colAx + 8 * k16, wdesc(k16, true), idesc_c, acc);
        if (split) {
          for (int k16 = 0; k16 < nx; ++k16)
            tc::mma_ts(tmem + colDc, tmem + colAxl + 8 * k16, wdesc(k16, true), idesc_c, true);
          for (int k16 = 0; k16 < nx; ++k16)
            tc::mma_ts(tmem + colDc, tmem + colAx + 8 * k16, wdesc(nx + k16, true), idesc_c, true);
        }
      }
      if (kLast && fc_pending) {                               // previous step's softmax, under the gate MMAs
        fc_finish(fc_t);
        fc_pending = false;
      }
      if (t + 1 < p.n) load_x(t + 1);                          // global loads in flight during the step
      tc::mbar_wait(&bars[0], phase);
      tc::fence_after_sync();

      // ---- epilogue 1: r, u;  A_h <- fp16(r * h)
      float u[kTcUnits];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t vr[16], vu[16];
        tc::ld16(tmem + lane_sel + colDg + u0 + 16 * c, vr);
        tc::ld16(tmem + lane_sel + colDg + kHidden + u0 + 16 * c, vu);
        tc::wait_ld();
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int j = 16 * c + i;
          const float r0 = fast_sigmoid(__uint_as_float(vr[i]) + sBias[u0 + j]);
          const float r1 = fast_sigmoid(__uint_as_float(vr[i + 1]) + sBias[u0 + j + 1]);
          u[j] = fast_sigmoid(__uint_as_float(vu[i]) + sBias[kHidden + u0 + j]);
          u[j + 1] = fast_sigmoid(__uint_as_float(vu[i + 1]) + sBias[kHidden + u0 + j + 1]);
          packed[i / 2] = tc::pack_half2(r0 * h[j], r1 * h[j + 1]);
        }
        tc::st8(my_ah + 8 * c, packed);
      }
      tc::wait_st();
      tc::fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        for (int k16 = 0; k16 < kHidden / 16; ++k16)
          tc::mma_ts(tmem + colDc, tmem + colAh + 8 * k16, tc::smem_desc(sW_addr + 32 * sbo + 256 * (p.kxw / 16 + k16), 128, sbo),
                     idesc_c, (p.kx > 0) || k16 > 0);
        tc::commit(&bars[1]);
      }
      tc::mbar_wait(&bars[1], phase);
      tc::fence_after_sync();
      phase ^= 1;

      // ---- epilogue 2: candidate, state update, outputs;  A_h <- fp16(h'),  A_x <- x_{t+1}
      const bool live = t < len;
#pragma unroll
      for (int c = 0; c < 8; ++c) fc_part[c] = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t vc[16];
        tc::ld16(tmem + lane_sel + colDc + u0 + 16 * c, vc);
        tc::wait_ld();
        uint32_t packed[8];
        uint32_t ypacked[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int j = 16 * c + i;
          const float c0 = fast_tanh(__uint_as_float(vc[i]) + sBias[2 * kHidden + u0 + j]);
          const float c1 = fast_tanh(__uint_as_float(vc[i + 1]) + sBias[2 * kHidden + u0 + j + 1]);
          const float n0 = u[j] * h[j] + (1.0f - u[j]) * c0;
          const float n1 = u[j + 1] * h[j + 1] + (1.0f - u[j + 1]) * c1;
          h[j] = live ? n0 : h[j];
          h[j + 1] = live ? n1 : h[j + 1];
          const float y0 = live ? n0 : 0.0f, y1 = live ? n1 : 0.0f;       // dynamic_rnn: zero output past the length
          packed[i / 2] = tc::pack_half2(h[j], h[j + 1]);
          if (kLast) {
            const float4 wa0 = *reinterpret_cast<const float4*>(sFcw + (u0 + j) * 8);
            const float4 wb0 = *reinterpret_cast<const float4*>(sFcw + (u0 + j) * 8 + 4);
            const float4 wa1 = *reinterpret_cast<const float4*>(sFcw + (u0 + j + 1) * 8);
            const float4 wb1 = *reinterpret_cast<const float4*>(sFcw + (u0 + j + 1) * 8 + 4);
            fc_part[0] = fmaf(y0, wa0.x, fc_part[0]); fc_part[1] = fmaf(y0, wa0.y, fc_part[1]);
            fc_part[2] = fmaf(y0, wa0.z, fc_part[2]); fc_part[3] = fmaf(y0, wa0.w, fc_part[3]);
            fc_part[4] = fmaf(y0, wb0.x, fc_part[4]); fc_part[5] = fmaf(y0, wb0.y, fc_part[5]);
            fc_part[6] = fmaf(y0, wb0.z, fc_part[6]); fc_part[7] = fmaf(y0, wb0.w, fc_part[7]);
            fc_part[0] = fmaf(y1, wa1.x, fc_part[0]); fc_part[1] = fmaf(y1, wa1.y, fc_part[1]);
            fc_part[2] = fmaf(y1, wa1.z, fc_part[2]); fc_part[3] = fmaf(y1, wa1.w, fc_part[3]);
            fc_part[4] = fmaf(y1, wb1.x, fc_part[4]); fc_part[5] = fmaf(y1, wb1.y, fc_part[5]);
            fc_part[6] = fmaf(y1, wb1.z, fc_part[6]); fc_part[7] = fmaf(y1, wb1.w, fc_part[7]);
          } else {
            ypacked[i / 2] = tc::pack_half2(y0, y1);
          }
        }
        tc::st8(my_ah + 8 * c, packed);
        if (!kLast && ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.y_f16 + (s * p.n + t) * static_cast<long>(kHidden) + u0 + 16 * c);
          dst[0] = make_uint4(ypacked[0], ypacked[1], ypacked[2], ypacked[3]);
          dst[1] = make_uint4(ypacked[4], ypacked[5], ypacked[6], ypacked[7]);
        }
      }
      if (t + 1 < p.n) store_x();
      if (kLast) {
        if (wg == 1) {
          *reinterpret_cast<float4*>(sXch + row * 8) = make_float4(fc_part[0], fc_part[1], fc_part[2], fc_part[3]);
          *reinterpret_cast<float4*>(sXch + row * 8 + 4) = make_float4(fc_part[4], fc_part[5], fc_part[6], fc_part[7]);
        }
        fc_pending = true;
        fc_t = t;
      }
    }
    if (kLast && fc_pending) {
      __syncthreads();
      fc_finish(fc_t);
      fc_pending = false;
    }
    if (ok) {
      float4* dst = reinterpret_cast<float4*>(p.h_out + s * kHidden + u0);
#pragma unroll
      for (int i = 0; i < kTcUnits / 4; ++i) dst[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    }
    // the next tile's first barrier orders these TMEM stores / smem reads against its MMAs
    __syncthreads();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static size_t gru_tc_smem_bytes(int ktot) {
  return static_cast<size_t>(384) * ktot * 2 + sizeof(float) * (384 + kHidden * 8 + 8 + kTcTile * 8) + 2 * sizeof(uint64_t) + 16;
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
  const int L = m->cfg.num_layers;
  if (a.n <= 0) {
    if (a.state_out != a.state_in)
      KWS_CUDA_OK(cudaMemcpyAsync(a.state_out, a.state_in, sizeof(float) * L * a.S * kHidden, cudaMemcpyDeviceToDevice, st));
    return KWS_OK;
  }
  if (L > 1) {
    const int rc = kws_model_reserve(m, a.S, a.n);
    if (rc != KWS_OK) return rc;
  }
  const long ntiles = ceil_div(a.S, kTcTile);
  const size_t per_buf = static_cast<size_t>(a.S) * a.n * kHidden;           // halves
  __half* seq = reinterpret_cast<__half*>(m->scratch_seq);
  for (int l = 0; l < L; ++l) {
    const bool last = l == L - 1;
    GruTcParams p;
    p.kx = m->layer[l].tc_kx;
    p.kxw = m->layer[l].tc_kxw;
    p.in_dim = m->layer[l].in_dim;
    p.S = a.S;
    p.n = a.n;
    p.x_f32 = l == 0 ? a.x : nullptr;
    p.x_f16 = l == 0 ? nullptr : seq + ((l - 1) & 1) * per_buf;
    p.y_f16 = last ? nullptr : seq + (l & 1) * per_buf;
    p.wpack = static_cast<const __half*>(m->layer[l].tc_wpack);
    p.bias = m->layer[l].tc_bias;
    p.h_in = a.state_in + static_cast<long>(l) * a.S * kHidden;
    p.h_out = a.state_out + static_cast<long>(l) * a.S * kHidden;
    p.seq_len = a.seq_len;
    p.zero_state = a.zero_state;
    p.fc_w = m->fc_w;
    p.fc_b = m->fc_b;
    p.C = m->cfg.num_classes;
    p.probs = a.probs;
    p.logits = a.logits;
    const size_t smem = gru_tc_smem_bytes(p.kxw + kHidden);
    const long blocks = ntiles < sm_count() ? ntiles : sm_count();
    const bool first = l == 0;
    auto launch = [&](auto kernel) -> int {
      KWS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kernel<<<static_cast<unsigned>(blocks), kTcThreads, smem, st>>>(p);
      return KWS_OK;
    };
    int rc;
    if (last) rc = first ? launch(gru_tc_kernel<true, true>) : launch(gru_tc_kernel<true, false>);
    else rc = first ? launch(gru_tc_kernel<false, true>) : launch(gru_tc_kernel<false, false>);
    if (rc != KWS_OK) return rc;
    KWS_LAUNCH_OK("gru_tc_kernel");
  }
  return KWS_OK;
}

}  // namespace kws
