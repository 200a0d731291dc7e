// K1 -- fused framing + 400-point real DFT magnitude + mel projection (+ the server's per-chunk pre-step).
//
// Replaces, in one pass over the PCM and without materialising frames or spectra:
//   tf_frame(signal, 400, 160)            utils/stft.py:27-81   (rect window, no padding)
//   tf.abs(tf.spectral.rfft(frames,[400])) models/rnn_ctc.py:137
//   tf.matmul(linearspec, mel_basis)       models/rnn_ctc.py:139-149
// and, when launched by the streaming server for a chunk that fits one work item, also
//   vad(data, 30)                          utils/basic_vad.py:17-18, detector.py:168
//   tail carry (len-400)%160+240           detector.py:179-183
// so the chunk's PCM is read from HBM exactly once.
//
// Work item = (stream, group of 16 consecutive frame PAIRS = 32 frames) handled by one 320-thread CTA:
// thread t = 20*pair + column owns one DFT-20 column of its pair in both stages of the packed-real 20x20 FFT
// (fft400.cuh).  The item's PCM (5360 samples) is staged once in shared memory as fp32 with coalesced 16-byte
// loads; every 320-sample block is skewed by 20 floats so that the 32 lanes of a warp -- which straddle two
// pairs -- read 32 distinct banks.  int16 samples are transformed unscaled (the DFT is linear) and 2^-15 is
// folded into the magnitude.  The mel projection uses the per-band non-zero bin ranges of the basis found at
// model creation, so it costs ~2*201 FMAs per frame for a triangular filterbank yet stays exact for any dense
// basis.  Persistent grid: (resident CTAs per SM) x 148 SMs, grid-stride over work items.
#include "common.cuh"
#include "fft400.cuh"

namespace kws {

using fft::cpx;

constexpr int kFePairs = 16;                         // frame pairs per work item
constexpr int kFeThreads = kFePairs * fft::kR;       // 320
constexpr int kFeItemFrames = 2 * kFePairs;          // 32
constexpr int kFeBlock = 2 * kHop;                   // 320 samples between consecutive pairs
constexpr int kFeSkew = 20;                          // floats of skew per 320-sample block
constexpr int kFeWinSamples = kFePairs * kFeBlock + (kFft - kHop);       // 5360
constexpr int kFeWinFloats = kFeWinSamples + kFeSkew * (kFePairs + 1);   // 5700
constexpr int kFeItemHop = kFePairs * kFeBlock;      // 5120 samples between consecutive items of a stream

struct FrontendParams {
  PcmSource src;
  long S;
  int max_frames;           // row stride of mel_out in frames
  int groups;               // work items per stream = ceil(max_frames / 32)
  const int* nframes;       // [S] or null -> frames from the signal length
  int n_mel;
  const cpx* twiddle;       // [400] k2-major (fft::twt_index)
  const int* mel_start;
  const int* mel_count;
  const int* mel_offset;
  const float* mel_weight;
  int mel_nnz;
  float* mel_out;           // [S, max_frames, n_mel]
  float mag_scale;          // 0.5 * (int16 input ? 2^-15 : 1)
  // fused server pre-step (only with groups == 1 and int16 input): VAD, frame count, next tail
  int fuse_pre;
  long long vad_limit;
  int16_t* tail_next;       // [S, 400]
  int* len_next;            // [S]
  unsigned char* silence;   // [S]
  int* nframes_out;         // [S]
};

__device__ __forceinline__ int win_pos(int i) { return i + kFeSkew * (i / kFeBlock); }

__global__ void __launch_bounds__(kFeThreads, 2)
frontend_kernel(const FrontendParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* twt = reinterpret_cast<cpx*>(smem_raw);                      // [400]
  cpx* buf = twt + fft::kN;                                         // [16][420]
  float* win = reinterpret_cast<float*>(buf + kFePairs * fft::kBufSlots);   // [5700]
  float* mel_w = win + kFeWinFloats;                                // [nnz]
  int* mel_start = reinterpret_cast<int*>(mel_w + p.mel_nnz);       // [M]
  int* mel_count = mel_start + p.n_mel;
  int* mel_off = mel_count + p.n_mel;
  int* red = mel_off + p.n_mel;                                     // [16] block reduction scratch

  const int tid = threadIdx.x;
  for (int i = tid; i < fft::kN; i += kFeThreads) twt[i] = p.twiddle[i];
  for (int i = tid; i < p.mel_nnz; i += kFeThreads) mel_w[i] = p.mel_weight[i];
  for (int i = tid; i < p.n_mel; i += kFeThreads) {
    mel_start[i] = p.mel_start[i];
    mel_count[i] = p.mel_count[i];
    mel_off[i] = p.mel_offset[i];
  }

  const int pair = tid / fft::kR;
  const int col = tid - pair * fft::kR;
  cpx* my_buf = buf + pair * fft::kBufSlots;
  const float* my_win = win + pair * (kFeBlock + kFeSkew) + col;
  const long items = p.S * p.groups;
  const bool i16 = p.src.body_dtype == KWS_PCM_I16;
  __syncthreads();                                    // tables staged
  // Barriers per item: window staged | Y' exchanged | Z mirror rows published | magnitudes published.  The
  // next item's staging only writes `win` (last read before the second barrier) and its stage 1 writes `buf`
  // after its own first barrier, so no barrier is needed at the loop boundary.

  for (long item = blockIdx.x; item < items; item += gridDim.x) {
    const long s = item / p.groups;
    const int g = static_cast<int>(item - s * p.groups);
    const int head_len = p.src.head_len ? p.src.head_len[s] : 0;
    const int total_len = head_len + p.src.body_len;
    const int nfr_sig = total_len >= kFft ? 1 + (total_len - kFft) / kHop : 0;
    int nfr = p.fuse_pre ? nfr_sig : (p.nframes ? p.nframes[s] : nfr_sig);
    if (nfr > p.max_frames) nfr = p.max_frames;
    const int q0 = g * kFeItemHop;                    // stream sample held at window position 0
    const int f0 = g * kFeItemFrames;

    // ---- stage the window: stream samples [q0, q0 + 5360), zero beyond the signal
    int vad_acc = 0;
    {
      // carried tail (int16), group 0 only
      for (int q = q0 + tid; q < head_len && q < q0 + kFeWinSamples; q += kFeThreads)
        win[win_pos(q - q0)] = static_cast<float>(p.src.head[s * p.src.ld_head + q]);
      int b_lo = q0 - head_len;
      if (b_lo < 0) b_lo = 0;
      int b_hi = q0 + kFeWinSamples - head_len;
      if (b_hi > p.src.body_len) b_hi = p.src.body_len;
      if (i16) {
        const int16_t* row = static_cast<const int16_t*>(p.src.body) + s * p.src.ld_body;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
        if (vec_ok) {
          for (int j = (b_lo >> 3) + tid; 8 * j < b_hi; j += kFeThreads) {
            const int b0 = 8 * j;
            int x[8];
            if (b0 + 8 <= p.src.body_len) {
              const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + b0));
              x[0] = static_cast<short>(v.x & 0xffffu); x[1] = static_cast<int>(v.x) >> 16;
              x[2] = static_cast<short>(v.y & 0xffffu); x[3] = static_cast<int>(v.y) >> 16;
              x[4] = static_cast<short>(v.z & 0xffffu); x[5] = static_cast<int>(v.z) >> 16;
              x[6] = static_cast<short>(v.w & 0xffffu); x[7] = static_cast<int>(v.w) >> 16;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = b0 + e < p.src.body_len ? row[b0 + e] : 0;
            }
            int i = head_len + b0 - q0;               // window position of element 0 (may be < 0)
            int r = (i + kFeBlock) % kFeBlock;        // i >= -7 here
            int pos = i + kFeSkew * ((i + kFeBlock) / kFeBlock - 1);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (r == kFeBlock) {
                r = 0;
                pos += kFeSkew;
              }
              const int b = b0 + e;
              if (b >= b_lo && b < b_hi) win[pos] = static_cast<float>(x[e]);
              vad_acc += abs(x[e]);
              ++pos;
              ++r;
            }
          }
        } else {
          for (int b = b_lo + tid; b < b_hi; b += kFeThreads) {
            const int x = row[b];
            win[win_pos(head_len + b - q0)] = static_cast<float>(x);
            vad_acc += abs(x);
          }
        }
      } else {
        const float* row = static_cast<const float*>(p.src.body) + s * p.src.ld_body;
        for (int b = b_lo + tid; b < b_hi; b += kFeThreads) win[win_pos(head_len + b - q0)] = __ldg(row + b);
      }
      // zero fill beyond the signal
      int z_lo = total_len - q0;
      if (z_lo < 0) z_lo = 0;
      for (int i = z_lo + tid; i < kFeWinSamples; i += kFeThreads) win[win_pos(i)] = 0.0f;
    }
    if (p.fuse_pre) {                                 // block sum of |x| over the chunk (exact integers)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vad_acc += __shfl_xor_sync(0xffffffffu, vad_acc, o);
      if ((tid & 31) == 0) red[tid >> 5] = vad_acc;
    }
    __syncthreads();

    if (p.fuse_pre) {
      // keep = (len-400)%160+240 last samples (detector.py:181-182); everything when no frame fits yet
      const int keep = total_len >= kFft ? (total_len - kFft) % kHop + (kFft - kHop) : total_len;
      const int start = total_len - keep;
      int16_t* tnext = p.tail_next + s * 400;
      for (int i = tid; i < keep; i += kFeThreads) tnext[i] = static_cast<int16_t>(win[win_pos(start + i)]);
      if (tid == 0) {
        long long sum = 0;
#pragma unroll
        for (int w = 0; w < kFeThreads / 32; ++w) sum += red[w];
        p.silence[s] = sum > p.vad_limit ? 0 : 1;
        p.nframes_out[s] = nfr;
        p.len_next[s] = keep;
      }
    }

    const int fa = f0 + 2 * pair;
    const bool live = fa < nfr;
    cpx v[20];
    if (live) {
#pragma unroll
      for (int n2 = 0; n2 < 20; ++n2) {
        v[n2].re = my_win[20 * n2 + (n2 >= 16 ? kFeSkew : 0)];
        v[n2].im = my_win[kHop + 20 * n2 + (n2 >= 8 ? kFeSkew : 0)];
      }
      fft::stage1_col(col, v, twt, my_buf);
    }
    __syncthreads();
    if (live) fft::stage2_col(col, my_buf, v);
    __syncthreads();
    float* mag = reinterpret_cast<float*>(my_buf);
    if (live) fft::untangle_col(col, v, my_buf, p.mag_scale, mag);
    __syncthreads();
    if (live) {
      const bool b_live = fa + 1 < nfr;
      float* out_a = p.mel_out + (s * p.max_frames + fa) * p.n_mel;
      for (int m = col; m < p.n_mel; m += fft::kR) {
        const int k0 = mel_start[m], c = mel_count[m];
        const float* wv = mel_w + mel_off[m];
        float acc_a = 0.0f, acc_b = 0.0f;
        for (int i = 0; i < c; ++i) {
          const float wgt = wv[i];
          acc_a = fmaf(mag[k0 + i], wgt, acc_a);
          acc_b = fmaf(mag[fft::kMagB + k0 + i], wgt, acc_b);
        }
        out_a[m] = acc_a;
        if (b_live) out_a[p.n_mel + m] = acc_b;
      }
    }
  }
}

static size_t frontend_smem_bytes(const kws_model* m) {
  return sizeof(cpx) * fft::kN + sizeof(cpx) * kFePairs * fft::kBufSlots + sizeof(float) * kFeWinFloats +
         sizeof(float) * m->mel.nnz + sizeof(int) * 3 * m->cfg.n_mel + sizeof(int) * 16;
}

bool frontend_can_fuse_pre(int chunk_len, int tail_cap) {
  return chunk_len + tail_cap - 1 <= kFeWinSamples;
}

int launch_frontend(const kws_model* m, const PcmSource& src, int64_t S, int32_t max_frames,
                    const int32_t* nframes, float* mel_out, cudaStream_t st, const FrontendPre* pre) {
  if (S <= 0) return KWS_OK;
  if (max_frames <= 0 && !pre) return KWS_OK;
  FrontendParams p;
  p.src = src;
  p.S = S;
  p.max_frames = max_frames;
  p.groups = static_cast<int>(ceil_div(max_frames > 0 ? max_frames : 1, kFeItemFrames));
  p.nframes = nframes;
  p.n_mel = m->cfg.n_mel;
  p.twiddle = reinterpret_cast<const cpx*>(m->twiddle400);
  p.mel_start = m->mel.start;
  p.mel_count = m->mel.count;
  p.mel_offset = m->mel.offset;
  p.mel_weight = m->mel.weight;
  p.mel_nnz = m->mel.nnz;
  p.mel_out = mel_out;
  p.mag_scale = src.body_dtype == KWS_PCM_I16 ? 0.5f / 32768.0f : 0.5f;
  p.fuse_pre = 0;
  p.vad_limit = 0;
  p.tail_next = nullptr;
  p.len_next = nullptr;
  p.silence = nullptr;
  p.nframes_out = nullptr;
  if (pre) {
    if (p.groups != 1 || src.body_dtype != KWS_PCM_I16)
      return fail(KWS_ERR_INVALID_ARGUMENT, "fused pre-step needs an int16 chunk that fits one work item");
    p.fuse_pre = 1;
    p.vad_limit = pre->vad_limit;
    p.tail_next = pre->tail_next;
    p.len_next = pre->len_next;
    p.silence = pre->silence;
    p.nframes_out = pre->nframes_out;
  }
  const size_t smem = frontend_smem_bytes(m);
  static thread_local size_t configured = 0;
  if (configured < smem) {
    KWS_CUDA_OK(cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    configured = smem;
  }
  int per_sm = 0;
  KWS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frontend_kernel, kFeThreads, smem));
  if (per_sm < 1) return fail(KWS_ERR_CUDA, "frontend_kernel does not fit on an SM (smem %zu)", smem);
  const long items = S * p.groups;
  long blocks = items;
  const long resident = static_cast<long>(per_sm) * sm_count();
  if (blocks > resident) blocks = resident;
  frontend_kernel<<<static_cast<unsigned>(blocks), kFeThreads, smem, st>>>(p);
  KWS_LAUNCH_OK("frontend_kernel");
  return KWS_OK;
}

}  // namespace kws
