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
// folded into the magnitude.  Magnitudes go to shared memory transposed ([bin][frame]) so that the mel
// projection runs with lane = frame and warp-uniform bands: it uses the per-band non-zero bin ranges of the basis
// found at model creation (~2*201 FMAs per frame for a triangular filterbank, exact for any dense basis) with
// broadcast weight reads and conflict-free magnitude reads; the [32, M] result tile leaves as one contiguous
// block.  Persistent grid: (resident CTAs per SM) x 148 SMs, grid-stride over work items.
#include "common.cuh"
#include "fft400.cuh"

namespace kws {

using fft::cpx;

constexpr int kFePairs = 16;                         // frame pairs per work item
constexpr int kFeThreads = kFePairs * fft::kR;       // 320
constexpr int kFeWarps = kFeThreads / 32;            // 10
constexpr int kFeItemFrames = 2 * kFePairs;          // 32
constexpr int kFeBlock = 2 * kHop;                   // 320 samples between consecutive pairs
constexpr int kFeSkew = 20;                          // floats of skew per 320-sample block
constexpr int kFeWinSamples = kFePairs * kFeBlock + (kFft - kHop);       // 5360
constexpr int kFeWinRounds = (kFeWinSamples + kFeThreads - 1) / kFeThreads;   // 17
constexpr int kFeItemHop = kFePairs * kFeBlock;      // 5120 samples between consecutive items of a stream
constexpr int kFeMagStride = 33;                     // floats per bin row of the transposed magnitudes [201][33]
// the window (5360 samples + skew) and the transposed magnitudes share one region: the window is dead after stage 1
constexpr int kFeWinFloats = kFeWinRounds * (kFeBlock + kFeSkew);        // 5780 (position tid + 340*round)
constexpr int kFeMagRows = fft::kBins + 3;           // band ranges are padded to multiples of 4 bins: rows 201..203 stay zero
constexpr int kFeRegionFloats = kFeMagRows * kFeMagStride > kFeWinFloats ? kFeMagRows * kFeMagStride : kFeWinFloats;
static_assert(kFeWinFloats <= fft::kBins * kFeMagStride, "the window must not reach the zero rows");
static_assert((kFeRegionFloats * 4) % 16 == 0, "mel weights follow the region and are read as float4");

struct FrontendParams {
  PcmSource src;
  long S;
  int max_frames;           // row stride of mel_out in frames
  int groups;               // work items per stream = ceil(max_frames / 32)
  const int* nframes;       // [S] or null -> frames from the signal length
  int n_mel;
  const cpx* twiddle;       // [20*52] periodic k2-major table (fft::kTwSlots)
  const int* mel_start;
  const int* mel_count;
  const int* mel_offset;
  const float* mel_weight;
  int mel_nnz;
  float* mel_out;           // [S, max_frames, n_mel], or stream-tiled (common.cuh) when tiled_out
  int tiled_out;
  float mag_scale;          // 0.5 * (int16 input ? 2^-15 : 1)
  // fused server pre-step (only with groups == 1 and int16 input): VAD, frame count, next tail
  int fuse_pre;
  long long vad_limit;
  int16_t* tail_next;       // [S, 400]
  int* len_next;            // [S]
  unsigned char* silence;   // [S]
  int* nframes_out;         // [S]
};

// Magnitude column ("frame slot") of frame fr of pair p in the transposed [bin][33] array.  Chosen so that the
// 32 lanes of a warp in the untangle phase (consecutive columns of two adjacent pairs) hit 32 distinct banks:
// consecutive pairs are 20 slots apart mod 32; the 2 x 16 frames still fill the 32 slots exactly once.
__device__ __forceinline__ int mag_slot(int pair, int fr) { return ((20 * pair) & 31) + 2 * (pair >> 3) + fr; }
// inverse: slot -> frame index 2*pair + fr within the item  (5*5 = 25 = 1 mod 8)
__device__ __forceinline__ int slot_frame(int slot) {
  const int pair = ((5 * (slot >> 2)) & 7) + 8 * ((slot >> 1) & 1);
  return 2 * pair + (slot & 1);
}

__global__ void __launch_bounds__(kFeThreads, 2)
frontend_kernel(const FrontendParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* twp = reinterpret_cast<cpx*>(smem_raw);                      // [1040]
  cpx* buf = twp + fft::kTwSlots;                                   // [16][420]; later the [32][M] output tile
  float* win = reinterpret_cast<float*>(buf + kFePairs * fft::kBufSlots);   // window, later magnitudes [201][33]
  float* mel_w = win + kFeRegionFloats;                             // [nnz]
  int* mel_start = reinterpret_cast<int*>(mel_w + p.mel_nnz);       // [M]
  int* mel_count = mel_start + p.n_mel;
  int* mel_off = mel_count + p.n_mel;
  int* red = mel_off + p.n_mel;                                     // [16] block reduction scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < fft::kTwSlots; i += kFeThreads) twp[i] = p.twiddle[i];
  for (int i = tid; i < p.mel_nnz; i += kFeThreads) mel_w[i] = p.mel_weight[i];
  for (int i = tid; i < p.n_mel; i += kFeThreads) {
    mel_start[i] = p.mel_start[i];
    mel_count[i] = p.mel_count[i];
    mel_off[i] = p.mel_offset[i];
  }
  for (int i = fft::kBins * kFeMagStride + tid; i < kFeRegionFloats; i += kFeThreads) win[i] = 0.0f;

  const int pair = tid / fft::kR;
  const int col = tid - pair * fft::kR;
  cpx* my_buf = buf + pair * fft::kBufSlots;
  const float* my_win = win + pair * (kFeBlock + kFeSkew) + col;
  const cpx* my_tw = twp + fft::tw_base(warp, lane);
  float* out_tile = reinterpret_cast<float*>(buf);
  const long items = p.S * p.groups;
  const bool i16 = p.src.body_dtype == KWS_PCM_I16;
  const int M = p.n_mel;
  __syncthreads();                                    // tables staged
  // Barriers per item: window staged | Y' exchanged | Z mirror rows published | magnitudes published | mel tile
  // complete.  The next item's staging writes the window/magnitude region (last read before the fifth barrier) and
  // its stage 1 writes `buf` after its own first barrier, so no barrier is needed at the loop boundary.

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

    // ---- stage the window: stream samples [q0, q0 + 5360) -> win[i + 20*(i/320)], zero beyond the signal.
    // Thread t takes samples t + 320*r, so the skewed position is simply t + 340*r; the loads are unit-stride
    // across the warp and all 17 of a thread are independent (one round trip to HBM).
    int vad_acc = 0;
    const bool last_ok = tid < kFeWinSamples - kFeBlock * (kFeWinRounds - 1);   // the 17th round is partial
    if (i16) {
      const int16_t* body = static_cast<const int16_t*>(p.src.body) + s * p.src.ld_body - head_len + q0 + tid;
      const int16_t* head = p.src.head + s * p.src.ld_head + q0 + tid;          // only dereferenced below head_len
      // generic element: carried tail, chunk, or zero beyond the signal
      auto fetch = [&](int r) {
        const int q = q0 + tid + kFeBlock * r;
        const bool in_head = q < head_len;
        const int16_t* ptr = in_head ? head + kFeBlock * r : body + kFeBlock * r;
        int x = 0;
        if (q < total_len) x = *ptr;
        return x;
      };
      // rounds 2..14 hold samples [q0+640, q0+4800): entirely inside the chunk for every steady-state item
      const bool fast = q0 + 2 * kFeBlock >= head_len && q0 + 15 * kFeBlock <= total_len;
      if (fast) {
        int x[kFeWinRounds];
        x[0] = fetch(0);
        x[1] = fetch(1);
#pragma unroll
        for (int r = 2; r < 15; ++r) x[r] = __ldg(body + kFeBlock * r);
        x[15] = fetch(15);
        x[16] = last_ok ? fetch(16) : 0;
#pragma unroll
        for (int r = 0; r < kFeWinRounds; ++r) {
          if (r < kFeWinRounds - 1 || last_ok) win[tid + (kFeBlock + kFeSkew) * r] = static_cast<float>(x[r]);
          if (r >= 2 || q0 + tid + kFeBlock * r >= head_len) vad_acc += abs(x[r]);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < kFeWinRounds; ++r) {
          const int x = (r < kFeWinRounds - 1 || last_ok) ? fetch(r) : 0;
          if (r < kFeWinRounds - 1 || last_ok) win[tid + (kFeBlock + kFeSkew) * r] = static_cast<float>(x);
          if (q0 + tid + kFeBlock * r >= head_len) vad_acc += abs(x);
        }
      }
    } else {
      const float* body = static_cast<const float*>(p.src.body) + s * p.src.ld_body + q0 + tid;
#pragma unroll
      for (int r = 0; r < kFeWinRounds; ++r) {
        const int q = q0 + tid + kFeBlock * r;
        const float v = q < total_len ? __ldg(body + kFeBlock * r) : 0.0f;
        if (r < kFeWinRounds - 1 || last_ok) win[tid + (kFeBlock + kFeSkew) * r] = v;
      }
    }
    if (p.fuse_pre) {                                 // block sum of |x| over the chunk (exact integers)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vad_acc += __shfl_xor_sync(0xffffffffu, vad_acc, o);
      if (lane == 0) red[warp] = vad_acc;
    }
    __syncthreads();

    if (p.fuse_pre) {
      // keep = (len-400)%160+240 last samples (detector.py:181-182); everything when no frame fits yet
      const int keep = total_len >= kFft ? (total_len - kFft) % kHop + (kFft - kHop) : total_len;
      const int start = total_len - keep;
      int16_t* tnext = p.tail_next + s * 400;
      for (int i = tid; i < keep; i += kFeThreads) {
        const int w = start + i;
        tnext[i] = static_cast<int16_t>(win[w + kFeSkew * (w / kFeBlock)]);
      }
      if (tid == 0) {
        long long sum = 0;
#pragma unroll
        for (int w = 0; w < kFeWarps; ++w) sum += red[w];
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
      fft::stage1_col(col, v, my_tw, my_buf);
    }
    __syncthreads();
    if (live) fft::stage2_col(col, my_buf, v);
    __syncthreads();
    // magnitudes, transposed: mag[bin*33 + slot]; the window is dead since the second barrier
    float* mag = win;
    if (live) fft::untangle_col(col, v, my_buf, p.mag_scale, mag + mag_slot(pair, 0), mag + mag_slot(pair, 1), kFeMagStride);
    __syncthreads();
    // ---- mel projection: lane = frame slot, warp = set of bands {warp, warp+10, ...}: band ranges and weights
    // are warp-uniform (broadcast reads, no divergence), magnitude reads are unit-stride across the lanes
    {
      const int fi = slot_frame(lane);
      const float* mcol = mag + lane;
      for (int m = warp; m < M; m += kFeWarps) {
        const int k0 = mel_start[m], c = mel_count[m];
        const float* wv = mel_w + mel_off[m];
        const float* mp = mcol + k0 * kFeMagStride;
        // ranges are padded to a multiple of 4 bins with zero weights (api.cu); the rows past bin 200 are zero
        float acc0 = 0.0f, acc1 = 0.0f;
        for (int i = 0; i < c; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wv + i);
          acc0 = fmaf(mp[0], w4.x, acc0);
          acc1 = fmaf(mp[kFeMagStride], w4.y, acc1);
          acc0 = fmaf(mp[2 * kFeMagStride], w4.z, acc0);
          acc1 = fmaf(mp[3 * kFeMagStride], w4.w, acc1);
          mp += 4 * kFeMagStride;
        }
        const float acc = acc0 + acc1;
        out_tile[fi * M + m] = acc;
      }
    }
    __syncthreads();
    // ---- the item's frames are one contiguous block of mel_out
    {
      int nfi = nfr - f0;
      if (nfi > kFeItemFrames) nfi = kFeItemFrames;
      const int total = nfi > 0 ? nfi * M : 0;
      if (p.tiled_out) {
        // float4 chunk c of frame f -> ((tile*n + f)*Q + c)*128 + stream%128
        const int Q = M >> 2;
        const float4* src4 = reinterpret_cast<const float4*>(out_tile);
        float4* dst4 = reinterpret_cast<float4*>(p.mel_out) + ((s >> 7) * p.max_frames + f0) * static_cast<long>(Q) * 128 + (s & 127);
        for (int i = tid; i < (total >> 2); i += kFeThreads) dst4[static_cast<long>(i) * 128] = src4[i];
      } else {
        float* dst = p.mel_out + (s * p.max_frames + f0) * M;
        if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mel_out) & 15) == 0) {
          const float4* src4 = reinterpret_cast<const float4*>(out_tile);
          float4* dst4 = reinterpret_cast<float4*>(dst);
          for (int i = tid; i < (total >> 2); i += kFeThreads) dst4[i] = src4[i];
        } else {
          for (int i = tid; i < total; i += kFeThreads) dst[i] = out_tile[i];
        }
      }
    }
  }
}

static size_t frontend_smem_bytes(const kws_model* m) {
  return sizeof(cpx) * fft::kTwSlots + sizeof(cpx) * kFePairs * fft::kBufSlots + sizeof(float) * kFeRegionFloats +
         sizeof(float) * m->mel.nnz + sizeof(int) * 3 * m->cfg.n_mel + sizeof(int) * 16;
}

bool frontend_can_fuse_pre(int chunk_len, int tail_cap) {
  return chunk_len + tail_cap - 1 <= kFeWinSamples;
}

int launch_frontend(const kws_model* m, const PcmSource& src, int64_t S, int32_t max_frames,
                    const int32_t* nframes, float* mel_out, cudaStream_t st, const FrontendPre* pre, bool tiled_out) {
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
  p.tiled_out = tiled_out ? 1 : 0;
  if (tiled_out && (m->cfg.n_mel & 3)) return fail(KWS_ERR_INVALID_ARGUMENT, "tiled mel output needs n_mel % 4 == 0");
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
