// K1 -- fused framing + 400-point real DFT magnitude + mel projection.
//
// Replaces, in one pass over the PCM and without materialising frames or spectra:
//   tf_frame(signal, 400, 160)            utils/stft.py:27-81   (rect window, no padding)
//   tf.abs(tf.spectral.rfft(frames,[400])) models/rnn_ctc.py:137
//   tf.matmul(linearspec, mel_basis)       models/rnn_ctc.py:139-149
//
// Work item = (stream, group of 6 consecutive frames) handled by one warp: three
// 10-thread teams each transform one frame PAIR with the packed-real 20x20 FFT of
// fft400.cuh.  The PCM window of each pair (560 samples) is staged in shared memory
// with coalesced loads; region strides are skewed (587 floats) so the three teams'
// strided reads fall on disjoint banks.  The mel projection uses the per-band
// non-zero bin ranges of the basis found at model creation, so it costs ~2*201 FMAs
// per frame for a triangular filterbank yet stays exact for any dense basis.
// Persistent grid: (resident CTAs per SM) x 148 SMs, grid-stride over work items.
#include "common.cuh"
#include "fft400.cuh"

namespace kws {

using fft::cpx;

constexpr int kFeWarps = 4;            // warps per CTA
constexpr int kFePairs = 3;            // frame pairs per warp per item
constexpr int kFeFramesPerItem = 2 * kFePairs;
constexpr int kFeWinStride = 587;      // 560 + 27: team bases land on banks 0, 11, 22

struct FrontendParams {
  PcmSource src;
  long S;
  int max_frames;           // row stride of mel_out in frames
  int groups;               // work items per stream = ceil(max_frames / 6)
  const int* nframes;       // [S] or null -> frames from the signal length
  int n_mel;
  const cpx* twiddle;       // [400]
  const int* mel_start;
  const int* mel_count;
  const int* mel_offset;
  const float* mel_weight;
  int mel_nnz;
  float* mel_out;           // [S, max_frames, n_mel]
};

__device__ __forceinline__ float load_sample(const PcmSource& src, long s, int i, int head_len) {
  if (i < head_len) return static_cast<float>(src.head[s * src.ld_head + i]) * (1.0f / 32768.0f);
  const long o = s * src.ld_body + (i - head_len);
  if (src.body_dtype == KWS_PCM_I16)
    return static_cast<float>(static_cast<const int16_t*>(src.body)[o]) * (1.0f / 32768.0f);
  return static_cast<const float*>(src.body)[o];
}

__global__ void __launch_bounds__(kFeWarps * 32)
frontend_kernel(const FrontendParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx* tw = reinterpret_cast<cpx*>(smem_raw);                       // [400]
  float* mel_w = reinterpret_cast<float*>(tw + fft::kN);            // [nnz]
  int* mel_start = reinterpret_cast<int*>(mel_w + p.mel_nnz);       // [M]
  int* mel_count = mel_start + p.n_mel;
  int* mel_off = mel_count + p.n_mel;
  // per-warp regions (8-byte aligned: everything before is a multiple of 4 bytes, pad to 8)
  size_t head_bytes = sizeof(cpx) * fft::kN + sizeof(float) * p.mel_nnz + sizeof(int) * 3 * p.n_mel;
  head_bytes = (head_bytes + 15) & ~static_cast<size_t>(15);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr size_t kWarpBytes = sizeof(cpx) * kFePairs * fft::kBufSlots + sizeof(float) * kFePairs * kFeWinStride;
  unsigned char* wbase = smem_raw + head_bytes + static_cast<size_t>(warp) * ((kWarpBytes + 15) & ~static_cast<size_t>(15));
  cpx* buf = reinterpret_cast<cpx*>(wbase);                                       // [3][420]
  float* win = reinterpret_cast<float*>(wbase + sizeof(cpx) * kFePairs * fft::kBufSlots);  // [3][587]

  for (int i = threadIdx.x; i < fft::kN; i += blockDim.x) tw[i] = p.twiddle[i];
  for (int i = threadIdx.x; i < p.mel_nnz; i += blockDim.x) mel_w[i] = p.mel_weight[i];
  for (int i = threadIdx.x; i < p.n_mel; i += blockDim.x) {
    mel_start[i] = p.mel_start[i];
    mel_count[i] = p.mel_count[i];
    mel_off[i] = p.mel_offset[i];
  }
  __syncthreads();

  const int team = lane / fft::kThreads;        // 0..2 active, 3 = two idle lanes
  const int j = lane - team * fft::kThreads;
  const bool active = team < kFePairs;
  const long items = p.S * p.groups;

  for (long item = blockIdx.x * static_cast<long>(kFeWarps) + warp; item < items;
       item += static_cast<long>(gridDim.x) * kFeWarps) {
    const long s = item / p.groups;
    const int g = static_cast<int>(item - s * p.groups);
    const int f0 = g * kFeFramesPerItem;
    const int head_len = p.src.head_len ? p.src.head_len[s] : 0;
    const int total_len = head_len + p.src.body_len;
    int nfr = p.nframes ? p.nframes[s] : (total_len >= kFft ? 1 + (total_len - kFft) / kHop : 0);
    if (nfr > p.max_frames) nfr = p.max_frames;

    // ---- stage the three pair windows (zero beyond the signal)
    for (int idx = lane; idx < kFePairs * fft::kPairWindow; idx += 32) {
      const int t = idx / fft::kPairWindow;
      const int o = idx - t * fft::kPairWindow;
      const int fa = f0 + 2 * t;
      const int i = fa * kHop + o;
      float v = 0.0f;
      if (fa < nfr && i < total_len) v = load_sample(p.src, s, i, head_len);
      win[t * kFeWinStride + o] = v;
    }
    __syncwarp();

    const int fa = f0 + 2 * team;
    const bool pair_live = active && fa < nfr;
    cpx* my_buf = buf + (active ? team : 0) * fft::kBufSlots;
    if (pair_live) fft::stage1(j, win + team * kFeWinStride, tw, my_buf);
    __syncwarp();
    if (pair_live) fft::stage2(j, my_buf);
    __syncwarp();
    float mag_a[21], mag_b[21];
    int cnt = 0;
    if (pair_live) cnt = fft::untangle(j, my_buf, mag_a, mag_b);
    __syncwarp();
    if (pair_live) {
      float* mg = reinterpret_cast<float*>(my_buf);     // [2][201] magnitudes over the dead Z buffer
#pragma unroll
      for (int i = 0; i < 21; ++i)
        if (i < cnt) {
          mg[j + 10 * i] = mag_a[i];
          mg[kBins + j + 10 * i] = mag_b[i];
        }
    }
    __syncwarp();
    if (pair_live) {
      const float* mg = reinterpret_cast<const float*>(my_buf);
      const bool b_live = fa + 1 < nfr;
      float* out_a = p.mel_out + (s * p.max_frames + fa) * p.n_mel;
      for (int m = j; m < p.n_mel; m += fft::kThreads) {
        const int k0 = mel_start[m], c = mel_count[m];
        const float* wv = mel_w + mel_off[m];
        float acc_a = 0.0f, acc_b = 0.0f;
        for (int i = 0; i < c; ++i) {
          const float wgt = wv[i];
          acc_a = fmaf(mg[k0 + i], wgt, acc_a);
          acc_b = fmaf(mg[kBins + k0 + i], wgt, acc_b);
        }
        out_a[m] = acc_a;
        if (b_live) out_a[p.n_mel + m] = acc_b;
      }
    }
    __syncwarp();
  }
}

static size_t frontend_smem_bytes(const kws_model* m) {
  size_t head = sizeof(cpx) * fft::kN + sizeof(float) * m->mel.nnz + sizeof(int) * 3 * m->cfg.n_mel;
  head = (head + 15) & ~static_cast<size_t>(15);
  size_t per_warp = sizeof(cpx) * kFePairs * fft::kBufSlots + sizeof(float) * kFePairs * kFeWinStride;
  per_warp = (per_warp + 15) & ~static_cast<size_t>(15);
  return head + per_warp * kFeWarps;
}

int launch_frontend(const kws_model* m, const PcmSource& src, int64_t S, int32_t max_frames,
                    const int32_t* nframes, float* mel_out, cudaStream_t st) {
  if (S <= 0 || max_frames <= 0) return KWS_OK;
  FrontendParams p;
  p.src = src;
  p.S = S;
  p.max_frames = max_frames;
  p.groups = static_cast<int>(ceil_div(max_frames, kFeFramesPerItem));
  p.nframes = nframes;
  p.n_mel = m->cfg.n_mel;
  p.twiddle = reinterpret_cast<const cpx*>(m->twiddle400);
  p.mel_start = m->mel.start;
  p.mel_count = m->mel.count;
  p.mel_offset = m->mel.offset;
  p.mel_weight = m->mel.weight;
  p.mel_nnz = m->mel.nnz;
  p.mel_out = mel_out;
  const size_t smem = frontend_smem_bytes(m);
  static thread_local size_t configured = 0;
  if (configured < smem) {
    KWS_CUDA_OK(cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    configured = smem;
  }
  int per_sm = 0;
  KWS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frontend_kernel, kFeWarps * 32, smem));
  if (per_sm < 1) return fail(KWS_ERR_CUDA, "frontend_kernel does not fit on an SM (smem %zu)", smem);
  const long items = S * p.groups;
  long blocks = ceil_div(items, kFeWarps);
  const long resident = static_cast<long>(per_sm) * sm_count();
  if (blocks > resident) blocks = resident;
  frontend_kernel<<<static_cast<unsigned>(blocks), kFeWarps * 32, smem, st>>>(p);
  KWS_LAUNCH_OK("frontend_kernel");
  return KWS_OK;
}

}  // namespace kws
