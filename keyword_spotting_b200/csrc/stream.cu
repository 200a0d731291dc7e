// Streaming server: the HotwordDetector.start loop (detector.py:148-209) for S lock-step
// streams, every piece of per-stream state resident in HBM.
//
// Per chunk and stream, in the reference's order:
//   1. VAD on the NEW samples: sum|x| > 30 (utils/basic_vad.py:17-18, detector.py:168).
//      Done in exact integer arithmetic on the int16 PCM: sum|x_i16| > 30*32768.
//      Silence -> GRU state zeroed and decision window cleared (detector.py:171-177).
//   2. carried tail ++ chunk; keep (len-400)%160+240 samples for next time (detector.py:179-183)
//   3. model call on the concatenation (detector.py:190-193)         -> K1, K2+K3
//   4. window.add(softmax) with at most 15 chunks (utils/queue.py:16-38, detector.py:195)
//   5. ctc_decode2 over the whole window (detector.py:197-200), ctc_predict(.., '1233')
//   6. trigger -> window cleared, state zeroed (detector.py:201-208)
// The window keeps, per frame, only what ctc_decode2 consumes: the above-threshold winner
// column or -1 (one byte), so re-decoding the 450-frame window every chunk costs 480 B of
// reads per stream instead of 10.8 KB of fp32 probabilities.
#include <atomic>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "decode_core.cuh"
#include "stream_internal.cuh"

namespace kws {

constexpr int kTailCap = 400;

// One warp per stream: VAD, frame count, next tail.
__global__ void __launch_bounds__(256)
stream_pre_kernel(const int16_t* __restrict__ pcm, long ld, int chunk_len, long S, long long vad_limit,
                  const int16_t* __restrict__ tail_cur, const int* __restrict__ len_cur,
                  int16_t* __restrict__ tail_next, int* __restrict__ len_next,
                  unsigned char* __restrict__ silence, int* __restrict__ nframes, int max_frames) {
  const long s = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  const int16_t* row = pcm + s * ld;
  long long sum = 0;
  if (((reinterpret_cast<uintptr_t>(row) & 3) == 0)) {
    const int pairs = chunk_len >> 1;
    const unsigned* row2 = reinterpret_cast<const unsigned*>(row);
    int acc = 0;
    for (int i = lane; i < pairs; i += 32) {
      const unsigned v = __ldg(row2 + i);
      const int a = static_cast<short>(v & 0xffffu), b = static_cast<short>(v >> 16);
      acc += abs(a) + abs(b);            // |int16| <= 32768; a lane sums < 2^31 for chunks < 2^21 samples
    }
    sum = acc;
    if ((chunk_len & 1) && lane == 0) sum += abs(static_cast<int>(row[chunk_len - 1]));
  } else {
    int acc = 0;
    for (int i = lane; i < chunk_len; i += 32) acc += abs(static_cast<int>(row[i]));
    sum = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const int head = len_cur[s];
  const int total = head + chunk_len;
  const int nfr = total >= kFft ? 1 + (total - kFft) / kHop : 0;
  const int keep = total >= kFft ? (total - kFft) % kHop + (kFft - kHop) : total;   // detector.py:181-182
  if (lane == 0) {
    silence[s] = sum > vad_limit ? 0 : 1;
    nframes[s] = nfr < max_frames ? nfr : max_frames;
    len_next[s] = keep;
  }
  const int16_t* tcur = tail_cur + s * kTailCap;
  int16_t* tnext = tail_next + s * kTailCap;
  for (int i = lane; i < keep; i += 32) {
    const int gidx = total - keep + i;
    tnext[i] = gidx < head ? tcur[gidx] : row[gidx - head];
  }
}

struct TokRow {
  const float* p;
  int C;
  __device__ __forceinline__ float operator()(int t, int c) const { return __ldg(p + t * C + 1 + c); }
};

// One thread per stream: push this chunk's tokens, re-decode the window, trigger side effects.
__global__ void __launch_bounds__(128)
stream_post_kernel(const float* __restrict__ probs, int n_step, int C, long S, int W, int fpad,
                   double thres, dec::Keyword kw, const unsigned char* __restrict__ silence,
                   const int* __restrict__ nframes, signed char* __restrict__ tok,
                   unsigned char* __restrict__ slot_frames, int* __restrict__ win_head,
                   int* __restrict__ win_n, float* __restrict__ state, int layers,
                   int* __restrict__ trigger_a, int* __restrict__ trigger_b) {
  const long s = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (s >= S) return;
  int head = win_head[s], n = win_n[s];
  if (silence[s]) {            // prob_queue.clear()  (detector.py:177)
    head = 0;
    n = 0;
  }
  // SimpleQueue.add (utils/queue.py:29-35)
  if (n == W) head = (head + 1) % W; else ++n;
  const int slot = (head + n - 1) % W;
  const int nf = nframes[s];
  signed char* trow = tok + (s * W + slot) * static_cast<long>(fpad);
  TokRow row{probs + s * static_cast<long>(n_step) * C, C};
  for (int t = 0; t < nf; ++t) trow[t] = static_cast<signed char>(dec::frame_token(row, t, C - 2, thres));
  slot_frames[s * W + slot] = static_cast<unsigned char>(nf);
  // ctc_decode2 over the concatenated window (detector.py:197-200) + ctc_predict
  dec::Sink sink;
  sink.init(nullptr, 0, kw);
  int prev = -1;
  for (int q = 0; q < n; ++q) {
    const int sl = (head + q) % W;
    const signed char* r = tok + (s * W + sl) * static_cast<long>(fpad);
    const int f = slot_frames[s * W + sl];
    for (int t = 0; t < f; ++t) {
      const int tk = r[t];
      if (tk >= 0) {
        if (prev == -1 || prev != tk) sink.push(tk + 1);
        prev = tk;
      } else {
        prev = -1;
      }
    }
  }
  const int hit = sink.hit;
  if (hit) {                   // detector.py:201-208: clear the window, zero the state
    head = 0;
    n = 0;
    for (int l = 0; l < layers; ++l) {
      float4* h = reinterpret_cast<float4*>(state + (static_cast<long>(l) * S + s) * kHidden);
      for (int i = 0; i < kHidden / 4; ++i) h[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  win_head[s] = head;
  win_n[s] = n;
  if (trigger_a) trigger_a[s] = hit;
  if (trigger_b) trigger_b[s] = hit;
}

__global__ void __launch_bounds__(128)
stream_labels_kernel(long S, int W, int fpad, dec::Keyword kw, const signed char* __restrict__ tok,
                     const unsigned char* __restrict__ slot_frames, const int* __restrict__ win_head,
                     const int* __restrict__ win_n, int* __restrict__ labels, int max_labels,
                     int* __restrict__ counts) {
  const long s = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (s >= S) return;
  const int head = win_head[s], n = win_n[s];
  dec::Sink sink;
  sink.init(labels ? labels + s * max_labels : nullptr, max_labels, kw);
  int prev = -1;
  for (int q = 0; q < n; ++q) {
    const int sl = (head + q) % W;
    const signed char* r = tok + (s * W + sl) * static_cast<long>(fpad);
    const int f = slot_frames[s * W + sl];
    for (int t = 0; t < f; ++t) {
      const int tk = r[t];
      if (tk >= 0) {
        if (prev == -1 || prev != tk) sink.push(tk + 1);
        prev = tk;
      } else {
        prev = -1;
      }
    }
  }
  sink.finish();
  if (counts) counts[s] = sink.count();
}

// zero the carried state of silent streams when a chunk produces no frame at all
__global__ void zero_silent_state_kernel(float* state, long S, int layers, const unsigned char* silence) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long total = static_cast<long>(layers) * S * kHidden;
  if (idx >= total) return;
  const long s = (idx / kHidden) % S;
  if (silence[s]) state[idx] = 0.0f;
}

template <typename T>
static int dev_alloc(T** p, size_t n, bool zero) {
  *p = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * (n ? n : 1));
  if (e != cudaSuccess) return fail(KWS_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", sizeof(T) * n, cudaGetErrorString(e));
  if (zero) KWS_CUDA_OK(cudaMemset(*p, 0, sizeof(T) * (n ? n : 1)));
  return KWS_OK;
}

static void free_stream(kws_stream* st) {
  if (!st) return;
  cudaFree(st->state);
  for (int i = 0; i < 2; ++i) {
    cudaFree(st->tail[i]);
    cudaFree(st->tail_len[i]);
    cudaFree(st->pcm_dev[i]);
    if (st->copied[i]) cudaEventDestroy(st->copied[i]);
    if (st->consumed[i]) cudaEventDestroy(st->consumed[i]);
  }
  if (st->copy_stream) cudaStreamDestroy(st->copy_stream);
  cudaFree(st->silence);
  cudaFree(st->nframes);
  cudaFree(st->mel);
  cudaFree(st->seq);
  cudaFree(st->y_rows);
  cudaFree(st->probs);
  cudaFree(st->tok);
  cudaFree(st->slot_frames);
  cudaFree(st->win_head);
  cudaFree(st->win_n);
  cudaFree(st->trigger);
  delete st;
}

static int frames_for_chunk(int chunk_len) {      // with the longest possible carried tail (399)
  const int total = chunk_len + kTailCap - 1;
  return total >= kFft ? 1 + (total - kFft) / kHop : 0;
}

}  // namespace kws

using namespace kws;

extern "C" int kws_stream_create(kws_model* m, const kws_stream_config* cfg, kws_stream** out) {
  clear_error();
  KWS_REQUIRE(m && cfg && out, "NULL argument");
  *out = nullptr;
  KWS_REQUIRE(cfg->n_streams >= 1, "n_streams must be >= 1");
  KWS_REQUIRE(cfg->max_chunk >= 1 && cfg->max_chunk <= (1 << 20), "max_chunk out of range");
  KWS_REQUIRE(cfg->window_chunks >= 1 && cfg->window_chunks <= 255, "window_chunks must be in [1, 255]");
  KWS_REQUIRE(cfg->vad_threshold >= 0, "vad_threshold must be >= 0");
  KWS_REQUIRE(m->cfg.num_classes >= 3 && m->cfg.num_classes <= 11, "streaming decode needs 3..11 classes");
  KWS_REQUIRE(std::memchr(cfg->keyword, 0, sizeof(cfg->keyword)) != nullptr && std::strlen(cfg->keyword) <= 16,
              "keyword must be a NUL-terminated string of at most 16 characters");
  const int mf = frames_for_chunk(cfg->max_chunk);
  KWS_REQUIRE(mf <= 255, "max_chunk produces more than 255 frames per chunk");
  KWS_CUDA_OK(cudaSetDevice(m->device));
  kws_stream* st = new kws_stream();
  st->model = m;
  st->device = m->device;
  st->cfg = *cfg;
  st->S = cfg->n_streams;
  st->max_frames = mf;
  st->fpad = (mf + 15) / 16 * 16;
  if (st->fpad == 0) st->fpad = 16;
  st->kw = dec::parse_keyword(cfg->keyword);
  const size_t S = static_cast<size_t>(st->S);
  const int L = m->cfg.num_layers, C = m->cfg.num_classes, M = m->cfg.n_mel, W = cfg->window_chunks;
  int rc = dev_alloc(&st->state, S * L * kHidden, true);
  for (int i = 0; i < 2 && rc == KWS_OK; ++i) {
    rc = dev_alloc(&st->tail[i], S * kTailCap, true);
    if (rc == KWS_OK) rc = dev_alloc(&st->tail_len[i], S, true);
  }
  if (rc == KWS_OK) rc = dev_alloc(&st->silence, S, true);
  if (rc == KWS_OK) rc = dev_alloc(&st->nframes, S, true);
  if (rc == KWS_OK) rc = dev_alloc(&st->mel, mel_scratch_elems(st->S, mf, M), false);
  if (rc == KWS_OK) rc = dev_alloc(&st->probs, S * (mf ? mf : 1) * C, false);
  if (rc == KWS_OK) rc = dev_alloc(&st->tok, S * W * st->fpad, true);
  if (rc == KWS_OK) rc = dev_alloc(&st->slot_frames, S * W, true);
  if (rc == KWS_OK) rc = dev_alloc(&st->win_head, S, true);
  if (rc == KWS_OK) rc = dev_alloc(&st->win_n, S, true);
  if (rc == KWS_OK) rc = dev_alloc(&st->trigger, S, true);
  if (rc == KWS_OK && mf > 0 && L > 1) rc = dev_alloc(&st->seq, seq_scratch_elems(st->S, mf, L), false);
  if (rc == KWS_OK && mf > 0 && m->octbit) rc = dev_alloc(&st->y_rows, S * mf * kHidden, false);
  if (rc != KWS_OK) {
    free_stream(st);
    return rc;
  }
  *out = st;
  return KWS_OK;
}

extern "C" int kws_stream_destroy(kws_stream* st) {
  clear_error();
  if (st) {
    cudaSetDevice(st->device);
    cudaDeviceSynchronize();
    free_stream(st);
  }
  return KWS_OK;
}

extern "C" int kws_stream_reset(kws_stream* st, void* stream) {
  clear_error();
  KWS_REQUIRE(st != nullptr, "stream handle is NULL");
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const size_t S = static_cast<size_t>(st->S);
  KWS_CUDA_OK(cudaSetDevice(st->model->device));
  KWS_CUDA_OK(cudaMemsetAsync(st->state, 0, sizeof(float) * S * st->model->cfg.num_layers * kHidden, cs));
  for (int i = 0; i < 2; ++i) KWS_CUDA_OK(cudaMemsetAsync(st->tail_len[i], 0, sizeof(int32_t) * S, cs));
  KWS_CUDA_OK(cudaMemsetAsync(st->win_head, 0, sizeof(int32_t) * S, cs));
  KWS_CUDA_OK(cudaMemsetAsync(st->win_n, 0, sizeof(int32_t) * S, cs));
  return KWS_OK;
}

extern "C" int32_t kws_stream_max_frames(const kws_stream* st) { return st ? st->max_frames : 0; }
extern "C" const float* kws_stream_state(const kws_stream* st) { return st ? st->state : nullptr; }

extern "C" int kws_stream_copy_state(kws_stream* st, float* state_out, void* stream) {
  clear_error();
  KWS_REQUIRE(st != nullptr && state_out != nullptr, "NULL argument");
  KWS_CUDA_OK(cudaSetDevice(st->model->device));
  KWS_CUDA_OK(cudaMemcpyAsync(state_out, st->state,
                              sizeof(float) * st->S * st->model->cfg.num_layers * kHidden,
                              cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return KWS_OK;
}

// Debug: device time of the three parts of a step (front end incl. the fused pre-step | GRU layers | decode + trigger),
// measured with CUDA events on the step's stream.  Enabling it makes every step synchronise -- never on in production.
static std::atomic<int> g_step_timing_on{0};
static std::mutex g_step_mutex;                    // the sums may be fed by stream objects on several host threads
static double g_step_ms[3] = {0.0, 0.0, 0.0};
static long g_step_count = 0;
extern "C" int kws_debug_step_timing(int enable, double* ms_out3, long long* steps_out) {
  kws::clear_error();
  std::lock_guard<std::mutex> lock(g_step_mutex);
  if (ms_out3) {
    for (int i = 0; i < 3; ++i) ms_out3[i] = g_step_ms[i];
  }
  if (steps_out) *steps_out = g_step_count;
  if (enable != g_step_timing_on) {
    g_step_timing_on = enable;
    g_step_ms[0] = g_step_ms[1] = g_step_ms[2] = 0.0;
    g_step_count = 0;
  }
  return KWS_OK;
}

extern "C" int kws_stream_step(kws_stream* st, const int16_t* pcm, int32_t chunk_len, int64_t ld_pcm,
                               int32_t* trigger_out, float* probs_out, int32_t* nframes_out, void* stream) {
  clear_error();
  KWS_REQUIRE(st != nullptr, "stream handle is NULL");
  KWS_REQUIRE(chunk_len >= 1 && chunk_len <= st->cfg.max_chunk, "chunk_len %d outside [1, max_chunk=%d]", chunk_len,
              st->cfg.max_chunk);
  KWS_REQUIRE(pcm != nullptr, "pcm is NULL");
  KWS_REQUIRE(ld_pcm >= chunk_len, "ld_pcm < chunk_len");
  kws_model* m = st->model;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  KWS_CUDA_OK(cudaSetDevice(m->device));
  const long S = st->S;
  const int cur = st->cur, nxt = cur ^ 1;
  const int n_step = frames_for_chunk(chunk_len);
  const int C = m->cfg.num_classes;

  const long long vad_limit = static_cast<long long>(st->cfg.vad_threshold) * 32768LL;
  PcmSource src;
  src.body = pcm;
  src.ld_body = ld_pcm;
  src.body_len = chunk_len;
  src.body_dtype = KWS_PCM_I16;
  src.head = st->tail[cur];
  src.ld_head = kTailCap;
  src.head_len = st->tail_len[cur];
  const bool fused = frontend_can_fuse_pre(m, chunk_len, kTailCap);
  const bool tiled = mel_can_tile(m);
  cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (g_step_timing_on) {
    for (int i = 0; i < 4; ++i) KWS_CUDA_OK(cudaEventCreate(&tev[i]));
    KWS_CUDA_OK(cudaEventRecord(tev[0], cs));
  }
  if (fused) {
    // one pass over the chunk: VAD + tail carry + frame count + framing/FFT/mel
    FrontendPre pre;
    pre.vad_limit = vad_limit;
    pre.tail_next = st->tail[nxt];
    pre.len_next = st->tail_len[nxt];
    pre.silence = st->silence;
    pre.nframes_out = st->nframes;
    int rc = launch_frontend(m, src, S, n_step, nullptr, st->mel, cs, &pre, tiled);
    if (rc != KWS_OK) return rc;
  } else {
    stream_pre_kernel<<<static_cast<unsigned>(ceil_div(S * 32, 256)), 256, 0, cs>>>(
        pcm, ld_pcm, chunk_len, S, vad_limit, st->tail[cur], st->tail_len[cur], st->tail[nxt], st->tail_len[nxt],
        st->silence, st->nframes, n_step);
    KWS_LAUNCH_OK("stream_pre_kernel");
  }

  if (n_step > 0) {
    if (!fused) {
      int rc = launch_frontend(m, src, S, n_step, st->nframes, st->mel, cs, nullptr, tiled);
      if (rc != KWS_OK) return rc;
    }
    if (tev[1]) KWS_CUDA_OK(cudaEventRecord(tev[1], cs));
    int rc = KWS_OK;
    GruArgs a;
    a.x = st->mel;
    a.x_tiled = tiled;
    a.seq_scratch = st->seq;
    a.y_rows_scratch = st->y_rows;
    a.S = S;
    a.n = n_step;
    a.seq_len = st->nframes;
    a.zero_state = st->silence;
    a.state_in = st->state;
    a.state_out = st->state;
    a.probs = st->probs;
    a.logits = nullptr;
    rc = launch_gru(m, a, cs);
    if (rc != KWS_OK) return rc;
  } else {
    const long total = static_cast<long>(m->cfg.num_layers) * S * kHidden;
    zero_silent_state_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, cs>>>(
        st->state, S, m->cfg.num_layers, st->silence);
    KWS_LAUNCH_OK("zero_silent_state_kernel");
  }
  if (tev[2]) KWS_CUDA_OK(cudaEventRecord(tev[2], cs));
  stream_post_kernel<<<static_cast<unsigned>(ceil_div(S, 128)), 128, 0, cs>>>(
      st->probs, n_step, C, S, st->cfg.window_chunks, st->fpad, st->cfg.decode_thres, st->kw, st->silence,
      st->nframes, st->tok, st->slot_frames, st->win_head, st->win_n, st->state, m->cfg.num_layers,
      st->trigger, trigger_out);
  KWS_LAUNCH_OK("stream_post_kernel");
  if (tev[0]) {
    KWS_CUDA_OK(cudaEventRecord(tev[3], cs));
    KWS_CUDA_OK(cudaEventSynchronize(tev[3]));
    if (n_step > 0) {
      std::lock_guard<std::mutex> lock(g_step_mutex);
      for (int i = 0; i < 3; ++i) {
        float ms = 0.0f;
        KWS_CUDA_OK(cudaEventElapsedTime(&ms, tev[i], tev[i + 1]));
        g_step_ms[i] += ms;
      }
      ++g_step_count;
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(tev[i]);
  }
  st->cur = nxt;
  if (probs_out && n_step > 0) {
    KWS_CUDA_OK(cudaMemcpy2DAsync(probs_out, sizeof(float) * st->max_frames * C, st->probs,
                                  sizeof(float) * n_step * C, sizeof(float) * n_step * C, S,
                                  cudaMemcpyDeviceToDevice, cs));
  }
  if (nframes_out)
    KWS_CUDA_OK(cudaMemcpyAsync(nframes_out, st->nframes, sizeof(int32_t) * S, cudaMemcpyDeviceToDevice, cs));
  return KWS_OK;
}

extern "C" int kws_stream_step_host(kws_stream* st, const int16_t* pcm_host, int32_t chunk_len,
                                    int32_t* trigger_host, void* stream) {
  clear_error();
  KWS_REQUIRE(st != nullptr, "stream handle is NULL");
  KWS_REQUIRE(chunk_len >= 1 && chunk_len <= st->cfg.max_chunk, "chunk_len %d outside [1, max_chunk=%d]", chunk_len,
              st->cfg.max_chunk);
  KWS_REQUIRE(pcm_host != nullptr, "pcm_host is NULL");
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  KWS_CUDA_OK(cudaSetDevice(st->model->device));
  const size_t S = static_cast<size_t>(st->S);
  if (!st->copy_stream) {
    KWS_CUDA_OK(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      int rc = dev_alloc(&st->pcm_dev[i], S * st->cfg.max_chunk, false);
      if (rc != KWS_OK) return rc;
      KWS_CUDA_OK(cudaEventCreateWithFlags(&st->copied[i], cudaEventDisableTiming));
      KWS_CUDA_OK(cudaEventCreateWithFlags(&st->consumed[i], cudaEventDisableTiming));
    }
  }
  const int b = st->host_buf;
  // the staging buffer may be overwritten only after the step that read it has finished
  if (st->consumed_valid[b]) KWS_CUDA_OK(cudaStreamWaitEvent(st->copy_stream, st->consumed[b], 0));
  KWS_CUDA_OK(cudaMemcpyAsync(st->pcm_dev[b], pcm_host, sizeof(int16_t) * S * chunk_len, cudaMemcpyHostToDevice,
                              st->copy_stream));
  KWS_CUDA_OK(cudaEventRecord(st->copied[b], st->copy_stream));
  KWS_CUDA_OK(cudaStreamWaitEvent(cs, st->copied[b], 0));
  int rc = kws_stream_step(st, st->pcm_dev[b], chunk_len, chunk_len, nullptr, nullptr, nullptr, stream);
  if (rc != KWS_OK) return rc;
  KWS_CUDA_OK(cudaEventRecord(st->consumed[b], cs));
  st->consumed_valid[b] = true;
  st->host_buf = b ^ 1;
  if (trigger_host)
    KWS_CUDA_OK(cudaMemcpyAsync(trigger_host, st->trigger, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, cs));
  return KWS_OK;
}

extern "C" int kws_stream_labels(kws_stream* st, int32_t* labels_out, int32_t max_labels, int32_t* counts_out,
                                 void* stream) {
  clear_error();
  KWS_REQUIRE(st != nullptr, "stream handle is NULL");
  KWS_REQUIRE(labels_out == nullptr || max_labels >= 1, "max_labels must be >= 1");
  KWS_CUDA_OK(cudaSetDevice(st->model->device));
  stream_labels_kernel<<<static_cast<unsigned>(ceil_div(st->S, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      st->S, st->cfg.window_chunks, st->fpad, st->kw, st->tok, st->slot_frames, st->win_head, st->win_n,
      labels_out, max_labels, counts_out);
  KWS_LAUNCH_OK("stream_labels_kernel");
  return KWS_OK;
}
