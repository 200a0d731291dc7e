// K4 -- batched CTC peak decoders + keyword test on the device.
// Reference: utils/prediction.py:18-118 (see decode_core.cuh for the per-stream logic).
// One thread per stream: each decoder is a strictly sequential scan with a lockout /
// mode state, and a stream's [T, C] probabilities are contiguous, so a thread walks its
// own cache lines.  Bit-exact by construction: integer outputs, fp32 inputs compared in
// double exactly as numpy did.
#include <cstring>

#include "common.cuh"
#include "decode_core.cuh"

namespace kws {

struct ProbRow {
  const float* p;   // stream base, [T, C]
  int C;
  __device__ __forceinline__ float operator()(int t, int c) const { return __ldg(p + t * C + 1 + c); }
};

__global__ void __launch_bounds__(128)
ctc_decode_kernel(const float* __restrict__ probs, long S, int T, int C, const int* __restrict__ lens,
                  dec::Params prm, dec::Keyword kw, int* __restrict__ labels, int max_labels,
                  int* __restrict__ counts, int* __restrict__ trigger) {
  const long s = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (s >= S) return;
  int len = lens ? lens[s] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  dec::Sink sink;
  sink.init(labels ? labels + s * max_labels : nullptr, max_labels, kw);
  ProbRow row{probs + s * static_cast<long>(T) * C, C};
  dec::decode(row, len, prm, sink);
  if (counts) counts[s] = sink.count();
  if (trigger) trigger[s] = sink.hit;
}

int decode_params_from_api(const kws_decode_params* in, int C, dec::Params* out) {
  KWS_REQUIRE(in != nullptr, "params is NULL");
  KWS_REQUIRE(in->mode >= KWS_DECODE_CTC && in->mode <= KWS_DECODE_STRICT, "unknown decode mode %d", in->mode);
  KWS_REQUIRE(C >= 3 && C <= 11, "num_classes must be in [3, 11] (labels are single decimal digits)");
  out->mode = in->mode;
  out->lockout = in->lockout;
  if (in->mode != KWS_DECODE_CTC2) KWS_REQUIRE(in->lockout >= 1, "lockout must be >= 1");
  const double def = in->mode == KWS_DECODE_CTC2 ? 0.4 : 0.5;      // prediction.py:18,65,89 defaults
  out->thres = in->thres < 0 ? def : in->thres;
  out->loose_thres = in->loose_thres;
  // ctc_decode slices columns 1:5 whatever the class count (:21); the others 1:classnum-1 (:67,:92)
  out->ncols = in->mode == KWS_DECODE_CTC ? (C - 1 < 4 ? C - 1 : 4) : C - 2;
  return KWS_OK;
}

// Levenshtein distance of S (reference, hypothesis) label pairs -- the integer the reference's WER divides by
// len(r) (utils/wer.py:4-41).  One thread per pair, single rolling row; lengths <= 254 as in the reference
// (its table is uint8).
constexpr int kWerMaxLen = 254;
__global__ void __launch_bounds__(128)
edit_distance_kernel(const int* __restrict__ ref, const int* __restrict__ ref_len, long ld_ref,
                     const int* __restrict__ hyp, const int* __restrict__ hyp_len, long ld_hyp, long S,
                     int* __restrict__ dist) {
  const long s = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (s >= S) return;
  const int nr = ref_len[s], nh = hyp_len[s];
  const int* r = ref + s * ld_ref;
  const int* h = hyp + s * ld_hyp;
  unsigned char row[kWerMaxLen + 1];              // d[i-1][*] rolling into d[i][*]
  for (int j = 0; j <= nh; ++j) row[j] = static_cast<unsigned char>(j);
  for (int i = 1; i <= nr; ++i) {
    int diag = row[0];                            // d[i-1][j-1]
    row[0] = static_cast<unsigned char>(i);
    const int ri = r[i - 1];
    for (int j = 1; j <= nh; ++j) {
      const int up = row[j];                      // d[i-1][j]
      int v;
      if (ri == h[j - 1]) {
        v = diag;
      } else {
        const int sub = diag + 1, ins = row[j - 1] + 1, del = up + 1;
        v = sub < ins ? sub : ins;
        v = v < del ? v : del;
      }
      diag = up;
      row[j] = static_cast<unsigned char>(v);
    }
  }
  dist[s] = row[nh];
}

}  // namespace kws

extern "C" int kws_edit_distance(const int32_t* ref, const int32_t* ref_len, int64_t ld_ref, const int32_t* hyp,
                                 const int32_t* hyp_len, int64_t ld_hyp, int64_t S, int32_t max_len,
                                 int32_t* dist_out, void* stream) {
  using namespace kws;
  clear_error();
  KWS_REQUIRE(S >= 0, "negative size");
  if (S == 0) return KWS_OK;
  KWS_REQUIRE(ref && ref_len && hyp && hyp_len && dist_out, "NULL pointer");
  KWS_REQUIRE(max_len >= 0 && max_len <= kWerMaxLen, "sequences longer than %d labels are outside the reference's domain "
              "(utils/wer.py:8, uint8 table)", kWerMaxLen);
  KWS_REQUIRE(ld_ref >= max_len && ld_hyp >= max_len, "row stride smaller than max_len");
  edit_distance_kernel<<<static_cast<unsigned>(ceil_div(S, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      ref, ref_len, ld_ref, hyp, hyp_len, ld_hyp, S, dist_out);
  KWS_LAUNCH_OK("edit_distance_kernel");
  return KWS_OK;
}

extern "C" int kws_ctc_decode(const float* probs, int64_t S, int32_t T, int32_t C, const int32_t* lens,
                              const kws_decode_params* params, const char* keyword, int32_t* labels_out,
                              int32_t max_labels, int32_t* counts_out, int32_t* trigger_out, void* stream) {
  using namespace kws;
  clear_error();
  KWS_REQUIRE(S >= 0 && T >= 0, "negative size");
  dec::Params prm;
  int rc = decode_params_from_api(params, C, &prm);
  if (rc != KWS_OK) return rc;
  KWS_REQUIRE(keyword != nullptr && std::strlen(keyword) <= 16, "keyword must be 0..16 characters");
  KWS_REQUIRE(labels_out == nullptr || max_labels >= 1, "max_labels must be >= 1");
  if (S == 0) return KWS_OK;
  KWS_REQUIRE(probs != nullptr || T == 0, "probs is NULL");
  const dec::Keyword kw = dec::parse_keyword(keyword);
  const int threads = 128;
  ctc_decode_kernel<<<static_cast<unsigned>(ceil_div(S, threads)), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      probs, S, T, C, lens, prm, kw, labels_out, max_labels, counts_out, trigger_out);
  KWS_LAUNCH_OK("ctc_decode_kernel");
  return KWS_OK;
}
