// Per-stream CTC peak decoders + keyword test, shared by the device kernels and by a
// CPU build used only to test this logic against the reference's golden vectors
// (tests/host/decode_host.cpp).
//
// Reference: utils/prediction.py
//   ctc_decode        :18-62   lockout / threshold / 'loose' mode after 1,2,3
//   ctc_decode2       :65-86   streaming decoder used by detector.py:200
//   ctc_decode_strict :89-108
//   ctc_predict       :111-118 decimal-concatenate labels, substring test
// Probabilities are fp32; thresholds are doubles and the comparison is done in
// double, which is what `np.float32 > python_float` did under numpy 1.x.
#pragma once

#include <stdint.h>

#include "../../include/kws_b200.h"

#if defined(__CUDACC__)
#define KWS_HD __host__ __device__ __forceinline__
#else
#define KWS_HD inline
#endif

namespace kws {
namespace dec {

struct Params {
  int mode;            // kws_decode_mode
  int lockout;
  double thres;
  double loose_thres;
  int ncols;           // number of word columns examined (starting at class 1)
};

struct Keyword {
  uint64_t pattern;    // 4 bits per label, most recent label in the low nibble
  uint64_t mask;
  int length;          // 0 => '' in text is always True
  int impossible;      // contains a character that no label can produce
};

KWS_HD Keyword parse_keyword(const char* s) {
  Keyword k;
  k.pattern = 0;
  k.mask = 0;
  k.length = 0;
  k.impossible = 0;
  for (int i = 0; s && s[i] != 0; ++i) {
    const int d = s[i] - '0';
    if (d < 1 || d > 9) k.impossible = 1;
    k.pattern = (k.pattern << 4) | static_cast<uint64_t>(d & 15);
    k.mask = (k.mask << 4) | 15u;
    ++k.length;
    if (k.length > 16) {
      k.impossible = 1;
      break;
    }
  }
  return k;
}

// Collects labels in the reference output form [0, l1, 0, l2, 0, ...] and tracks the
// keyword as a substring of the label string.
struct Sink {
  int32_t* out;        // may be null
  int max_out;
  int nlabels;
  uint64_t recent;     // shift register of the last 16 labels
  int hit;
  Keyword kw;

  KWS_HD void init(int32_t* o, int mx, const Keyword& k) {
    out = o;
    max_out = mx;
    nlabels = 0;
    recent = 0;
    kw = k;
    hit = (k.length == 0 && !k.impossible) ? 1 : 0;
    if (out && max_out > 0) out[0] = 0;
  }
  KWS_HD void push(int label) {
    const int pos = 2 * nlabels + 1;
    if (out) {
      if (pos < max_out) out[pos] = label;
      if (pos + 1 < max_out) out[pos + 1] = 0;
    }
    ++nlabels;
    recent = (recent << 4) | static_cast<uint64_t>(label & 15);
    if (!kw.impossible && kw.length > 0 && nlabels >= kw.length && (recent & kw.mask) == kw.pattern) hit = 1;
  }
  KWS_HD int count() const { return 2 * nlabels + 1; }
  KWS_HD void finish() {
    if (out)
      for (int i = count(); i < max_out; ++i) out[i] = -1;
  }
};

// frame summary over the word columns: max value, first arg-max
template <typename RowFn>
KWS_HD void row_max(const RowFn& row, int t, int ncols, float& mx, int& arg) {
  mx = row(t, 0);
  arg = 0;
  for (int c = 1; c < ncols; ++c) {
    const float v = row(t, c);
    if (v > mx) {      // strict: numpy argmax keeps the first maximum
      mx = v;
      arg = c;
    }
  }
}

// row(t, c): probability of class (c + 1) at frame t
template <typename RowFn>
KWS_HD void decode(const RowFn& row, int T, const Params& p, Sink& sink) {
  if (p.ncols <= 0) {          // empty slice: max() of nothing -- the reference raises; we emit nothing
    sink.finish();
    return;
  }
  if (p.mode == KWS_DECODE_CTC2) {
    int prev = -1;
    for (int t = 0; t < T; ++t) {
      float mx;
      int arg;
      row_max(row, t, p.ncols, mx, arg);
      if (static_cast<double>(mx) > p.thres) {
        if (prev == -1 || prev != arg) sink.push(arg + 1);
        prev = arg;
      } else {
        prev = -1;
      }
    }
  } else if (p.mode == KWS_DECODE_STRICT) {
    int t = 0;
    while (t < T) {
      float mx;
      int arg;
      row_max(row, t, p.ncols, mx, arg);
      if (static_cast<double>(mx) > p.thres) {
        sink.push(arg + 1);
        t += p.lockout;
      } else {
        ++t;
      }
    }
  } else {
    // ctc_decode: l1 is the newest label, l3 the third newest; last_t the frame of the newest
    int l1 = 0, l2 = 0, l3 = 0, last_t = 0;
    bool loose = false;
    int t = 0;
    while (t < T) {
      float mx;
      int arg;
      row_max(row, t, p.ncols, mx, arg);
      if (loose) {
        if (static_cast<double>(mx) < p.loose_thres) {
          if (l1 != 3) {
            t += p.lockout;
            loose = false;
            continue;
          }
        } else if (p.ncols > 2 && static_cast<double>(row(t, 2)) > p.loose_thres) {
          sink.push(3);
          l3 = l2; l2 = l1; l1 = 3; last_t = t;
          t += p.lockout;
          loose = false;
          continue;
        } else if (static_cast<double>(mx) > 0.6 && last_t + p.lockout < t) {
          sink.push(arg + 1);
          l3 = l2; l2 = l1; l1 = arg + 1; last_t = t;
        }
      } else if (static_cast<double>(mx) > p.thres) {
        sink.push(arg + 1);
        l3 = l2; l2 = l1; l1 = arg + 1; last_t = t;
        t += p.lockout;
        if (l3 == 1 && l2 == 2 && l1 == 3) loose = true;
        continue;
      }
      ++t;
    }
  }
  sink.finish();
}

// ---- per-frame token of ctc_decode2: winner column if above threshold, else -1
template <typename RowFn>
KWS_HD int frame_token(const RowFn& row, int t, int ncols, double thres) {
  float mx;
  int arg;
  row_max(row, t, ncols, mx, arg);
  return static_cast<double>(mx) > thres ? arg : -1;
}

}  // namespace dec
}  // namespace kws
