// CPU build of the device decoders (keyword_spotting_b200/csrc/decode_core.cuh), used only
// to check that logic against the reference's golden vectors without a GPU.
#include "../../keyword_spotting_b200/csrc/decode_core.cuh"

using namespace kws::dec;

struct Row {
  const float* p;
  int C;
  float operator()(int t, int c) const { return p[t * C + 1 + c]; }
};

extern "C" int decode_host(const float* probs, int T, int C, int mode, int lockout, double thres,
                           double loose, const char* keyword, int32_t* out, int max_out, int* trigger) {
  Params prm;
  prm.mode = mode;
  prm.lockout = lockout;
  prm.thres = thres;
  prm.loose_thres = loose;
  prm.ncols = mode == KWS_DECODE_CTC ? (C - 1 < 4 ? C - 1 : 4) : C - 2;
  Sink sink;
  sink.init(out, max_out, parse_keyword(keyword));
  Row row{probs, C};
  decode(row, T, prm, sink);
  *trigger = sink.hit;
  return sink.count();
}
