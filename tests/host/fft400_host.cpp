// CPU emulation of the 20-thread frame-pair FFT used by the K1 front-end kernel
// (keyword_spotting_b200/csrc/fft400.cuh).  Built with g++ by tests/test_host.py.
#include <cmath>
#include "../../keyword_spotting_b200/csrc/fft400.cuh"

using namespace kws::fft;

extern "C" void fft400_pair_mags(const float* win560, float* mag_a201, float* mag_b201) {
  static cpx twp[kTwSlots];
  static bool init = false;
  if (!init) {
    for (int k2 = 0; k2 < kR; ++k2)
      for (int j = 0; j < kTwStride; ++j) {
        const double a = -2.0 * M_PI * ((j % kR) * k2) / kN;
        twp[k2 * kTwStride + j].re = (float)std::cos(a);
        twp[k2 * kTwStride + j].im = (float)std::sin(a);
      }
    init = true;
  }
  cpx buf[kBufSlots];
  cpx regs[kR][20];                                 // per-"thread" register state
  for (int c = 0; c < kR; ++c) {
    for (int n2 = 0; n2 < 20; ++n2) {
      regs[c][n2].re = win560[c + 20 * n2];
      regs[c][n2].im = win560[160 + c + 20 * n2];
    }
    // thread t = 20*pair + c of a CTA; emulate a pair that straddles a warp boundary (pair 1 -> t = 20 + c)
    const int t = 20 + c;
    stage1_col(c, regs[c], twp + tw_base(t >> 5, t & 31), buf);
  }                                                  // --- barrier ---
  for (int c = 0; c < kR; ++c) stage2_col(c, buf, regs[c]);   // --- barrier ---
  for (int c = kR - 1; c >= 0; --c) untangle_col(c, regs[c], buf, mag_a201 + c, mag_b201 + c, kR);   // 2|A|, 2|B|
  for (int k = 0; k < kBins; ++k) { mag_a201[k] *= 0.5f; mag_b201[k] *= 0.5f; }
}

extern "C" void dft20_host(const float* in40, float* out40) {
  cpx v[20];
  for (int i = 0; i < 20; ++i) { v[i].re = in40[2 * i]; v[i].im = in40[2 * i + 1]; }
  dft20_pfa(v);
  for (int k = 0; k < 20; ++k) { out40[2 * k] = v[pfa_slot(k)].re; out40[2 * k + 1] = v[pfa_slot(k)].im; }
}
