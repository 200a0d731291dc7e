// CPU emulation of the 20-thread frame-pair FFT used by the K1 front-end kernel
// (keyword_spotting_b200/csrc/fft400.cuh).  Built with g++ by tests/test_host.py.
#include <cmath>
#include "../../keyword_spotting_b200/csrc/fft400.cuh"

using namespace kws::fft;

extern "C" void fft400_pair_mags(const float* win560, float* mag_a201, float* mag_b201) {
  static cpx twt[kN];
  static bool init = false;
  if (!init) {
    for (int n1 = 0; n1 < kR; ++n1)
      for (int k2 = 0; k2 < kR; ++k2) {
        const double a = -2.0 * M_PI * (n1 * k2) / kN;
        twt[twt_index(n1, k2)].re = (float)std::cos(a);
        twt[twt_index(n1, k2)].im = (float)std::sin(a);
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
    stage1_col(c, regs[c], twt, buf);
  }                                                  // --- barrier ---
  for (int c = 0; c < kR; ++c) stage2_col(c, buf, regs[c]);   // --- barrier ---
  // emulate the in-place reuse: every thread reads its mirror slots first only if no writer got there before;
  // on the GPU the written region (rows 0..9) and the read region (rows 10..19) are disjoint, so order is free.
  float* mag = reinterpret_cast<float*>(buf);
  for (int c = kR - 1; c >= 0; --c) untangle_col(c, regs[c], buf, 0.5f, mag);
  for (int k = 0; k < kBins; ++k) {
    mag_a201[k] = mag[k];
    mag_b201[k] = mag[kMagB + k];
  }
}

extern "C" void dft20_host(const float* in40, float* out40) {
  cpx v[20];
  for (int i = 0; i < 20; ++i) { v[i].re = in40[2 * i]; v[i].im = in40[2 * i + 1]; }
  dft20_pfa(v);
  for (int k = 0; k < 20; ++k) { out40[2 * k] = v[pfa_slot(k)].re; out40[2 * k + 1] = v[pfa_slot(k)].im; }
}
