// CPU emulation of the 10-thread frame-pair FFT used by the K1 front-end kernel
// (keyword_spotting_b200/csrc/fft400.cuh).  Built with g++ by tests/test_fft400_host.py.
#include <cmath>
#include "../../keyword_spotting_b200/csrc/fft400.cuh"

using namespace kws::fft;

extern "C" void fft400_pair_mags(const float* win560, float* mag_a201, float* mag_b201) {
  static cpx tw[kN];
  static bool init = false;
  if (!init) {
    for (int m = 0; m < kN; ++m) {
      const double a = -2.0 * M_PI * m / kN;
      tw[m].re = (float)std::cos(a);
      tw[m].im = (float)std::sin(a);
    }
    init = true;
  }
  cpx buf[kBufSlots];
  for (int j = 0; j < kThreads; ++j) stage1(j, win560, tw, buf);   // --- __syncwarp ---
  for (int j = 0; j < kThreads; ++j) stage2(j, buf);               // --- __syncwarp ---
  for (int j = 0; j < kThreads; ++j) {
    float ma[21], mb[21];
    const int cnt = untangle(j, buf, ma, mb);
    for (int i = 0; i < cnt; ++i) {
      mag_a201[j + 10 * i] = ma[i];
      mag_b201[j + 10 * i] = mb[i];
    }
  }
}

extern "C" void dft20_host(const float* in40, float* out40) {
  cpx v[20];
  for (int i = 0; i < 20; ++i) { v[i].re = in40[2 * i]; v[i].im = in40[2 * i + 1]; }
  dft20_pfa(v);
  for (int k = 0; k < 20; ++k) { out40[2 * k] = v[pfa_slot(k)].re; out40[2 * k + 1] = v[pfa_slot(k)].im; }
}
