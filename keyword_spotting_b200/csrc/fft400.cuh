// 400-point DFT of two real frames at once, split over 10 cooperating threads.
//
// Replaces tf.abs(tf.spectral.rfft(frames, [400])) (models/rnn_ctc.py:137) for
// rectangular-window frames (utils/stft.py:27-81).
//
// Two consecutive real frames a, b are packed as z[n] = a[n] + i*b[n]; one complex
// 400-point DFT Z gives both spectra:  A[k] = (Z[k] + conj(Z[400-k]))/2,
// B[k] = (Z[k] - conj(Z[400-k]))/(2i).  400 = 20 x 20 (Cooley-Tukey):
//   n = n1 + 20*n2,  k = 20*k1 + k2
//   stage 1: Y[n1][k2]  = sum_n2 z[n1+20*n2] * W20^(n2*k2)        (DFT-20 per n1)
//   twiddle: Y'[n1][k2] = Y[n1][k2] * W400^(n1*k2)
//   stage 2: Z[20*k1+k2] = sum_n1 Y'[n1][k2] * W20^(n1*k1)        (DFT-20 per k2)
// Each DFT-20 is a twiddle-free Good-Thomas 4 x 5 prime-factor transform held in
// registers.  Thread j of the 10 owns n1 in {j, j+10} in stage 1 and k2 in
// {j, j+10} in stage 2; the exchange goes through a [20][21]-slot complex buffer
// (row stride padded to 21 to keep the 64-bit accesses on distinct banks).
//
// Everything here is __host__ __device__ so the index logic is unit-tested on the
// CPU (tests/test_fft400_host.py builds tests/fft400_host.cpp with g++).
#pragma once

#if defined(__CUDACC__)
#define KWS_HD __host__ __device__ __forceinline__
#else
#define KWS_HD inline
#endif

namespace kws {
namespace fft {

struct alignas(8) cpx {
  float re, im;
};

constexpr int kN = 400;
constexpr int kR = 20;            // 400 = kR * kR
constexpr int kRowStride = 21;    // padded slots per row of the exchange buffer
constexpr int kBufSlots = kR * kRowStride;   // 420 complex slots
constexpr int kThreads = 10;      // threads cooperating on one frame pair
constexpr int kPairWindow = 560;  // samples spanned by two consecutive frames (160 + 400)

KWS_HD cpx cadd(cpx a, cpx b) { return cpx{a.re + b.re, a.im + b.im}; }
KWS_HD cpx csub(cpx a, cpx b) { return cpx{a.re - b.re, a.im - b.im}; }
KWS_HD cpx cmul(cpx a, cpx b) { return cpx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// forward DFT-4, in place (W4 = -i)
KWS_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
  const cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cpx{t1.re + t3.im, t1.im - t3.re};   // t1 - i*t3
  a3 = cpx{t1.re - t3.im, t1.im + t3.re};   // t1 + i*t3
}

// forward DFT-5, in place (W5 = exp(-2*pi*i/5))
KWS_HD void dft5(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4) {
  const float c1 = 0.30901699437494742f;    // cos(2pi/5)
  const float c2 = -0.80901699437494742f;   // cos(4pi/5)
  const float s1 = 0.95105651629515357f;    // sin(2pi/5)
  const float s2 = 0.58778525229247313f;    // sin(4pi/5)
  const cpx t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const cpx m1 = cpx{a0.re + c1 * t1.re + c2 * t2.re, a0.im + c1 * t1.im + c2 * t2.im};
  const cpx m2 = cpx{a0.re + c2 * t1.re + c1 * t2.re, a0.im + c2 * t1.im + c1 * t2.im};
  const cpx n1 = cpx{s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
  const cpx n2 = cpx{s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
  a0 = cpx{a0.re + t1.re + t2.re, a0.im + t1.im + t2.im};
  a1 = cpx{m1.re + n1.im, m1.im - n1.re};   // m1 - i*n1
  a4 = cpx{m1.re - n1.im, m1.im + n1.re};   // m1 + i*n1
  a2 = cpx{m2.re + n2.im, m2.im - n2.re};   // m2 - i*n2
  a3 = cpx{m2.re - n2.im, m2.im + n2.re};   // m2 + i*n2
}

// where output k of the prime-factor DFT-20 sits in the in-place array
KWS_HD constexpr int pfa_slot(int k) { return (5 * (k % 4) + 4 * (k % 5)) % 20; }

// forward DFT-20 in place on v[0..19] (natural input order).  Afterwards
// X[k] == v[pfa_slot(k)].  Input map n = (5*n1 + 4*n2) % 20, output map
// k = (5*k1 + 16*k2) % 20 (Good-Thomas, no twiddles).
KWS_HD void dft20_pfa(cpx (&v)[20]) {
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2)
    dft4(v[(4 * n2) % 20], v[(5 + 4 * n2) % 20], v[(10 + 4 * n2) % 20], v[(15 + 4 * n2) % 20]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    dft5(v[(5 * k1) % 20], v[(5 * k1 + 4) % 20], v[(5 * k1 + 8) % 20], v[(5 * k1 + 12) % 20],
         v[(5 * k1 + 16) % 20]);
}

KWS_HD constexpr int slot_of(int k) { return (k / kR) * kRowStride + (k % kR); }

// stage 1 + twiddle for thread j: columns n1 = j, j+10 of the pair window.
//   win  : this pair's samples; frame a = win[0..399], frame b = win[160..559]
//   tw   : tw[m] = exp(-2*pi*i*m/400), m < 400
//   buf  : exchange buffer, Y'[n1][k2] -> buf[n1*21 + k2]
KWS_HD void stage1(int j, const float* win, const cpx* tw, cpx* buf) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int n1 = j + 10 * c;
    cpx v[20];
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) {
      v[n2].re = win[n1 + 20 * n2];
      v[n2].im = win[160 + n1 + 20 * n2];
    }
    dft20_pfa(v);
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) buf[n1 * kRowStride + k2] = cmul(v[pfa_slot(k2)], tw[n1 * k2]);
  }
}

// stage 2 for thread j: columns k2 = j, j+10; in place (a thread only touches its own columns)
KWS_HD void stage2(int j, cpx* buf) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int k2 = j + 10 * c;
    cpx v[20];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) v[n1] = buf[n1 * kRowStride + k2];
    dft20_pfa(v);
#pragma unroll
    for (int k1 = 0; k1 < 20; ++k1) buf[k1 * kRowStride + k2] = v[pfa_slot(k1)];   // Z[20*k1 + k2]
  }
}

// |A[k]| and |B[k]| for the bins k = j, j+10, ... <= 200 owned by thread j.
// mag_a / mag_b get up to 21 values each (index i <-> bin j + 10*i).
KWS_HD int untangle(int j, const cpx* buf, float (&mag_a)[21], float (&mag_b)[21]) {
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 21; ++i) {
    const int k = j + 10 * i;
    if (k <= 200) {
      const cpx zk = buf[slot_of(k)];
      const cpx zm = buf[slot_of((kN - k) % kN)];
      const float ar = zk.re + zm.re, ai = zk.im - zm.im;   // 2*A[k]
      const float br = zk.re - zm.re, bi = zk.im + zm.im;   // 2i*B[k] rotated: same modulus
#if defined(__CUDA_ARCH__)
      mag_a[i] = 0.5f * sqrtf(ar * ar + ai * ai);
      mag_b[i] = 0.5f * sqrtf(br * br + bi * bi);
#else
      mag_a[i] = 0.5f * __builtin_sqrtf(ar * ar + ai * ai);
      mag_b[i] = 0.5f * __builtin_sqrtf(br * br + bi * bi);
#endif
      cnt = i + 1;
    }
  }
  return cnt;
}

}  // namespace fft
}  // namespace kws
