// 400-point DFT of two real frames at once, one DFT-20 column per thread, 20 threads per frame pair.
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
// registers.  Thread c of the pair's 20 owns column n1 = c in stage 1 and column
// k2 = c in stage 2; the single exchange goes through a [20][21]-slot complex buffer
// (row stride 21: a half-warp's 64-bit accesses land on 16 distinct bank pairs in both
// directions).  The thread that ends stage 2 with Z[20*k1+k2] in registers untangles its
// own bins k <= 200 and only needs Z[400-k] from its mirror column, so stage 2 publishes
// just rows 10..19.
//
// Everything here is __host__ __device__ so the index logic is unit-tested on the
// CPU (tests/test_host.py builds tests/host/fft400_host.cpp with g++).
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
constexpr int kR = 20;            // 400 = kR * kR; also the threads cooperating on one frame pair
constexpr int kRowStride = 21;    // padded slots per row of the exchange buffer
constexpr int kBufSlots = kR * kRowStride;   // 420 complex slots
constexpr int kPairWindow = 560;  // samples spanned by two consecutive frames (160 + 400)
constexpr int kBins = kN / 2 + 1; // 201

// Complex add / subtract / real scaling / rotation by +-i are pairwise operations, so on the device each is ONE
// packed f32x2 instruction (FADD2 / FFMA2 / FMUL2 of sm_100a; the half swap of the +-i rotations is a free
// register-half selector of the packed operand) instead of two scalar ones -- K1 is bound by instruction issue.
#if defined(__CUDA_ARCH__)
KWS_HD float2 as_f2(cpx a) { return make_float2(a.re, a.im); }
KWS_HD cpx as_cpx(float2 a) { return cpx{a.x, a.y}; }
KWS_HD cpx cadd(cpx a, cpx b) { return as_cpx(__fadd2_rn(as_f2(a), as_f2(b))); }
KWS_HD cpx csub(cpx a, cpx b) { return as_cpx(__ffma2_rn(as_f2(b), make_float2(-1.0f, -1.0f), as_f2(a))); }
KWS_HD cpx caxpy(float s, cpx b, cpx a) { return as_cpx(__ffma2_rn(as_f2(b), make_float2(s, s), as_f2(a))); }   // a + s*b
KWS_HD cpx cscale(float s, cpx b) { return as_cpx(__fmul2_rn(as_f2(b), make_float2(s, s))); }                     // s*b
KWS_HD cpx csub_i(cpx a, cpx b) { return as_cpx(__ffma2_rn(make_float2(b.im, b.re), make_float2(1.0f, -1.0f), as_f2(a))); }  // a - i*b
KWS_HD cpx cadd_i(cpx a, cpx b) { return as_cpx(__ffma2_rn(make_float2(b.im, b.re), make_float2(-1.0f, 1.0f), as_f2(a))); }  // a + i*b
// (a.re^2 + b.im^2, a.im^2 + b.re^2)
KWS_HD cpx cross_sq(cpx a, cpx b) {
  const float2 bs = make_float2(b.im, b.re);
  return as_cpx(__ffma2_rn(bs, bs, __fmul2_rn(as_f2(a), as_f2(a))));
}
#else
KWS_HD cpx cadd(cpx a, cpx b) { return cpx{a.re + b.re, a.im + b.im}; }
KWS_HD cpx csub(cpx a, cpx b) { return cpx{a.re - b.re, a.im - b.im}; }
KWS_HD cpx caxpy(float s, cpx b, cpx a) { return cpx{a.re + s * b.re, a.im + s * b.im}; }
KWS_HD cpx cscale(float s, cpx b) { return cpx{s * b.re, s * b.im}; }
KWS_HD cpx csub_i(cpx a, cpx b) { return cpx{a.re + b.im, a.im - b.re}; }
KWS_HD cpx cadd_i(cpx a, cpx b) { return cpx{a.re - b.im, a.im + b.re}; }
KWS_HD cpx cross_sq(cpx a, cpx b) { return cpx{a.re * a.re + b.im * b.im, a.im * a.im + b.re * b.re}; }
#endif
KWS_HD cpx cmul(cpx a, cpx b) { return cpx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// forward DFT-4, in place (W4 = -i)
KWS_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
  const cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = csub_i(t1, t3);
  a3 = cadd_i(t1, t3);
}

// forward DFT-5, in place (W5 = exp(-2*pi*i/5))
KWS_HD void dft5(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4) {
  const float c1 = 0.30901699437494742f;    // cos(2pi/5)
  const float c2 = -0.80901699437494742f;   // cos(4pi/5)
  const float s1 = 0.95105651629515357f;    // sin(2pi/5)
  const float s2 = 0.58778525229247313f;    // sin(4pi/5)
  const cpx t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const cpx m1 = caxpy(c2, t2, caxpy(c1, t1, a0));
  const cpx m2 = caxpy(c1, t2, caxpy(c2, t1, a0));
  const cpx n1 = caxpy(s2, t4, cscale(s1, t3));
  const cpx n2 = caxpy(-s1, t4, cscale(s2, t3));
  a0 = cadd(cadd(a0, t1), t2);
  a1 = csub_i(m1, n1);
  a4 = cadd_i(m1, n1);
  a2 = csub_i(m2, n2);
  a3 = cadd_i(m2, n2);
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

// Twiddle table used by stage 1: k2-major rows of kTwStride = 52 slots, periodically extended in n1:
//   twp[k2*52 + j] = exp(-2*pi*i*(j % 20)*k2/400),  j < 52.
// Thread t of a CTA owns column n1 = t % 20; with j = (12*warp) % 20 + lane (== n1 mod 20) the 32 lanes of a
// warp -- which straddle two frame pairs -- read 32 CONSECUTIVE slots for a fixed k2, i.e. conflict-free.
constexpr int kTwStride = 52;
constexpr int kTwSlots = kR * kTwStride;       // 1040
KWS_HD constexpr int tw_base(int warp, int lane) { return (12 * warp) % kR + lane; }

// stage 1 + twiddle for column n1.  v[n2] = a[n1+20*n2] + i*b[n1+20*n2] on entry (filled by the caller).
//   tw  : the table above, already offset by this thread's tw_base
//   buf : exchange buffer, Y'[n1][k2] -> buf[n1*21 + k2]
KWS_HD void stage1_col(int n1, cpx (&v)[20], const cpx* tw, cpx* buf) {
  dft20_pfa(v);
  buf[n1 * kRowStride] = v[pfa_slot(0)];                                // W^0 = 1
#pragma unroll
  for (int k2 = 1; k2 < kR; ++k2) buf[n1 * kRowStride + k2] = cmul(v[pfa_slot(k2)], tw[k2 * kTwStride]);
}

// stage 2 for column k2 (after a barrier).  On return Z[20*k1 + k2] == v[pfa_slot(k1)]; rows 10..19 of the
// column are written back in place (a column is only ever touched by its owner until the next barrier).
KWS_HD void stage2_col(int k2, cpx* buf, cpx (&v)[20]) {
#pragma unroll
  for (int n1 = 0; n1 < kR; ++n1) v[n1] = buf[n1 * kRowStride + k2];
  dft20_pfa(v);
#pragma unroll
  for (int k1 = kR / 2; k1 < kR; ++k1) buf[k1 * kRowStride + k2] = v[pfa_slot(k1)];
}

KWS_HD float sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return __builtin_sqrtf(x);
#endif
}

// 2|A[k]|, 2|B[k]| for the bins k = 20*k1 + k2 <= 200 of column k2 (after a barrier).  `v` is the register
// state left by stage2_col; Z[400-k] comes from the mirror column's published rows 10..19.  Results go to
// mag_a[k1*k1_stride] and mag_b[k1*k1_stride]: the caller passes the address of bin k2 and the distance between
// bins 20 apart (any memory that nobody is reading during this phase).
// The factor 1/2 (and the sample scale: the DFT is linear, so int16 samples are transformed unscaled) is left
// to the consumer -- the mel projection applies it once per band instead of once per bin.
KWS_HD void untangle_col(int k2, const cpx (&v)[20], const cpx* buf, float* mag_a, float* mag_b, int k1_stride) {
  const int mcol = k2 == 0 ? 0 : kR - k2;          // column of Z[400 - k]
  const int mrow0 = k2 == 0 ? kR : kR - 1;         // its row is mrow0 - k1
#pragma unroll
  for (int k1 = 0; k1 < kR / 2; ++k1) {
    const cpx zk = v[pfa_slot(k1)];
    int mrow = mrow0 - k1;
    const bool self = mrow == kR;                   // k == 0: Z[400] = Z[0] = zk
    if (self) mrow = kR - 1;                        // any published slot; the value is discarded
    cpx zm = buf[mrow * kRowStride + mcol];
    if (self) zm = zk;
    // 2*A[k] = (sp.re, sm.im);  2i*B[k] (same modulus as 2*B[k]) = (sm.re, sp.im)
    const cpx sp = cadd(zk, zm), sm = csub(zk, zm);
    const cpx sq = cross_sq(sp, sm);                // (|2A|^2, |2B|^2)
    mag_a[k1 * k1_stride] = sqrt_fast(sq.re);
    mag_b[k1 * k1_stride] = sqrt_fast(sq.im);
  }
  if (k2 == 0) {                                    // k = 200 = 400 - 200: its own mirror
    const cpx z = v[pfa_slot(kR / 2)];
    mag_a[(kR / 2) * k1_stride] = 2.0f * (z.re < 0.0f ? -z.re : z.re);
    mag_b[(kR / 2) * k1_stride] = 2.0f * (z.im < 0.0f ? -z.im : z.im);
  }
}

}  // namespace fft
}  // namespace kws
