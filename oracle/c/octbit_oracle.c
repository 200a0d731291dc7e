/* Oracle (TEST INFRASTRUCTURE ONLY): scalar plain-C restatement of
 * OctbitMatMulOp::Compute, octbit/octbit_mat_mul_op.cc:90-181 -- no intrinsics,
 * the maddubs pair saturation and the four-lane accumulation written out.
 * Pinned against octbit/octbit_ops_test.py:24-53 and the unmodified reference
 * kernel in oracle/_ref.  Built with -ffp-contract=off (the reference is
 * compiled without FMA: octbit/op_compile.py:64-72). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* :90-124  returns 1 if the signed branch was taken; *bscale_out = bscale */
int oracle_octbit_quantize(const float *x, long n, uint8_t *q, float *bscale_out)
{
    float min_v = 3.402823466e+38F, max_v = -3.402823466e+38F;
    for (long i = 0; i < n; ++i) {
        if (x[i] < min_v) min_v = x[i];
        if (x[i] > max_v) max_v = x[i];
    }
    int is_signed = min_v < 0;
    float bscale;
    if (is_signed) {
        float m = -min_v > max_v ? -min_v : max_v;
        bscale = m / 127;
    } else {
        bscale = max_v / 254;
    }
    *bscale_out = bscale;
    for (long i = 0; i < n; ++i) {
        if (bscale == 0) { q[i] = 0; continue; }       /* 0/0: undefined in the reference */
        double r = round((double)(x[i] / bscale));
        q[i] = (uint8_t)(int)(is_signed ? r + 127 : r);
    }
    return is_signed;
}

static inline int32_t sat16(int32_t v)
{
    return v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
}

/* x[A,K] f32, w[B,K] s8, bias[B] f32, attr scale -> out[A,B] f32 */
int oracle_octbit_matmul(const float *x, const int8_t *w, const float *bias, float scale_attr,
                         int A, int B, int K, float *out)
{
    if (K % 64 != 0) return -1;
    uint8_t *q = (uint8_t *)malloc((size_t)A * K + 1);
    float bscale;
    int is_signed = oracle_octbit_quantize(x, (long)A * K, q, &bscale);
    float scale = scale_attr * bscale;                                  /* :108,:117 */
    for (int i = 0; i < B; ++i) {
        const int8_t *wr = w + (long)i * K;
        for (int a = 0; a < A; ++a) {
            const uint8_t *qr = q + (long)a * K;
            uint32_t lane[4] = {0, 0, 0, 0};                            /* wrapping like paddd */
            for (int blk = 0; blk < K / 16; ++blk)
                for (int p = 0; p < 8; ++p) {
                    int k = blk * 16 + 2 * p;
                    int32_t s = sat16((int32_t)qr[k] * wr[k] + (int32_t)qr[k + 1] * wr[k + 1]);
                    lane[p & 3] += (uint32_t)s;
                }
            float o = 0.0f;
            for (int m = 0; m < 4; ++m) o += (float)(int32_t)lane[m];   /* :172-175 */
            if (is_signed) o -= bias[i];                                /* :176-178 */
            o *= scale;                                                 /* :179 */
            out[(long)a * B + i] = o;
        }
    }
    free(q);
    return 0;
}
