/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of
 * PositionalEncodingOp::Compute, positional_encoding/positional_encoding_op.cc:44-49.
 * Pinned against the unmodified reference kernel in oracle/_ref. */
#include <math.h>

void oracle_positional_encoding(int max_position, int encoding_size, float *out)
{
    for (int p = 0; p < max_position; ++p) {
        float *row = out + (long)p * encoding_size;
        for (int i = 0; i < encoding_size / 2; ++i) {
            double angle = p / pow(10000.0, 2.0 * i / encoding_size);
            row[2 * i] = (float)sin(angle);
            row[2 * i + 1] = (float)cos(angle);
        }
    }
}
