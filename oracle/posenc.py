"""Oracle (TEST INFRASTRUCTURE ONLY): positional encoding table.

PINNED bit-for-bit against the unmodified reference kernel built into
oracle/_ref (tests/test_oracle_posenc.py).  Follows
``PositionalEncodingOp::Compute`` positional_encoding/positional_encoding_op.cc:32-50:
double-precision ``sin/cos(p / pow(10000, 2i/size))`` stored as float32; for odd
``encoding_size`` the last column is never written (left as ``fill``).
"""
from __future__ import annotations

import math

import numpy as np


def positional_encoding(max_position: int, encoding_size: int, fill: float = 0.0) -> np.ndarray:
    if encoding_size < 1:
        raise ValueError("encoding_size must be >= 1")      # Attr("encoding_size: int >= 1") :64
    out = np.full((max(0, int(max_position)), int(encoding_size)), fill, dtype=np.float32)
    for i in range(encoding_size // 2):
        denom = math.pow(10000.0, 2.0 * i / encoding_size)
        for p in range(out.shape[0]):
            a = p / denom
            out[p, 2 * i] = math.sin(a)
            out[p, 2 * i + 1] = math.cos(a)
    return out
