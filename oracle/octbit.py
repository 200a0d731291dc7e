"""Oracle (TEST INFRASTRUCTURE ONLY): octbit 8-bit matmul, restated in numpy.

PINNED: reproduces octbit/octbit_ops_test.py:24-34 (-> [[-6048.]]) and :41-53
(-> [[-128,-4032,-4032,-4032]]x2) and is compared bit-for-bit with the
unmodified reference kernel built into oracle/_ref (tests/test_oracle_octbit.py).

Follows ``OctbitMatMulOp::Compute`` octbit/octbit_mat_mul_op.cc:49-183 and the
weight recipe ``octize_weight_int8_signed`` octbit/octbit_graph.py:191-215.
"""
from __future__ import annotations

import numpy as np


class InvalidArgument(ValueError):
    """Stands for TF's ``errors::InvalidArgument`` raised by the op."""


def _round_half_away(v64: np.ndarray) -> np.ndarray:
    """C ``round()`` (octbit_mat_mul_op.cc:112,121) on float64 values."""
    return np.sign(v64) * np.floor(np.abs(v64) + 0.5)


def quantize_activations(x: np.ndarray):
    """octbit_mat_mul_op.cc:90-124.  Tensor-wide dynamic range -> u8.

    Returns ``(q u8[A,K], bscale f32, signed_flag)``.
    signed   : bscale = max(-min, max)/127, q = round(x/bscale) + 127   (0..254)
    unsigned : bscale = max/254,            q = round(x/bscale)         (0..254)
    All-zero input (bscale == 0, 0/0) is defined here as q == 0; the reference
    leaves that cast undefined.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.size == 0:
        return np.zeros(x.shape, np.uint8), np.float32(0), False
    min_v = np.float32(x.min())
    max_v = np.float32(x.max())
    signed_flag = bool(min_v < 0)
    if signed_flag:
        bscale = np.float32(max(np.float32(-min_v), max_v)) / np.float32(127)
        offset = 127.0
    else:
        bscale = max_v / np.float32(254)
        offset = 0.0
    bscale = np.float32(bscale)
    if bscale == 0:
        return np.zeros(x.shape, np.uint8), bscale, signed_flag
    ratio = (x / bscale).astype(np.float32)          # fp32 IEEE division (:112,:121)
    q = _round_half_away(ratio.astype(np.float64)) + offset
    return q.astype(np.int64).astype(np.uint8), bscale, signed_flag


def lane_sums(q: np.ndarray, w: np.ndarray) -> np.ndarray:
    """octbit_mat_mul_op.cc:137-170: the four int32 lanes of ``sum``.

    ``_mm_maddubs_epi16`` forms, per 16-byte block, eight int16 values
    ``sat16(q[2p]*w[2p] + q[2p+1]*w[2p+1])``; lane m accumulates pairs p with
    ``p % 4 == m`` (lo = pairs 0..3, hi = pairs 4..7, added lane-wise).
    q ``[A,K]`` u8, w ``[B,K]`` i8 -> ``[A,B,4]`` int32 (wrapping adds).
    """
    A, K = q.shape
    B = w.shape[0]
    qi = q.astype(np.int32).reshape(A, 1, K // 2, 2)
    wi = w.astype(np.int32).reshape(1, B, K // 2, 2)
    pair = (qi * wi).sum(axis=3)
    pair = np.clip(pair, -32768, 32767)                # the int16 saturation
    lanes = pair.reshape(A, B, K // 8, 4).astype(np.int64).sum(axis=2)
    return ((lanes + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)


def octbit_mat_mul(x, w, transpose_a=False, transpose_b=True, scale=0.0, bias=(0,)):
    """Same signature/defaults as octbit/octbit_ops.py:17-26.

    x ``[A,K]`` float32, w ``[B,K]`` int8 (already transposed), bias ``[B]``.
    Checks mirror the op: octbit_mat_mul_op.cc:41-46 (attrs) and :61-73.
    """
    if not transpose_b:
        raise InvalidArgument("b need to be transposed")
    if transpose_a:
        raise InvalidArgument("a cannot to be transposed")
    if not (np.float32(scale) > 0):
        raise InvalidArgument("scale has to be positive")
    x = np.asarray(x, dtype=np.float32)
    w = np.asarray(w)
    if x.ndim != 2:
        raise InvalidArgument("In[0] is not a matrix")
    if w.ndim != 2:
        raise InvalidArgument("In[1] is not a matrix")
    if x.shape[1] != w.shape[1]:
        raise InvalidArgument("f is not equal in filter and input")
    if x.shape[1] % 64 != 0:
        raise InvalidArgument("we need to be 16 aligned.")
    w = w.astype(np.int8)
    bias = np.asarray(bias, dtype=np.float32).reshape(-1)
    A, K = x.shape
    B = w.shape[0]
    q, bscale, signed_flag = quantize_activations(x)
    total_scale = np.float32(np.float32(scale) * bscale)          # :108,:117
    out = np.zeros((A, B), dtype=np.float32)
    rows = max(1, (1 << 22) // max(1, B * K))
    for a0 in range(0, A, rows):
        lanes = lane_sums(q[a0:a0 + rows], w)
        acc = np.zeros(lanes.shape[:2], dtype=np.float32)
        for m in range(4):                                        # :172-175, fp32 adds in lane order
            acc = (acc + lanes[:, :, m].astype(np.float32)).astype(np.float32)
        if signed_flag:
            acc = (acc - bias[None, :B]).astype(np.float32)       # :176-178
        out[a0:a0 + rows] = (acc * total_scale).astype(np.float32)  # :179
    return out


def octize_weight_int8_signed(weight: np.ndarray):
    """octbit/octbit_graph.py:191-215.  W ``[in,out]`` float32 ->
    ``(W_q^T int8 [out,in], scale float64, bias float64 [out])`` with
    ``scale = max|W|/127``, ``W_q = np.round(W/scale)`` (half-to-even) and
    ``bias[j] = 127 * sum_i W_q[i,j]``.

    Promotion pinned to the numpy-1.x rules the reference ran under:
    ``nmax`` is an fp32 scalar, ``nmax / 127.`` is a float64 scalar, and
    ``tensor_value / scale`` (fp32 array / float64 scalar) is evaluated in fp32
    with ``scale`` cast to fp32 -- which is also the value the op receives,
    because the ``scale`` attr is a 32-bit float (octbit_ops_reg.cc:12)."""
    weight = np.asarray(weight, dtype=np.float32)
    nmax = max(abs(np.float32(weight.max())), abs(np.float32(weight.min())))
    scale = float(nmax) / 127.0
    wq = np.round((weight / np.float32(scale)).astype(np.float32))
    bias = (wq.astype(np.float64) * 127.0).sum(axis=0)
    return np.ascontiguousarray(wq.T).astype(np.int8), float(scale), bias


def default_octbit_matmul_name_check(name: str) -> bool:
    """octbit/octbit_graph.py:218-225: which MatMuls get converted."""
    return name != "model/linear/linear/MatMul" and "MatMul" in name and "cell_0" not in name
