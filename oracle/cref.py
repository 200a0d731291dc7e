"""Oracle (TEST INFRASTRUCTURE ONLY): ctypes loaders for the compiled checkers.

``load_c()``   -> oracle/_build/liboracle_c.so   (plain-C restatement, oracle/c)
``load_ref()`` -> oracle/_ref/libkws_ref_ops.so  (the reference's own kernels,
                  compiled unmodified from /root/reference behind oracle/tf_shim)

Both are built by ``make -C oracle [ref]`` (``__graft_entry__.build()`` runs it).
``/root/reference`` only exists in the build container; on the GPU box the
prebuilt ``.so`` files that travelled with the snapshot are used as they are.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
C_LIB = os.path.join(HERE, "_build", "liboracle_c.so")
REF_LIB = os.path.join(HERE, "_ref", "libkws_ref_ops.so")
REFERENCE_DIR = os.environ.get("KWS_REFERENCE_DIR", "/root/reference")


def build(ref: bool = True, quiet: bool = True) -> None:
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=out)
    if ref and os.path.isdir(REFERENCE_DIR):
        subprocess.check_call(["make", "-C", HERE, "ref", "REFERENCE=" + REFERENCE_DIR], stdout=out)


def load_c():
    if not os.path.exists(C_LIB):
        build(ref=False)
    return ctypes.CDLL(C_LIB)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def load_ref():
    if not have_ref():
        if not os.path.isdir(REFERENCE_DIR):
            raise FileNotFoundError("oracle/_ref is not built and %s is absent" % REFERENCE_DIR)
        build(ref=True)
    return ctypes.CDLL(REF_LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def c_octbit_matmul(x, w, bias, scale):
    lib = load_c()
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.int8)
    bias = np.ascontiguousarray(bias, np.float32)
    A, K = x.shape
    B = w.shape[0]
    out = np.zeros((A, B), np.float32)
    rc = lib.oracle_octbit_matmul(_p(x), _p(w), _p(bias), ctypes.c_float(scale), A, B, K, _p(out))
    if rc != 0:
        raise ValueError("oracle_octbit_matmul rc=%d" % rc)
    return out


def c_positional_encoding(max_position, encoding_size, fill=0.0):
    lib = load_c()
    out = np.full((max_position, encoding_size), fill, np.float32)
    lib.oracle_positional_encoding(int(max_position), int(encoding_size), _p(out))
    return out


class RefError(ValueError):
    pass


def ref_octbit_matmul(x, w, bias, scale, transpose_a=False, transpose_b=True):
    """Run the reference's own OctbitMatMulOp (octbit/octbit_mat_mul_op.cc)."""
    lib = load_ref()
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.int8)
    bias = np.ascontiguousarray(bias, np.float32)
    A, K = x.shape
    B, Kw = w.shape
    out = np.zeros((A, B), np.float32)
    err = ctypes.create_string_buffer(512)
    rc = lib.ref_octbit_matmul(_p(x), _p(w), _p(bias), ctypes.c_float(scale), int(transpose_a),
                               int(transpose_b), A, B, K, Kw, _p(out), err, 512)
    if rc != 0:
        raise RefError(err.value.decode())
    return out


def ref_positional_encoding(max_position, encoding_size, fill=0.0):
    """Run the reference's own PositionalEncodingOp (positional_encoding_op.cc)."""
    lib = load_ref()
    out = np.full((max_position, encoding_size), fill, np.float32)
    err = ctypes.create_string_buffer(512)
    rc = lib.ref_positional_encoding(int(max_position), int(encoding_size), _p(out), err, 512)
    if rc != 0:
        raise RefError(err.value.decode())
    return out
