"""The weight-quantisation recipe of octbit/octbit_graph.py (the part of the graph
rewriter that defines what ``octbit_mat_mul`` consumes).  The GraphDef surgery itself
(:404-550) is TensorFlow-protobuf specific and out of scope.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib, _tensors


def octize_weight_int8_signed(weight, device=None):
    """octbit/octbit_graph.py:191-215 on the device.

    weight ``[in, out]`` float32 -> ``(W_q^T int8 [out, in], scale float, bias float32 [out])``
    with ``scale = max|W|/127``, ``W_q = round_half_even(W/scale)``, ``bias[j] = 127*sum_i W_q[i,j]``.
    """
    lib = _lib.load()
    host = _tensors.is_host(weight)
    dev = _tensors.require_cuda(device if device is not None or host else weight.device)
    w = _tensors.to_device(weight, torch.float32, dev)
    if w.dim() != 2:
        raise _lib.InvalidArgumentError("weight must be [in, out]")
    in_dim, out_dim = w.shape
    wq_t = torch.empty((out_dim, in_dim), dtype=torch.int8, device=dev)
    bias = torch.empty((out_dim,), dtype=torch.float32, device=dev)
    scale = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        _lib.check(lib.kws_octize_weight(_tensors.ptr(w), in_dim, out_dim, _tensors.ptr(wq_t), _tensors.ptr(bias),
                                         ctypes.byref(scale), _tensors.stream_ptr(dev)))
    if host:
        torch.cuda.current_stream(dev).synchronize()
        return _tensors.to_host(wq_t), float(scale.value), _tensors.to_host(bias)
    return wq_t, float(scale.value), bias


def default_octbit_matmul_name_check(name):
    """octbit/octbit_graph.py:218-225: layer-1 GRU matmuls only (layer 0 has K = n_mel+128,
    not a multiple of 64, which the op rejects)."""
    return name != "model/linear/linear/MatMul" and "MatMul" in name and "cell_0" not in name
