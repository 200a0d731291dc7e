"""Drop-in for octbit/octbit_ops.py: ``octbit_mat_mul`` with the reference signature.

The reference compiles octbit_mat_mul_op.cc at import and loads it as a TF op
(octbit/octbit_ops.py:8-13); here the op is ``kws_octbit_matmul`` in
libkws_b200.so (built at import by ``_build``), an int8 tensor-core kernel that is
bit-exact with the CPU op.  Errors the op raised as ``InvalidArgument``
(octbit_mat_mul_op.cc:41-46,56-73) surface as ``InvalidArgumentError`` (a
``ValueError``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, _tensors

_octbit_ops_so = _lib.load()
assert _octbit_ops_so, "Could not load libkws_b200.so."

_workspaces = {}


def _workspace(device, nbytes):
    ws = _workspaces.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[device] = ws
    return ws


def _prepare(x1, x2, bias, device):
    host = _tensors.is_host(x1)
    dev = _tensors.require_cuda(device if device is not None or host else x1.device)
    x = _tensors.to_device(x1, torch.float32, dev)
    w = x2 if isinstance(x2, torch.Tensor) else np.asarray(x2)
    if not isinstance(w, torch.Tensor) and w.dtype.names:        # np.dtype([("qint8", np.int8, 1)]) of the reference test
        w = w[w.dtype.names[0]].reshape(w.shape)
    w = _tensors.to_device(w, torch.int8, dev)
    b = _tensors.to_device(np.asarray(bias, np.float32).reshape(-1) if not isinstance(bias, torch.Tensor) else bias.reshape(-1),
                           torch.float32, dev)
    if x.dim() != 2:
        raise _lib.InvalidArgumentError("In[0] is not a matrix")          # octbit_mat_mul_op.cc:70-71
    if w.dim() != 2:
        raise _lib.InvalidArgumentError("In[1] is not a matrix")          # :72-73
    if x.shape[1] != w.shape[1]:
        raise _lib.InvalidArgumentError("f is not equal in filter and input")   # :61-63
    if b.numel() < w.shape[0]:
        raise _lib.InvalidArgumentError("bias has %d entries, need %d" % (b.numel(), w.shape[0]))
    return host, dev, x, w, b


def octbit_mat_mul(x1, x2, transpose_a=False, transpose_b=True, scale=0.0, bias=[0], device=None,
                   _exact=False):
    """octbit/octbit_ops.py:17-26.  x1 ``[A,K]`` float, x2 ``[B,K]`` int8 -> ``[A,B]`` float32."""
    host, dev, x, w, b = _prepare(x1, x2, bias, device)
    A, K = x.shape
    B = w.shape[0]
    out = torch.empty((A, B), dtype=torch.float32, device=dev)
    nbytes = _octbit_ops_so.kws_octbit_workspace_bytes(A, K)
    ws = _workspace(dev, nbytes)
    with torch.cuda.device(dev):
        if _exact:
            if not transpose_b or transpose_a:
                raise _lib.InvalidArgumentError("b need to be transposed")
            rc = _octbit_ops_so.kws_octbit_matmul_exact(_tensors.ptr(x), _tensors.ptr(w), _tensors.ptr(b), float(scale),
                                                        A, B, K, _tensors.ptr(out), _tensors.ptr(ws), ws.numel(),
                                                        _tensors.stream_ptr(dev))
        else:
            rc = _octbit_ops_so.kws_octbit_matmul(_tensors.ptr(x), _tensors.ptr(w), _tensors.ptr(b), float(scale),
                                                  int(bool(transpose_a)), int(bool(transpose_b)), A, B, K,
                                                  _tensors.ptr(out), _tensors.ptr(ws), ws.numel(),
                                                  _tensors.stream_ptr(dev))
        _lib.check(rc)
    if host:
        torch.cuda.current_stream(dev).synchronize()
        return _tensors.to_host(out)
    return out
