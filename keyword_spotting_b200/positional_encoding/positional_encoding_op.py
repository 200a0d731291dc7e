"""Drop-in for positional_encoding/positional_encoding_op.py.

The reference compiles positional_encoding_op.cc at import and exposes
``positional_encoding(max_position, encoding_size)`` (:14-24); here the table is
filled by ``kws_positional_encoding`` (K6) and returned as a CUDA tensor
``[max_position, encoding_size]`` float32 (``as_numpy=True`` for a host array).
"""
from __future__ import annotations

import torch

from .. import _lib, _tensors

_positional_encoding_module = _lib.load()


def positional_encoding(max_position, encoding_size, device=None, as_numpy=False, fill=0.0):
    dev = _tensors.require_cuda(device)
    max_position = int(max_position)
    encoding_size = int(encoding_size)
    if encoding_size < 1:
        raise _lib.InvalidArgumentError("encoding_size must be >= 1")       # Attr("encoding_size: int >= 1")
    if max_position < 0:
        raise _lib.InvalidArgumentError("max_position must be >= 0")
    # odd sizes: the reference never writes the last column (positional_encoding_op.cc:45);
    # `fill` is what the caller sees there.
    out = torch.full((max_position, encoding_size), float(fill), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_positional_encoding_module.kws_positional_encoding(max_position, encoding_size, _tensors.ptr(out),
                                                                      _tensors.stream_ptr(dev)))
    if as_numpy:
        torch.cuda.current_stream(dev).synchronize()
        return _tensors.to_host(out)
    return out
