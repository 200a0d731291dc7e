"""Drop-in for utils/prediction.py: same function names, arguments and return values,
computed by the device decoder (kws_ctc_decode, K4).

``ctc_decode`` / ``ctc_decode2`` / ``ctc_decode_strict`` take one ``[T, C]`` softmax
and return the reference's ``np.int32`` array ``[0, l1, 0, l2, 0, ...]``
(utils/prediction.py:58-62).  ``decode_batch`` is the batched form the B200 path is
meant for.  ``ctc_predict`` and ``evaluate`` are integer bookkeeping on the already
decoded labels (utils/prediction.py:111-118, 203-210) and stay on the host.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib, _tensors

MODE_CTC_DECODE = _lib.DECODE_CTC
MODE_CTC_DECODE2 = _lib.DECODE_CTC2
MODE_CTC_DECODE_STRICT = _lib.DECODE_STRICT


def decode_batch(softmax, lens=None, mode=MODE_CTC_DECODE, lockout=3, thres=None, loose_thres=0.2,
                 keyword="1233", max_labels=None, device=None, want_labels=True):
    """softmax ``[S, T, C]`` -> (labels ``[S, max_labels]`` padded with -1, counts ``[S]``, trigger ``[S]``)."""
    lib = _lib.load()
    host = _tensors.is_host(softmax)
    dev = _tensors.require_cuda(device if device is not None or host else softmax.device)
    p = _tensors.to_device(softmax, torch.float32, dev)
    if p.dim() != 3:
        raise _lib.InvalidArgumentError("softmax must be [S, T, C]")
    S, T, C = p.shape
    if max_labels is None:
        max_labels = 2 * T + 1
    ln = None if lens is None else _tensors.to_device(lens, torch.int32, dev)
    labels = torch.empty((S, max_labels), dtype=torch.int32, device=dev) if want_labels else None
    counts = torch.empty((S,), dtype=torch.int32, device=dev)
    trig = torch.empty((S,), dtype=torch.int32, device=dev)
    prm = _lib.DecodeParams(int(mode), int(lockout), -1.0 if thres is None else float(thres), float(loose_thres))
    with torch.cuda.device(dev):
        _lib.check(lib.kws_ctc_decode(_tensors.ptr(p), S, T, C, _tensors.ptr(ln), ctypes.byref(prm),
                                      keyword.encode("ascii"), _tensors.ptr(labels), int(max_labels),
                                      _tensors.ptr(counts), _tensors.ptr(trig), _tensors.stream_ptr(dev)))
    if host:
        torch.cuda.current_stream(dev).synchronize()
        return (None if labels is None else _tensors.to_host(labels)), _tensors.to_host(counts), _tensors.to_host(trig)
    return labels, counts, trig


def _single(softmax, mode, **kw):
    sm = np.asarray(softmax, dtype=np.float32)
    if sm.ndim != 2:
        raise _lib.InvalidArgumentError("softmax must be [T, C]")
    if sm.shape[0] == 0:
        return np.asarray([0], dtype=np.int32)
    labels, counts, _ = decode_batch(sm[None], mode=mode, **kw)
    return np.ascontiguousarray(labels[0, :counts[0]], dtype=np.int32)


def ctc_decode(softmax, lockout=3, thres=0.5, loose_thres=0.2):
    """utils/prediction.py:18-62."""
    return _single(softmax, MODE_CTC_DECODE, lockout=lockout, thres=thres, loose_thres=loose_thres)


def _label_columns(softmax, classnum):
    """The reference slices ``softmax[:, 1:classnum-1]`` (utils/prediction.py:67,91); numpy clamps the end to the
    actual column count, so ``classnum`` larger than the array keeps the LAST column too.  The device decoder
    reads columns ``1..C-2`` of what it is given: hand it the kept columns plus one unused closing column."""
    sm = np.asarray(softmax, dtype=np.float32)
    if sm.ndim != 2:
        raise _lib.InvalidArgumentError("softmax must be [T, C]")
    end = max(1, min(int(classnum) - 1, sm.shape[1]))
    return np.concatenate([sm[:, :end], np.zeros((sm.shape[0], 1), np.float32)], axis=1)


def ctc_decode2(softmax, classnum, thres=0.4):
    """utils/prediction.py:65-86 (the streaming decoder of detector.py:200)."""
    return _single(_label_columns(softmax, classnum), MODE_CTC_DECODE2, thres=thres)


def ctc_decode_strict(softmax, classnum, lockout=3, thres=0.5):
    """utils/prediction.py:89-108."""
    return _single(_label_columns(softmax, classnum), MODE_CTC_DECODE_STRICT, lockout=lockout, thres=thres)


def ctc_predict(seq, label="1233"):
    """utils/prediction.py:111-118."""
    text = ""
    for v in seq:
        if v < 0:
            break
        if v > 0:
            text += str(int(v))
    return 1 if label in text else 0


def evaluate(result, target):
    """utils/prediction.py:203-210 -> (miss, number of targets, false accepts)."""
    assert len(result) == len(target)
    xor = [int(a) ^ int(b) for a, b in zip(target, result)]
    miss = sum(a & int(b) for a, b in zip(xor, target))
    false_accept = sum(a & int(b) for a, b in zip(xor, result))
    return miss, sum(int(t) for t in target), false_accept
