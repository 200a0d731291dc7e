"""utils/wer.py of the reference (Levenshtein WER, WERCalculator) over the device edit-distance kernel.

Same names and return conventions as the reference: ``wer(r, h)`` -> float distance / len(r) (the raw distance when
``r`` is empty, utils/wer.py:36-41), ``batch_wer`` over sparse (index, value) pairs (:44-77), ``WERCalculator``
with ``remove_residual`` / ``cal_batch_wer`` / ``cal_topk_wers`` (:80-124).  Sequences of up to 254 labels, as in
the reference (its DP table is uint8).  No CPU fallback: the distances come from ``kws_edit_distance``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, _tensors

MAX_LEN = 254


def edit_distance_batch(refs, hyps, device=None):
    """Lists of label sequences -> int32 numpy array of Levenshtein distances (one per pair)."""
    if len(refs) != len(hyps):
        raise _lib.InvalidArgumentError("refs and hyps must have the same length")
    S = len(refs)
    if S == 0:
        return np.zeros(0, np.int32)
    rl = np.asarray([len(r) for r in refs], np.int32)
    hl = np.asarray([len(h) for h in hyps], np.int32)
    max_len = int(max(rl.max(), hl.max(), 1))
    if max_len > MAX_LEN:
        raise _lib.InvalidArgumentError("sequences longer than %d labels are outside the reference's domain" % MAX_LEN)
    R = np.zeros((S, max_len), np.int32)
    H = np.zeros((S, max_len), np.int32)
    for i, (r, h) in enumerate(zip(refs, hyps)):
        R[i, :len(r)] = np.asarray(r, np.int64)
        H[i, :len(h)] = np.asarray(h, np.int64)
    dev = _tensors.require_cuda(device)
    lib = _lib.load()
    t = [_tensors.to_device(a, torch.int32, dev) for a in (R, rl, H, hl)]
    out = torch.empty(S, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.kws_edit_distance(_tensors.ptr(t[0]), _tensors.ptr(t[1]), max_len, _tensors.ptr(t[2]),
                                         _tensors.ptr(t[3]), max_len, S, max_len, _tensors.ptr(out),
                                         _tensors.stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()
    return _tensors.to_host(out)


def wer(r, h):
    """utils/wer.py:4-41."""
    d = float(edit_distance_batch([list(r)], [list(h)])[0])
    return d if len(r) == 0 else d / float(len(r))


def batch_wer(batch_size, r_index, r_value, h_index, h_value):
    """utils/wer.py:44-77: mean WER of a batch given as sparse (index, value) arrays."""
    refs = [[] for _ in range(batch_size)]
    hyps = [[] for _ in range(batch_size)]
    for idx, v in zip(r_index, r_value):
        if 0 <= idx[0] < batch_size:
            refs[idx[0]].append(v)
    for idx, v in zip(h_index, h_value):
        if 0 <= idx[0] < batch_size:
            hyps[idx[0]].append(v)
    d = edit_distance_batch(refs, hyps).astype(np.float64)
    lens = np.asarray([len(r) for r in refs], np.float64)
    return np.mean(np.where(lens > 0, d / np.maximum(lens, 1.0), d))


class WERCalculator(object):
    """utils/wer.py:80-124."""

    def __init__(self, ignore_label_list):
        self._ignore_label_set = set(ignore_label_list)

    def remove_residual(self, inputs):
        """Labels up to (not including) the first -1 terminator, without the ignored ones (utils/wer.py:84-92)."""
        seq = np.asarray(inputs).ravel()
        stop = np.flatnonzero(seq == -1)
        if stop.size:
            seq = seq[:stop[0]]
        if self._ignore_label_set:
            seq = seq[~np.isin(seq, list(self._ignore_label_set))]
        return seq

    def cal_batch_wer(self, batch_r, batch_h):
        refs = [self.remove_residual(r) for r in batch_r]
        hyps = [self.remove_residual(h) if len(r) else np.asarray([]) for r, h in zip(refs, batch_h)]
        d = edit_distance_batch([list(r) for r in refs], [list(h) for h in hyps]).astype(np.float64)
        lens = np.asarray([len(r) for r in refs], np.float64)
        return np.where(lens > 0, d / np.maximum(lens, 1.0), 0.0)          # empty reference -> 0. (:97-99)

    def cal_topk_wers(self, batch_r, batch_h, batch_size, nums_gpu, topk, max_topk):
        """Best-of-topk WER per utterance for hypotheses laid out [gpu][k][utterance] (utils/wer.py:108-124)."""
        best = []
        for g in range(int(nums_gpu)):
            refs = batch_r[g * batch_size:(g + 1) * batch_size]
            hyp0 = g * batch_size * max_topk
            per_k = [self.cal_batch_wer(refs, batch_h[hyp0 + k * batch_size: hyp0 + (k + 1) * batch_size])
                     for k in range(int(topk))]
            best.extend(np.min(np.vstack(per_k), axis=0))
        return best
