"""Oracle (TEST INFRASTRUCTURE ONLY): restatement of utils/prediction.py decoders.

PINNED against golden vectors produced by importing the reference's own
utils/prediction.py (tests/golden/make_golden.py -> tests/golden/decode_*.npz).

Threshold comparisons are done in float64 on the (fp32) probabilities: the
reference ran under numpy 1.x, where ``np.float32 > python_float`` promotes to
double.  numpy 2 would compare in fp32; we pin the 2017 rule.
"""
from __future__ import annotations

import numpy as np

MODE_CTC_DECODE = 0         # utils/prediction.py:18-62   (offline / validation)
MODE_CTC_DECODE2 = 1        # utils/prediction.py:65-86   (streaming, detector.py:200)
MODE_CTC_DECODE_STRICT = 2  # utils/prediction.py:89-108


def _interleave(labels):
    """``[0, l1, 0, l2, 0, ...]`` int32 -- utils/prediction.py:58-62."""
    out = np.zeros(2 * len(labels) + 1, dtype=np.int32)
    out[1::2] = labels
    return out


def _word_cols(softmax, classnum):
    return np.asarray(softmax)[:, 1:classnum - 1].astype(np.float64)


def ctc_decode(softmax, lockout=3, thres=0.5, loose_thres=0.2):
    """utils/prediction.py:18-62.  Peak picking on columns 1..4 with a lockout
    and a 'loose' mode entered after the labels 1,2,3 were seen in a row."""
    p = np.asarray(softmax)[:, 1:5].astype(np.float64)
    T = p.shape[0]
    labels, frames = [], []
    loose = False
    i = 0
    while i < T:
        row = p[i]
        top = row.max() if row.size else 0.0
        if loose:
            if top < loose_thres:
                if labels[-1] != 3:                      # :31-34
                    i += lockout
                    loose = False
                    continue
            elif row[2] > loose_thres:                   # :36-40  (label 3 == column 2)
                labels.append(3)
                frames.append(i)
                i += lockout
                loose = False
                continue
            else:                                        # :42-45
                k = int(row.argmax())
                if row[k] > 0.6 and frames[-1] + lockout < i:
                    labels.append(k + 1)
                    frames.append(i)
        elif top > thres:                                # :48-55
            labels.append(int(row.argmax()) + 1)
            frames.append(i)
            i += lockout
            if labels[-3:] == [1, 2, 3]:
                loose = True
            continue
        i += 1
    return _interleave(labels)


def ctc_decode2(softmax, classnum, thres=0.4):
    """utils/prediction.py:65-86.  Emit ``argmax+1`` whenever the per-frame
    winner over columns 1..classnum-2 exceeds ``thres`` and differs from the
    previous above-threshold winner; a sub-threshold frame forgets the winner."""
    p = _word_cols(softmax, classnum)
    labels = []
    prev = -1
    for row in p:
        if row.size and row.max() > thres:
            k = int(row.argmax())
            if prev == -1 or prev != k:
                labels.append(k + 1)
            prev = k
        else:
            prev = -1
    return _interleave(labels)


def ctc_decode_strict(softmax, classnum, lockout=3, thres=0.5):
    """utils/prediction.py:89-108.  Threshold peak + lockout, nothing else."""
    p = _word_cols(softmax, classnum)
    T = p.shape[0]
    labels = []
    i = 0
    while i < T:
        row = p[i]
        if row.size and row.max() > thres:
            labels.append(int(row.argmax()) + 1)
            i += lockout
        else:
            i += 1
    return _interleave(labels)


def ctc_predict(seq, label="1233"):
    """utils/prediction.py:111-118.  Decimal-concatenate the positive entries
    up to the first negative one; 1 iff ``label`` occurs as a substring."""
    digits = []
    for v in seq:
        v = int(v)
        if v < 0:
            break
        if v > 0:
            digits.append(str(v))
    return 1 if label in "".join(digits) else 0


def evaluate(result, target):
    """utils/prediction.py:203-210 -> (miss, number of targets, false_accept)."""
    result = np.asarray(result, dtype=np.int64)
    target = np.asarray(target, dtype=np.int64)
    assert len(result) == len(target)
    xor = result ^ target
    miss = int((xor & target).sum())
    false_accept = int((xor & result).sum())
    return miss, int(target.sum()), false_accept


def decode(softmax, mode, classnum=6, lockout=3, thres=None, loose_thres=0.2):
    """Dispatch by ``mode`` with the reference defaults."""
    if mode == MODE_CTC_DECODE:
        return ctc_decode(softmax, lockout, 0.5 if thres is None else thres, loose_thres)
    if mode == MODE_CTC_DECODE2:
        return ctc_decode2(softmax, classnum, 0.4 if thres is None else thres)
    if mode == MODE_CTC_DECODE_STRICT:
        return ctc_decode_strict(softmax, classnum, lockout, 0.5 if thres is None else thres)
    raise ValueError("unknown decode mode %r" % (mode,))
