"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's WER (utils/wer.py:4-41).

Levenshtein distance with a uint8 table (the reference's dtype: only sequences up to 254 labels are in its
domain), normalised by len(r); the raw distance when r is empty.  Pinned against the reference's own function
in tests/golden/wer_golden.npz (tests/golden/make_golden.py).
"""
import numpy as np


def edit_distance(r, h) -> int:
    d = np.zeros((len(r) + 1, len(h) + 1), dtype=np.uint8)
    d[:, 0] = np.arange(len(r) + 1)
    d[0, :] = np.arange(len(h) + 1)
    for i in range(1, len(r) + 1):
        for j in range(1, len(h) + 1):
            if r[i - 1] == h[j - 1]:
                d[i, j] = d[i - 1, j - 1]
            else:
                d[i, j] = min(int(d[i - 1, j - 1]) + 1, int(d[i, j - 1]) + 1, int(d[i - 1, j]) + 1)
    return int(d[len(r), len(h)])


def wer(r, h) -> float:
    d = float(edit_distance(r, h))
    return d if len(r) == 0 else d / float(len(r))
