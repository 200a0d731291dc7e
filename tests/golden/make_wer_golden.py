#!/usr/bin/env python
"""Golden vectors for the WER path, produced by the REFERENCE's own utils/wer.py (build container only).

    python tests/golden/make_wer_golden.py

Writes tests/golden/wer_golden.npz: ragged (reference, hypothesis) label pairs, the reference's wer() values,
WERCalculator.cal_batch_wer() with ignore labels and -1 terminators, and batch_wer() on sparse inputs.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("KWS_REFERENCE_DIR", "/root/reference")
sys.dont_write_bytecode = True


def load_ref_module(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pack(seqs):
    off = np.zeros(len(seqs) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.concatenate([np.asarray(s, np.int32) for s in seqs]) if off[-1] else np.zeros(0, np.int32)
    return flat.astype(np.int32), off


def main():
    ref = load_ref_module("utils/wer.py", "ref_wer")
    rng = np.random.default_rng(4242)
    refs, hyps = [], []
    for n in range(160):
        lr = int(rng.choice([0, 1, 2, 5, 9, 30, 120, 254]))
        r = rng.integers(1, 6, lr).tolist()
        kind = n % 4
        if kind == 0:                                   # unrelated
            h = rng.integers(1, 6, int(rng.integers(0, 60))).tolist()
        elif kind == 1:                                 # edited copy
            h = list(r)
            for _ in range(int(rng.integers(0, 8))):
                op = rng.integers(0, 3)
                pos = int(rng.integers(0, len(h) + 1))
                if op == 0 and len(h) < 254:
                    h.insert(pos, int(rng.integers(1, 6)))
                elif op == 1 and h:
                    h.pop(min(pos, len(h) - 1))
                elif h:
                    h[min(pos, len(h) - 1)] = int(rng.integers(1, 6))
        elif kind == 2:                                 # identical
            h = list(r)
        else:                                           # empty hypothesis
            h = []
        refs.append(r)
        hyps.append(h[:254])
    wers = np.asarray([ref.wer(r, h) for r, h in zip(refs, hyps)], np.float64)
    # WERCalculator with ignore labels {0, 5} and -1 terminated, padded rows
    calc = ref.WERCalculator([0, 5])
    B, L = 24, 40
    br = rng.integers(-1, 6, (B, L))
    bh = rng.integers(-1, 6, (B, L))
    br[:, 0] = rng.integers(0, 6, B)
    br[3, 0] = -1                                       # empty reference row -> 0.
    batch = np.asarray(calc.cal_batch_wer(br, bh), np.float64)
    # batch_wer on sparse inputs
    bs = 6
    r_index, r_value, h_index, h_value = [], [], [], []
    for b in range(bs):
        for t in range(int(rng.integers(1, 9))):
            r_index.append([b, t]); r_value.append(int(rng.integers(1, 6)))
        for t in range(int(rng.integers(0, 9))):
            h_index.append([b, t]); h_value.append(int(rng.integers(1, 6)))
    bw = float(ref.batch_wer(bs, np.asarray(r_index), np.asarray(r_value), np.asarray(h_index), np.asarray(h_value)))
    rf, ro = pack(refs)
    hf, ho = pack(hyps)
    np.savez_compressed(os.path.join(HERE, "wer_golden.npz"), ref=rf, ref_off=ro, hyp=hf, hyp_off=ho, wer=wers,
                        calc_r=br.astype(np.int32), calc_h=bh.astype(np.int32), calc_wer=batch,
                        sp_bs=np.int32(bs), sp_r_index=np.asarray(r_index, np.int32), sp_r_value=np.asarray(r_value, np.int32),
                        sp_h_index=np.asarray(h_index, np.int32), sp_h_value=np.asarray(h_value, np.int32), sp_wer=np.float64(bw))
    print("wer_golden.npz:", len(refs), "pairs; mean wer", wers.mean(), "batch", batch.mean(), "sparse", bw)


if __name__ == "__main__":
    main()
