#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE's own code (build container only).

Runs, unmodified and from where they lie under /root/reference:
  * utils/prediction.py  (ctc_decode, ctc_decode2, ctc_decode_strict, ctc_predict, evaluate)
  * utils/basic_vad.py   (vad)
  * utils/queue.py       (SimpleQueue)
  * octbit/octbit_mat_mul_op.cc and positional_encoding/positional_encoding_op.cc
    through oracle/_ref (compiled behind oracle/tf_shim; see oracle/Makefile)
and writes small fixtures next to this script.  /root/reference does not exist on
the GPU box, so tests only ever read the committed .npz files.

    python tests/golden/make_golden.py

The probability sequences are handed to the reference decoders as float64 arrays
holding fp32-representable values: that reproduces the numpy-1.x semantics the
reference was written for (fp32 value compared with a Python float in double),
independently of the numpy version running this script.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("KWS_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def load_ref_module(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_prob_sequence(rng, T, C=6):
    """Peaky CTC-style posteriors that exercise thresholds, ties and lockouts."""
    kind = rng.integers(0, 4)
    logits = rng.normal(0.0, 1.0, (T, C))
    logits[:, C - 1] += 4.0                                  # blank dominates
    if kind == 0:                                            # scripted keyword-ish events
        t = int(rng.integers(0, 6))
        script = rng.choice([1, 2, 3, 3, 1, 2, 3, 4, 2, 1], size=int(rng.integers(1, 12)))
        if rng.random() < 0.5:
            script = np.array([1, 2, 3, 3] + list(script[:4]))
        for lab in script:
            dur = int(rng.integers(1, 6))
            peak = rng.choice([1.0, 2.5, 4.5, 6.0, 9.0])
            for d in range(dur):
                if t + d < T:
                    logits[t + d, lab] += peak + rng.normal(0, 0.7)
            t += dur + int(rng.integers(0, 7))
    elif kind == 1:                                          # dense random winners
        win = rng.integers(0, C, T)
        logits[np.arange(T), win] += rng.choice([3.0, 5.0, 8.0], T)
    elif kind == 2:                                          # flat / near-threshold
        logits = rng.normal(0.0, 0.4, (T, C))
    z = logits - logits.max(axis=1, keepdims=True) if T else logits
    p = np.exp(z)
    p = p / p.sum(axis=1, keepdims=True) if T else p
    p = p.astype(np.float32)
    if T and kind == 3:                                      # exact threshold / tie values
        vals = np.array([0.2, 0.4, 0.5, 0.6, 0.25, 0.19999999, 0.40000004, 0.5000001, 0.6000001],
                        dtype=np.float32)
        rows = rng.integers(0, T, max(1, T // 3))
        for r in rows:
            v = rng.choice(vals)
            cols = rng.choice(np.arange(1, 5), size=int(rng.integers(1, 3)), replace=False)
            p[r, 1:5] = np.float32(0.01)
            p[r, cols] = v                                   # ties -> first-max argmax rule
    return p


def main():
    pred = load_ref_module("utils/prediction.py", "ref_prediction")
    bvad = load_ref_module("utils/basic_vad.py", "ref_basic_vad")
    rqueue = load_ref_module("utils/queue.py", "ref_queue")
    rng = np.random.default_rng(20171017)

    # ---------------- decoders
    lengths = [0, 1, 2, 3, 4, 7, 28, 30, 30, 58, 60, 90, 150, 298, 298, 450, 450]
    seqs, outs = [], {k: [] for k in ("decode", "decode2", "strict")}
    preds = {k: [] for k in ("decode", "decode2", "strict")}
    for rep in range(14):
        for T in lengths:
            p = synth_prob_sequence(rng, T)
            seqs.append(p)
            p64 = p.astype(np.float64)
            r0 = pred.ctc_decode(p64) if T else np.asarray([0], np.int32)
            r1 = pred.ctc_decode2(p64, 6)
            r2 = pred.ctc_decode_strict(p64, 6)
            for k, r in (("decode", r0), ("decode2", r1), ("strict", r2)):
                outs[k].append(np.asarray(r, np.int32))
                preds[k].append(pred.ctc_predict(r, "1233"))
    # non-default parameters
    extra = []
    for rep in range(40):
        T = int(rng.integers(1, 120))
        p = synth_prob_sequence(rng, T)
        lockout = int(rng.integers(1, 6))
        thres = float(rng.choice([0.3, 0.45, 0.5, 0.7]))
        loose = float(rng.choice([0.1, 0.2, 0.3]))
        p64 = p.astype(np.float64)
        extra.append(dict(
            p=p, lockout=lockout, thres=thres, loose=loose,
            decode=np.asarray(pred.ctc_decode(p64, lockout, thres, loose), np.int32),
            decode2=np.asarray(pred.ctc_decode2(p64, 6, thres), np.int32),
            strict=np.asarray(pred.ctc_decode_strict(p64, 6, lockout, thres), np.int32)))

    def pack(list_of_arrays):
        off = np.cumsum([0] + [len(a) for a in list_of_arrays]).astype(np.int64)
        flat = np.concatenate(list_of_arrays) if list_of_arrays else np.zeros(0)
        return flat, off

    probs_flat, probs_off = pack([s.reshape(-1) for s in seqs])
    save = dict(probs=probs_flat.astype(np.float32), probs_off=probs_off)
    for k in outs:
        f, o = pack(outs[k])
        save[k] = f.astype(np.int32)
        save[k + "_off"] = o
        save[k + "_pred"] = np.asarray(preds[k], np.int32)
    ep, eo = pack([e["p"].reshape(-1) for e in extra])
    save.update(extra_probs=ep.astype(np.float32), extra_probs_off=eo,
                extra_lockout=np.asarray([e["lockout"] for e in extra], np.int32),
                extra_thres=np.asarray([e["thres"] for e in extra], np.float64),
                extra_loose=np.asarray([e["loose"] for e in extra], np.float64))
    for k in ("decode", "decode2", "strict"):
        f, o = pack([e[k] for e in extra])
        save["extra_" + k] = f.astype(np.int32)
        save["extra_" + k + "_off"] = o
    # ctc_predict / evaluate on hand-made sequences
    pred_cases = [[0, 1, 0, 2, 0, 3, 0, 3, 0], [0, 1, 0, 2, 0, 3, 0], [1, 2, 3, 3], [0], [],
                  [0, 1, 0, 2, 0, 3, -1, 3, 0], [4, 1, 2, 3, 3, 4], [1, 2, 3, 0, 0, 3], [1, 1, 2, 3, 3],
                  [1, 2, 3, 4, 3], [3, 3, 2, 1], [1, 2, 3, 3, 1, 2, 3, 3]]
    save["predict_cases"], save["predict_cases_off"] = pack([np.asarray(c, np.int32) for c in pred_cases])
    save["predict_cases"] = save["predict_cases"].astype(np.int32)
    save["predict_out"] = np.asarray([pred.ctc_predict(c, "1233") for c in pred_cases], np.int32)
    save["predict_out_123"] = np.asarray([pred.ctc_predict(c, "123") for c in pred_cases], np.int32)
    res = rng.integers(0, 2, 64)
    tgt = rng.integers(0, 2, 64)
    save["eval_result"], save["eval_target"] = res.astype(np.int32), tgt.astype(np.int32)
    save["eval_out"] = np.asarray(pred.evaluate(res.tolist(), tgt.tolist()), np.int64)
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **save)

    # ---------------- VAD + queue
    sigs, vout = [], []
    for i in range(48):
        n = int(rng.choice([160, 3600, 4800]))
        amp = float(rng.choice([1e-4, 1e-3, 5e-3, 6.2e-3, 6.3e-3, 1e-2, 0.1]))
        sig = (rng.standard_normal(n) * amp).astype(np.float32)
        sigs.append(sig)
        vout.append([bool(bvad.vad(sig, 30)), bool(bvad.vad(sig))])
    sflat, soff = pack(sigs)
    ops, snaps = [], []
    q = rqueue.SimpleQueue(15)
    for i in range(60):
        r = rng.random()
        if r < 0.08:
            q.clear()
            ops.append(-1)
        else:
            q.add(i)
            ops.append(i)
        snaps.append(np.asarray(list(q.get_all()) + [-1] * (15 - len(q.get_all())), np.int32))
    np.savez_compressed(os.path.join(HERE, "vad_queue_golden.npz"),
                        sig=sflat.astype(np.float32), sig_off=soff, vad=np.asarray(vout, np.bool_),
                        queue_ops=np.asarray(ops, np.int32), queue_snap=np.stack(snaps))

    # ---------------- native ops through the unmodified reference kernels
    from oracle import cref
    cref.build(ref=True)
    oct_cases = {}
    shapes = [(1, 1, 64), (2, 4, 64), (1, 256, 256), (30, 256, 256), (30, 128, 256), (30, 6, 128),
              (7, 3, 192), (5, 10, 512), (3, 2, 1024)]
    for ci, (A, B, K) in enumerate(shapes):
        for mode in ("signed", "unsigned", "saturating"):
            x = rng.standard_normal((A, K)).astype(np.float32)
            w = rng.integers(-127, 128, (B, K)).astype(np.int8)
            if mode == "unsigned":
                x = np.abs(x)
            if mode == "saturating":
                x = (np.abs(x) * 0.02 + 1.0).astype(np.float32)
                x[0, K // 2] = -1.0 if ci % 2 else x[0, K // 2]
                w = np.where(rng.random((B, K)) < 0.6, np.int8(127) if ci % 3 else np.int8(-128), w).astype(np.int8)
            bias = (127.0 * w.astype(np.float64).sum(axis=1)).astype(np.float32)
            scale = float(np.float32(rng.uniform(0.001, 0.05)))
            key = "c%d_%s" % (ci, mode)
            oct_cases[key + "_x"] = x
            oct_cases[key + "_w"] = w
            oct_cases[key + "_bias"] = bias
            oct_cases[key + "_scale"] = np.float32(scale)
            oct_cases[key + "_out"] = cref.ref_octbit_matmul(x, w, bias, scale)
    np.savez_compressed(os.path.join(HERE, "octbit_golden.npz"), **oct_cases)
    pe_cases = {}
    for (mp, sz) in [(10, 16), (400, 128), (7, 5), (3, 1), (1, 2), (33, 60), (1000, 64)]:
        pe_cases["pe_%d_%d" % (mp, sz)] = cref.ref_positional_encoding(mp, sz, fill=-9.0)
    np.savez_compressed(os.path.join(HERE, "posenc_golden.npz"), **pe_cases)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print("  %-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
