"""GPU parity, through the C-ABI: octbit_mat_mul (bit-exact), positional_encoding, CTC decoders."""
import numpy as np
import pytest

from tests._util import golden, unpack

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- positional encoding (K6)
def test_posenc_bit_exact_vs_reference_kernel_golden():
    from keyword_spotting_b200.positional_encoding.positional_encoding_op import positional_encoding
    g = golden("posenc_golden.npz")
    total = mism = 0
    worst = 0.0
    for key in g.files:
        _, mp, sz = key.split("_")
        want = g[key]
        got = positional_encoding(int(mp), int(sz), as_numpy=True, fill=-9.0)
        assert got.shape == want.shape and got.dtype == np.float32
        total += want.size
        mism += int((got.view(np.uint32) != want.view(np.uint32)).sum())
        worst = max(worst, float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) if want.size else 0.0)
    # double sin/cos may differ from glibc in the last bit of the DOUBLE; after rounding to float
    # that shows up at most once per ~1e8 elements.  Bar: within 1 fp32 ulp everywhere, and
    # bit-identical for all but at most 2 of the ~120k golden values.
    assert worst <= 6e-8, worst
    assert mism <= 2, (mism, total)


def test_posenc_odd_size_and_oracle_agree():
    from keyword_spotting_b200.positional_encoding.positional_encoding_op import positional_encoding
    from oracle import posenc as ope
    for mp, sz in [(0, 8), (5, 1), (9, 7), (400, 128), (798 // 2 + 1, 128), (2048, 96)]:
        got = positional_encoding(mp, sz, as_numpy=True, fill=3.5)
        want = ope.positional_encoding(mp, sz, fill=3.5)
        assert got.shape == want.shape
        if want.size:
            assert np.abs(got - want).max() <= 6e-8
        if sz % 2 and mp:
            assert (got[:, -1] == 3.5).all()           # never written, positional_encoding_op.cc:45
    with pytest.raises(ValueError):
        positional_encoding(4, 0)


# ---------------------------------------------------------------- octbit (K5)
def _oct(x, w, bias, scale, **kw):
    from keyword_spotting_b200.octbit.octbit_ops import octbit_mat_mul
    return octbit_mat_mul(x, w, scale=scale, bias=bias, **kw)


def test_octbit_reference_known_answers_on_gpu():
    # octbit/octbit_ops_test.py:24-34 and :41-53, assertAllEqual
    out = _oct([[-1.0] * 64], np.arange(64, dtype=np.int8)[None], [127 * 2016.0], 3.0)
    np.testing.assert_array_equal(out, [[-6048.0]])
    w = np.stack([np.ones(64)] + [np.arange(64)] * 3).astype(np.int8)
    out = _oct(-np.ones((2, 64), np.float32), w, [127 * 64.0, 127 * 2016.0, 127 * 2016.0, 127 * 2016.0], 2.0)
    np.testing.assert_array_equal(out, [[-128.0, -4032.0, -4032.0, -4032.0]] * 2)
    out = _oct(-np.ones((2, 64), np.float32), w, [127 * 64.0, 127 * 2016.0, 127 * 2016.0, 127 * 2016.0], 2.0, _exact=True)
    np.testing.assert_array_equal(out, [[-128.0, -4032.0, -4032.0, -4032.0]] * 2)


def test_octbit_bit_exact_vs_reference_kernel_golden():
    g = golden("octbit_golden.npz")
    keys = sorted(k[:-2] for k in g.files if k.endswith("_x"))
    for k in keys:
        x, w, bias, scale, want = g[k + "_x"], g[k + "_w"], g[k + "_bias"], float(g[k + "_scale"]), g[k + "_out"]
        got = _oct(x, w, bias, scale)
        assert got.tobytes() == want.tobytes(), (k, np.abs(got - want).max())
        got = _oct(x, w, bias, scale, _exact=True)
        assert got.tobytes() == want.tobytes(), ("exact", k, np.abs(got - want).max())


def test_octbit_bit_exact_vs_oracle_model_shapes_and_adversarial():
    from oracle import octbit as ooct
    rng = np.random.default_rng(42)
    # the matmuls the graph rewriter converts: K=256->256, K=256->128, K=128->6
    for (A, B, K) in [(1, 256, 256), (30, 128, 256), (300, 6, 128), (1000, 256, 256), (77, 100, 512), (5, 9, 64)]:
        for mode in ("signed", "unsigned", "sat_pos", "sat_neg", "w128"):
            x = rng.standard_normal((A, K)).astype(np.float32)
            w = rng.integers(-127, 128, (B, K)).astype(np.int8)
            if mode == "unsigned":
                x = np.abs(x)
            if mode.startswith("sat"):
                x = (1.0 + 0.01 * rng.random((A, K))).astype(np.float32)       # q ~ 252..254 everywhere
                w[:] = 127 if mode == "sat_pos" else -127                      # every pair saturates
                w[rng.random((B, K)) < 0.2] = 3
            if mode == "w128":
                w[rng.random((B, K)) < 0.3] = -128
                x = np.abs(x) + 0.5
            bias = (127.0 * w.astype(np.float64).sum(axis=1)).astype(np.float32)
            scale = 0.0173
            want = ooct.octbit_mat_mul(x, w, scale=scale, bias=bias)
            got = _oct(x, w, bias, scale)
            assert got.tobytes() == want.tobytes(), (A, B, K, mode, np.abs(got - want).max())
            got = _oct(x, w, bias, scale, _exact=True)
            assert got.tobytes() == want.tobytes(), ("exact", A, B, K, mode)


def test_octbit_large_k_uses_exact_lanes():
    from oracle import octbit as ooct
    rng = np.random.default_rng(1)
    for K in (576, 1024, 4096):
        x = (rng.standard_normal((3, K)) * 5).astype(np.float32)
        w = rng.integers(-128, 128, (5, K)).astype(np.int8)
        bias = rng.standard_normal(5).astype(np.float32)
        want = ooct.octbit_mat_mul(x, w, scale=0.5, bias=bias)
        assert _oct(x, w, bias, 0.5).tobytes() == want.tobytes(), K


def test_octbit_cuda_tensors_and_quantised_model_weights():
    """config 4: W_ih/W_hh/FC of layer 1 quantised with the rewriter's recipe, then the op."""
    import torch
    from keyword_spotting_b200.octbit.octbit_graph import octize_weight_int8_signed
    from oracle import model as om, octbit as ooct
    w = om.init_weights(seed=1234)
    rng = np.random.default_rng(7)
    for W, A in ((w.gates_kernel[1], 30), (w.cand_kernel[1], 30), (w.fc_w, 64)):
        wq_t, scale, bias = octize_weight_int8_signed(W)
        owq, oscale, obias = ooct.octize_weight_int8_signed(W)
        np.testing.assert_array_equal(wq_t, owq)
        assert scale == oscale
        np.testing.assert_array_equal(bias, obias.astype(np.float32))
        x = rng.uniform(-1, 1, (A, W.shape[0])).astype(np.float32)
        want = ooct.octbit_mat_mul(x, owq, scale=np.float32(oscale), bias=obias)
        xd = torch.from_numpy(x).cuda()
        got = _oct(xd, torch.from_numpy(wq_t).cuda(), torch.from_numpy(bias).cuda(), scale)
        assert got.is_cuda
        assert got.cpu().numpy().tobytes() == want.tobytes()
        # and, where no adjacent pair saturates int16 (the reference's maddubs clips those),
        # the quantised product tracks the fp32 matmul -- sanity of the whole recipe
        q, _, _ = ooct.quantize_activations(x)
        pair = (q.astype(np.int32)[:, None, :] * owq.astype(np.int32)[None]).reshape(A, owq.shape[0], -1, 2).sum(-1)
        clean = (np.abs(pair) <= 32767).all(axis=2)
        ref = x @ W
        assert clean.any()
        assert np.abs(got.cpu().numpy() - ref)[clean].max() < 0.03 * np.abs(ref).max() + 0.03


def test_octbit_errors_follow_the_op():
    from keyword_spotting_b200 import InvalidArgumentError
    x = np.zeros((2, 64), np.float32)
    w = np.zeros((3, 64), np.int8)
    b = np.zeros(3, np.float32)
    from keyword_spotting_b200.octbit.octbit_ops import octbit_mat_mul
    with pytest.raises(InvalidArgumentError, match="scale has to be positive"):
        octbit_mat_mul(x, w, bias=b)                                   # default scale=0.0 (:46)
    with pytest.raises(InvalidArgumentError, match="b need to be transposed"):
        octbit_mat_mul(x, w, transpose_b=False, scale=1.0, bias=b)     # :41
    with pytest.raises(InvalidArgumentError, match="a cannot to be transposed"):
        octbit_mat_mul(x, w, transpose_a=True, scale=1.0, bias=b)      # :42
    with pytest.raises(InvalidArgumentError, match="16 aligned"):
        octbit_mat_mul(x[:, :32], w[:, :32], scale=1.0, bias=b)        # :65-67
    with pytest.raises(InvalidArgumentError, match="f is not equal"):
        octbit_mat_mul(x, w[:, :32], scale=1.0, bias=b)                # :61-63
    with pytest.raises(InvalidArgumentError, match="not a matrix"):
        octbit_mat_mul(x[0], w, scale=1.0, bias=b)                     # :70-71
    assert octbit_mat_mul(np.zeros((0, 64), np.float32), w, scale=1.0, bias=b).shape == (0, 3)
    np.testing.assert_array_equal(octbit_mat_mul(x, w, scale=1.0, bias=b), np.zeros((2, 3), np.float32))  # all-zero x


def test_octbit_full_size_properties():
    """config 4 at A = 131072*30 is too big for the numpy oracle; check size-independent properties on a
    1M-row slab: (i) row independence given the same tensor-wide range, (ii) tensor-core == exact path."""
    import torch
    from oracle import octbit as ooct
    rng = np.random.default_rng(3)
    A, B, K = 1 << 20, 256, 256
    base = rng.standard_normal((4096, K)).astype(np.float32)
    base[0, 0] = 6.0
    base[1, 1] = -6.5                       # pins min/max inside the first block
    w = rng.integers(-127, 128, (B, K)).astype(np.int8)
    bias = (127.0 * w.astype(np.float64).sum(axis=1)).astype(np.float32)
    xd = torch.from_numpy(base).cuda().repeat(A // 4096, 1)
    wd, bd = torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    got = _oct(xd, wd, bd, 0.01)
    want_block = ooct.octbit_mat_mul(base, w, scale=0.01, bias=bias)
    wb = torch.from_numpy(want_block).cuda()
    assert torch.equal(got.view(A // 4096, 4096, B), wb.unsqueeze(0).expand(A // 4096, 4096, B))
    assert torch.equal(got, _oct(xd, wd, bd, 0.01, _exact=True))


# ---------------------------------------------------------------- decoders (K4)
@pytest.mark.parametrize("mode,key", [(0, "decode"), (1, "decode2"), (2, "strict")])
def test_decode_bit_exact_vs_reference_golden(mode, key):
    from keyword_spotting_b200.utils import prediction as P
    g = golden("decode_golden.npz")
    po = g["probs_off"]
    n = len(po) - 1
    Tmax = max((po[i + 1] - po[i]) // 6 for i in range(n))
    batch = np.zeros((n, Tmax, 6), np.float32)
    lens = np.zeros(n, np.int32)
    for i in range(n):
        p = unpack(g["probs"], po, i, 6)
        batch[i, :len(p)] = p
        batch[i, len(p):] = 0.9          # garbage past the length must be ignored
        lens[i] = len(p)
    labels, counts, trig = P.decode_batch(batch, lens=lens, mode=mode)
    for i in range(n):
        want = unpack(g[key], g[key + "_off"], i)
        assert counts[i] == len(want), (i, counts[i], len(want))
        np.testing.assert_array_equal(labels[i, :counts[i]], want)
        assert (labels[i, counts[i]:] == -1).all()
        assert trig[i] == g[key + "_pred"][i]
    # the single-sequence reference signatures
    for i in (0, 5, 17, 100, 200):
        p = unpack(g["probs"], po, i, 6)
        fn = {0: lambda q: P.ctc_decode(q), 1: lambda q: P.ctc_decode2(q, 6), 2: lambda q: P.ctc_decode_strict(q, 6)}[mode]
        out = fn(p)
        assert out.dtype == np.int32
        np.testing.assert_array_equal(out, unpack(g[key], g[key + "_off"], i))


def test_decode_nondefault_params_and_truncation():
    from keyword_spotting_b200.utils import prediction as P
    g = golden("decode_golden.npz")
    eo = g["extra_probs_off"]
    for i in range(len(eo) - 1):
        p = unpack(g["extra_probs"], eo, i, 6)
        lo, th, ls = int(g["extra_lockout"][i]), float(g["extra_thres"][i]), float(g["extra_loose"][i])
        np.testing.assert_array_equal(P.ctc_decode(p, lo, th, ls), unpack(g["extra_decode"], g["extra_decode_off"], i))
        np.testing.assert_array_equal(P.ctc_decode2(p, 6, th), unpack(g["extra_decode2"], g["extra_decode2_off"], i))
        np.testing.assert_array_equal(P.ctc_decode_strict(p, 6, lo, th), unpack(g["extra_strict"], g["extra_strict_off"], i))
    # max_labels smaller than the sequence: labels truncated, count and trigger still exact
    p = np.zeros((1, 40, 6), np.float32)
    for t, lab in enumerate([1, 2, 3, 3] * 10):
        p[0, t, lab] = 0.9 if t % 2 == 0 else 0.0
    labels, counts, trig = P.decode_batch(p, mode=P.MODE_CTC_DECODE2, max_labels=5)
    assert labels.shape == (1, 5) and counts[0] == 41 and (labels[0] == [0, 1, 0, 3, 0]).all()
    assert trig[0] == 0
    labels, counts, trig = P.decode_batch(p, mode=P.MODE_CTC_DECODE2, keyword="13")
    assert trig[0] == 1


def test_decode_full_size_properties():
    """4096 x 298 (config 2 shape): device decode == oracle on a sample; trigger == substring test of labels."""
    import torch
    from keyword_spotting_b200.utils import prediction as P
    from oracle import prediction as op
    rng = np.random.default_rng(9)
    S, T = 4096, 298
    logits = rng.normal(0, 1, (S, T, 6)).astype(np.float32)
    logits[:, :, 5] += 2.5
    win = rng.integers(1, 5, (S, T))
    boost = rng.random((S, T)) < 0.15
    np.add.at(logits, (np.nonzero(boost)[0], np.nonzero(boost)[1], win[boost]), 6.0)
    probs = torch.softmax(torch.from_numpy(logits).cuda(), dim=-1)
    for mode in (0, 1, 2):
        labels, counts, trig = P.decode_batch(probs, mode=mode)
        labels, counts, trig = labels.cpu().numpy(), counts.cpu().numpy(), trig.cpu().numpy()
        pn = probs.cpu().numpy()
        for s in rng.integers(0, S, 48):
            want = op.decode(pn[s], mode)
            np.testing.assert_array_equal(labels[s, :counts[s]], want)
        for s in range(0, S, 7):
            assert trig[s] == op.ctc_predict(labels[s], "1233")
        assert trig.sum() > 0


def test_wer_matches_reference_golden():
    """kws_edit_distance + the utils/wer.py mirrors vs the reference's own utils/wer.py outputs."""
    from keyword_spotting_b200.utils import wer as kw
    g = golden("wer_golden.npz")
    refs = [g["ref"][g["ref_off"][i]:g["ref_off"][i + 1]].tolist() for i in range(len(g["ref_off"]) - 1)]
    hyps = [g["hyp"][g["hyp_off"][i]:g["hyp_off"][i + 1]].tolist() for i in range(len(g["hyp_off"]) - 1)]
    d = kw.edit_distance_batch(refs, hyps)
    lens = np.asarray([len(r) for r in refs], np.float64)
    got = np.where(lens > 0, d / np.maximum(lens, 1), d)
    np.testing.assert_array_equal(got, g["wer"])
    assert kw.wer(refs[5], hyps[5]) == g["wer"][5]
    calc = kw.WERCalculator([0, 5])
    np.testing.assert_array_equal(calc.cal_batch_wer(g["calc_r"], g["calc_h"]), g["calc_wer"])
    bw = kw.batch_wer(int(g["sp_bs"]), g["sp_r_index"], g["sp_r_value"], g["sp_h_index"], g["sp_h_value"])
    assert abs(bw - float(g["sp_wer"])) < 1e-12
    with pytest.raises(ValueError):
        kw.edit_distance_batch([[1] * 255], [[1]])


def test_validation_harness_matches_reference_loop():
    """evaluation.evaluate_softmax / validate vs the reference's validation arithmetic restated with the oracle
    decoders (main.py:290-304): ctc_decode -> ctc_predict -> evaluate, summed over batches."""
    import torch
    from keyword_spotting_b200 import DeployModel, evaluation
    from oracle import model as om, prediction as op
    from tests._util import make_config, to_product_weights
    g = golden("decode_golden.npz")
    po = g["probs_off"]
    T = 120
    seqs = []
    for i in range(len(po) - 1):
        p = unpack(g["probs"], po, i, 6)
        if len(p) >= T:
            seqs.append(p[:T])
    probs = np.stack(seqs[:64]).astype(np.float32)
    rng = np.random.default_rng(3)
    correct = rng.integers(0, 2, len(probs)).astype(np.int32)
    want_res = [op.ctc_predict(op.ctc_decode(p), "1233") for p in probs]
    want = op.evaluate(want_res, correct.tolist())
    miss, target, fa, res = evaluation.evaluate_softmax(probs, correct)
    assert (miss, target, fa) == tuple(want)
    np.testing.assert_array_equal(res, want_res)
    # whole loop on the model: mel batches with ragged lengths
    ow = om.init_weights(seed=7, n_mel=40)
    ow.fc_w = (ow.fc_w * 6).astype(np.float32)
    dm = DeployModel(make_config(40), to_product_weights(ow), precision="fp32")
    batches, tot = [], [0, 0, 0, 0]
    for b in range(3):
        mel = (np.abs(rng.standard_normal((16, 60, 40))) * 2).astype(np.float32)
        lens = rng.integers(20, 61, 16).astype(np.int32)
        corr = rng.integers(0, 2, 16).astype(np.int32)
        batches.append((mel, lens, corr))
        p, _, _ = om.mel_forward(mel, np.zeros((2, 16, 128), np.float32), ow, seq_len=lens)
        r = [op.ctc_predict(op.ctc_decode(p[i, :lens[i]]), "1233") for i in range(16)]
        m, t, f = op.evaluate(r, corr.tolist())
        tot = [tot[0] + m, tot[1] + t, tot[2] + f, tot[3] + 16]
    out = evaluation.validate(dm, batches)
    assert [out["miss"], out["target"], out["false_accept"], out["total"]] == tot
    assert "miss rate: %d/%d" % (tot[0], tot[1]) in out.report()
    dm.close()
