"""BASELINE configs[3]: the octbit-rewritten rnn_ctc graph (graph_octbit.pb: float cell_0, OctbitMatMul for the
gates / candidate of the upper layer and for the FC) on the GPU against the oracle loop that calls the UNMODIFIED
reference op (oracle/_ref) once per stream, step and MatMul -- the reference's batch-1 deployment semantics.

What can be asserted: the op itself is bit-exact (tests/test_gpu_ops.py); the graph around it contains sigmoid /
tanh / exp, whose last bit differs between CUDA and numpy, and the op's u8 quantiser turns a last-bit difference
into a full quantum (~1e-3 on a pre-activation) whenever a value sits on a rounding boundary.  So: the great
majority of streams must agree to float round-off, and every stream to within quantisation noise.
"""
import numpy as np
import pytest

from tests._util import make_config, synth_pcm16, to_product_weights

pytestmark = pytest.mark.gpu

TOL_ROUNDOFF = 2e-5      # sigmoid/tanh/exp last-bit differences carried through a chunk
TOL_QUANTUM = 3e-2       # one or a few flipped u8 codes (observed: a few 1e-3)


def _models(seed=1234, fc=True):
    from keyword_spotting_b200 import DeployModel, OctbitModelWeights
    from oracle import model as om
    ow = om.init_weights(seed=seed, n_mel=40)
    octw = om.octize_model(ow, fc=fc)
    dm = DeployModel(make_config(40), to_product_weights(ow))
    pw = OctbitModelWeights.from_float(dm.weights, fc=fc)
    # the device recipe (kws_octize_weight) equals the numpy restatement of octize_weight_int8_signed
    for l in octw["gates"]:
        np.testing.assert_array_equal(pw.gates[l].weight_q, octw["gates"][l][0])
        np.testing.assert_array_equal(pw.candidate[l].weight_q, octw["candidate"][l][0])
        assert np.float32(pw.gates[l].scale) == np.float32(octw["gates"][l][1])
        np.testing.assert_array_equal(pw.gates[l].bias, octw["gates"][l][2])
    dm.set_octbit(pw)
    return ow, octw, dm


def _agreement(got, want, axis_streams):
    err = np.abs(got - want)
    per_stream = err.max(axis=tuple(i for i in range(err.ndim) if i != axis_streams))
    return float(err.max()), float((per_stream < TOL_ROUNDOFF).mean())


@pytest.mark.parametrize("fc", [True, False], ids=["fc-octbit", "fc-float"])
def test_octbit_graph_single_step_from_given_state(fc):
    """One time step from a random state: every op call sees inputs equal to the oracle's up to the last bit of a
    sigmoid, so (nearly) every stream must agree to round-off."""
    from oracle import model as om
    ow, octw, dm = _models(fc=fc)
    rng = np.random.default_rng(3)
    S = 200
    mel = (np.abs(rng.standard_normal((S, 1, 40))) * rng.uniform(0.1, 5.0, (S, 1, 1))).astype(np.float32)
    st = (rng.uniform(-1, 1, (2, S, 128)) * 0.8).astype(np.float32)
    p_want, s_want, l_want = om.octbit_mel_forward(mel, st, ow, octw)
    p_got, s_got, l_got = dm.run_mel(mel, st, want_logits=True)
    worst_s, ok_s = _agreement(s_got, s_want, 1)
    worst_p, ok_p = _agreement(p_got, p_want, 0)
    assert ok_s >= 0.97 and ok_p >= 0.95, (ok_s, ok_p, worst_s, worst_p)
    assert worst_s < TOL_QUANTUM and worst_p < TOL_QUANTUM
    assert np.abs(l_got - l_want).max() < 10 * TOL_QUANTUM
    # and it is NOT the float graph: quantisation moves the state by ~1e-2
    _, s_float, _ = om.mel_forward(mel, st, ow)
    assert np.abs(s_got - s_float).max() > 1e-3
    dm.close()


def test_octbit_graph_chunk_forward_and_deploy_call():
    """30 carried steps on 70 streams (a full tile of 64 + a partial one), mel input and PCM input."""
    from oracle import model as om
    ow, octw, dm = _models()
    assert dm.octbit is not None
    rng = np.random.default_rng(4)
    S, n = 70, 30
    mel = (np.abs(rng.standard_normal((S, n, 40))) * rng.uniform(0.1, 5.0, (S, 1, 1))).astype(np.float32)
    st = (rng.uniform(-1, 1, (2, S, 128)) * 0.5).astype(np.float32)
    lens = rng.integers(0, n + 1, S).astype(np.int32)
    lens[:4] = [0, n, 1, n]
    for seq_len in (None, lens):
        p_want, s_want, _ = om.octbit_mel_forward(mel, st, ow, octw, seq_len=seq_len)
        p_got, s_got = dm.run_mel(mel, st, seq_len=seq_len)
        worst_s, ok_s = _agreement(s_got, s_want, 1)
        worst_p, ok_p = _agreement(p_got, p_want, 0)
        assert ok_s >= 0.75 and ok_p >= 0.7, (ok_s, ok_p, worst_s, worst_p)
        assert worst_s < TOL_QUANTUM and worst_p < TOL_QUANTUM, (worst_s, worst_p)
        np.testing.assert_allclose(p_got.sum(-1), 1.0, atol=1e-5)
    np.testing.assert_array_equal(s_got[:, 0], st[:, 0])              # length 0: state untouched
    pcm = synth_pcm16(rng, 20, 5120, silent_frac=0.0)
    z = np.zeros((2, 20, 128), np.float32)
    p_want, s_want, _ = om.octbit_deploy_forward(om.pcm16_to_float(pcm), z, ow, octw)
    p_got, s_got = dm(pcm, z)
    assert np.abs(s_got - s_want).max() < TOL_QUANTUM and np.abs(p_got - p_want).max() < TOL_QUANTUM
    # back to the float graph
    dm.set_octbit(None)
    p_f, s_f = dm(pcm, z)
    p_fw, s_fw, _ = om.deploy_forward(om.pcm16_to_float(pcm), z, ow)
    assert np.abs(p_f - p_fw).max() < 1e-3 and np.abs(s_f - s_fw).max() < 1e-3
    dm.close()


def test_octbit_graph_file_round_trip():
    """DeployModel.from_octbit_graph on a serialised graph_octbit.pb == set_octbit with the same constants."""
    from keyword_spotting_b200 import DeployModel
    from oracle import model as om
    from tests.test_graph_pb import _frozen_graph
    ow, octw, dm = _models(seed=12)
    data, _ = _frozen_graph(ow, "kernel", octbit=True, octbit_fc=True)
    dg = DeployModel.from_octbit_graph(data, n_mel=40)
    assert dg.octbit is not None and sorted(dg.octbit.gates) == [1] and dg.octbit.fc is not None
    rng = np.random.default_rng(9)
    pcm = synth_pcm16(rng, 33, 5120, silent_frac=0.0)
    z = np.zeros((2, 33, 128), np.float32)
    p1, s1 = dm(pcm, z)
    p2, s2 = dg(pcm, z)
    np.testing.assert_array_equal(p1, p2)
    np.testing.assert_array_equal(s1, s2)
    dm.close()
    dg.close()


def test_octbit_graph_streaming_server():
    """The detector loop on the octbit graph: VAD reset, tail carry, window decode, trigger -- triggers and labels
    bit-identical to the oracle loop once the frames whose decision lies within the quantisation tolerance of the
    threshold are taken from the GPU."""
    from keyword_spotting_b200 import StreamingDetector
    from oracle import model as om, streaming as ost
    from tests.test_gpu_parity_scale import _MarginJudge
    ow, octw, dm = _models()
    ow3 = om.init_weights(seed=1234, n_mel=40)
    S, chunks, chunk = 48, 6, 4800
    rng = np.random.default_rng(8)
    pcm = synth_pcm16(rng, S, chunk * chunks, silent_frac=0.2)
    det = StreamingDetector(dm, S, keyword="12", decode_thres=0.2)
    orc = ost.StreamOracle(ow, S, label="12", decode_thres=0.2,
                           forward=lambda full, state: om.octbit_deploy_forward(full, state, ow, octw))
    judge = _MarginJudge(0.2, tol=TOL_QUANTUM)
    n_lab = 0
    for c in range(chunks):
        blk = pcm[:, c * chunk:(c + 1) * chunk]
        trig, probs, nfr = det.step(blk, want_probs=True)
        labels, counts = det.window_labels()
        judge.gpu = probs
        want = orc.step(blk, decide_on=judge)
        n = want["softmax"].shape[1]
        assert np.abs(probs[:, :n] - want["softmax"]).max() < TOL_QUANTUM, c
        assert np.abs(det.state().cpu().numpy() - want["state"]).max() < TOL_QUANTUM, c
        np.testing.assert_array_equal(trig, want["trigger"])
        for s in range(S):
            if not want["trigger"][s]:
                np.testing.assert_array_equal(labels[s, :counts[s]], want["labels"][s])
                n_lab += len(want["labels"][s]) // 2
    assert n_lab > 10
    det.close()
    dm.close()
