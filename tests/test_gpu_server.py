"""The wave server (kws_server_*, the HotwordDetector.start loop at scale, detector.py:148-212) against ONE big
stream object fed the same chunks: graph replay, direct enqueue, the pipelined serve loop and the copy-only mode."""
import numpy as np
import pytest

from tests._util import make_config, synth_pcm16, to_product_weights

pytestmark = pytest.mark.gpu

CHUNK = 4800


def _model(precision="tc", gain=3.0, seed=21):
    from keyword_spotting_b200 import DeployModel
    from oracle import model as om
    ow = om.init_weights(seed=seed, n_mel=40)
    ow.fc_w = (ow.fc_w * gain).astype(np.float32)
    return DeployModel(make_config(40), to_product_weights(ow), precision=precision)


@pytest.mark.parametrize("use_graphs", [True, False], ids=["graphs", "direct"])
@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_wave_server_matches_one_big_stream_object(use_graphs, precision):
    import torch
    from keyword_spotting_b200 import StreamingDetector, WaveServer
    dm = _model(precision)
    W, Sw, chunks = 4, 160, 9
    S = W * Sw
    rng = np.random.default_rng(77)
    pcm = synth_pcm16(rng, S, CHUNK * chunks, silent_frac=0.0)
    quiet = rng.random((S, chunks)) < 0.3
    for s, c in zip(*np.nonzero(quiet)):
        pcm[s, c * CHUNK:(c + 1) * CHUNK] = rng.integers(-2, 3, CHUNK)
    big = StreamingDetector(dm, S, keyword="1")
    srv = WaveServer(dm, S, waves=W, use_graphs=use_graphs, keyword="1")
    assert srv.graphs == use_graphs and srv.streams_per_wave == Sw
    n_trig = 0
    for c in range(chunks):
        blk = pcm[:, c * CHUNK:(c + 1) * CHUNK]
        want = big.step(blk)
        got = srv.step(blk)
        np.testing.assert_array_equal(got, want, err_msg="chunk %d" % c)
        n_trig += int(want.sum())
        assert torch.equal(srv.state(), big.state()), c
        lw, cw = big.window_labels(max_labels=64)
        lg, cg = srv.window_labels(max_labels=64)
        np.testing.assert_array_equal(cg, cw)
        np.testing.assert_array_equal(lg, lw)
    assert n_trig >= 5
    st = srv.stats()
    assert st["samples"] == W * chunks and st["triggers"] == n_trig and 0 < st["p50"] <= st["p99"] <= st["max"]
    srv.close()
    big.close()
    dm.close()


def test_wave_server_serve_loop_equals_step_by_step_and_copy_only_leaves_state_alone():
    """kws_server_serve (depth waves in flight, two ingest slots alternating) == the same chunks fed one at a time;
    the copy-only mode moves bytes but never touches stream state."""
    import torch
    from keyword_spotting_b200 import StreamingDetector, WaveServer
    dm = _model("tc")
    W, Sw, rounds = 8, 128, 7
    S = W * Sw
    rng = np.random.default_rng(5)
    two = synth_pcm16(rng, S, CHUNK * 2, silent_frac=0.1)          # the two chunks that alternate in the slots
    srv = WaveServer(dm, S, waves=W, keyword="1")
    for w in range(W):
        for b in range(2):                                          # fill both ingest slots of every wave
            srv.ingest_slot(w)[...] = two[w * Sw:(w + 1) * Sw, b * CHUNK:(b + 1) * CHUNK]
            srv.submit(w)
        for b in range(2):
            srv.wait(w)
    srv.reset()
    srv.stats(reset=True)
    srv.set_copy_only(True)
    srv.serve(3, depth=3)
    assert srv.stats(reset=True)["samples"] == 3 * W
    assert float(srv.state().abs().max()) == 0.0                    # copy-only: no kernel ran
    srv.set_copy_only(False)
    srv.serve(rounds, depth=3)
    st = srv.stats()
    assert st["samples"] == rounds * W
    big = StreamingDetector(dm, S, keyword="1")
    n_trig = 0
    # copy-only submissions advanced the slot parity by 3: the first served chunk is slot (2 + 3) % 2 = 1
    for k in range(rounds):
        b = (k + 1) % 2
        n_trig += int(big.step(two[:, b * CHUNK:(b + 1) * CHUNK]).sum())
    assert torch.equal(srv.state(), big.state())
    assert st["triggers"] == n_trig
    lw, cw = big.window_labels(max_labels=64)
    lg, cg = srv.window_labels(max_labels=64)
    np.testing.assert_array_equal(cg, cw)
    np.testing.assert_array_equal(lg, lw)
    srv.close()
    big.close()
    dm.close()


def test_wave_server_argument_checks():
    from keyword_spotting_b200 import InvalidArgumentError, WaveServer
    dm = _model("tc")
    with pytest.raises(InvalidArgumentError):
        WaveServer(dm, 100, waves=3)                                # not a multiple
    srv = WaveServer(dm, 256, waves=2)
    with pytest.raises(InvalidArgumentError):
        srv.wait(0)                                                 # nothing in flight
    srv.submit(0)
    srv.submit(0)
    with pytest.raises(InvalidArgumentError):
        srv.submit(0)                                               # two chunks already in flight
    with pytest.raises(InvalidArgumentError):
        srv.set_copy_only(True)                                     # chunks in flight
    srv.wait(0)
    srv.wait(0)
    with pytest.raises(InvalidArgumentError):
        srv.ingest_slot(5)
    srv.close()
    dm.close()
