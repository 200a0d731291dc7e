"""The tensor-core formulation of the front end (csrc/frontend_tc.cu: hop-block partial DFTs on tcgen05, twiddles in
tensor memory) against the float64 oracle and against the FFT kernel, through the same C-ABI entry points.
It is selectable per model (``DeployModel(frontend="tc")``); the FFT kernel is the default."""
import numpy as np
import pytest

from tests._util import make_config, synth_pcm16, to_product_weights

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return float(np.abs(got - want).max() / max(1e-12, np.abs(want).max()))


@pytest.fixture(scope="module", params=[40, 60], ids=lambda m: "mel%d" % m)
def pair(request):
    from keyword_spotting_b200 import DeployModel
    from oracle import model as om
    ow = om.init_weights(seed=1234, n_mel=request.param)
    tc = DeployModel(make_config(request.param), to_product_weights(ow), precision="fp32", frontend="tc")
    fft = DeployModel(make_config(request.param), to_product_weights(ow), precision="fp32", frontend="fft")
    assert tc.frontend_kind == "tc" and fft.frontend_kind == "fft"
    yield ow, tc, fft
    tc.close()
    fft.close()


def test_tc_frontend_matches_float64_oracle_and_fft_kernel(pair):
    from oracle import model as om
    ow, tc, fft = pair
    rng = np.random.default_rng(5678)
    # one frame, one chunk, several items per stream (30-frame groups), ragged ends, unaligned row strides, > 148 items
    for S, L in [(1, 400), (1, 4800), (3, 5120), (7, 559), (5, 560), (4, 1361), (2, 48000), (3, 12345), (400, 5120)]:
        pcm16 = synth_pcm16(rng, S, L, silent_frac=0.2)
        want64 = om.pcm_to_mel(om.pcm16_to_float(pcm16).astype(np.float64), ow, np.float64)
        got = tc.frontend(pcm16)
        assert got.shape == want64.shape
        assert _rel(got, want64) < 2e-6, (S, L, _rel(got, want64))
        assert _rel(got, fft.frontend(pcm16)) < 2e-6
    assert tc.frontend(np.zeros((2, 399), np.int16)).shape == (2, 0, ow.n_mel)
    # float PCM has no exact fp16 split: it is served by the FFT kernel whatever the model's setting
    pcmf = om.pcm16_to_float(synth_pcm16(rng, 3, 5120, silent_frac=0.0))
    np.testing.assert_array_equal(tc.frontend(pcmf), fft.frontend(pcmf))


def test_tc_frontend_accuracy_is_relative_to_each_signals_own_level(pair):
    """int16 x = fp16(x) + residual is exact at every level: quiet streams are as accurate as loud ones."""
    from oracle import model as om
    ow, tc, _ = pair
    rng = np.random.default_rng(1)
    L = 5120
    t = np.arange(L) / 16000.0
    signals = [rng.standard_normal(L) * 3, rng.standard_normal(L) * 30, rng.standard_normal(L) * 3000,
               rng.standard_normal(L) * 30000, 30000 * np.sin(2 * np.pi * 997.3 * t), 32767 * np.sign(np.sin(2 * np.pi * 440 * t)),
               15000 + rng.standard_normal(L) * 100, np.full(L, -32768.0)]
    for x in signals:
        pcm16 = np.clip(np.rint(x), -32768, 32767).astype(np.int16)[None, :]
        want = om.pcm_to_mel(om.pcm16_to_float(pcm16).astype(np.float64), ow, np.float64)
        got = tc.frontend(pcm16)
        scale = max(np.abs(want).max(), 1e-3)              # a pure DC signal has no energy above 300 Hz: absolute floor
        assert np.abs(got - want).max() / scale < 5e-5, np.abs(got - want).max() / scale


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_streaming_server_on_the_tc_frontend(precision):
    """The fused pre-step of the tensor-core kernel (VAD, tail carry, frame count) against the detector-loop oracle,
    with chunk sizes that leave unaligned tails, and the server's steady state."""
    from keyword_spotting_b200 import DeployModel, StreamingDetector
    from oracle import model as om, streaming as ost
    ow = om.init_weights(seed=99, n_mel=40)
    ow.fc_w = (ow.fc_w * 3).astype(np.float32)
    dm = DeployModel(make_config(40), to_product_weights(ow), precision=precision, frontend="tc")
    tol = 1e-4 if precision == "fp32" else 1e-3
    S = 150
    rng = np.random.default_rng(4)
    det = StreamingDetector(dm, S, max_chunk=4801, keyword="4321", decode_thres=2.0)     # never fires: state is never reset by a trigger
    orc = ost.StreamOracle(ow, S, label="4321", decode_thres=2.0)
    sizes = [4800, 4800, 3600, 4801, 1234, 400, 4800, 4799, 2000, 4800, 4800, 4800]
    for i, n in enumerate(sizes):
        blk = synth_pcm16(rng, S, n, silent_frac=0.15)
        want = orc.step(blk)
        trig, probs, nfr = det.step(blk, want_probs=True)
        nf = want["softmax"].shape[1]
        assert (nfr == nf).all(), (i, nfr[:4], nf)
        assert np.abs(probs[:, :nf] - want["softmax"]).max() < tol, i
        assert np.abs(det.state().cpu().numpy() - want["state"]).max() < tol, i
        np.testing.assert_array_equal(trig, want["trigger"])
    det.close()
    dm.close()
