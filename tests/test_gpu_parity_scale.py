"""Parity of the BENCHMARKED path (tensor-core precision, `kws_stream_step`) at the benchmarked size, and the
rigorous form of "bit-exact labels and triggers under a float tolerance".

north_star: labels / triggers bit-exact, per-frame softmax and carried state within max-abs 1e-3.  With a float
tolerance on the probabilities, bit-exact decisions can only be demanded where the decision is not within that
tolerance of a threshold.  The statement asserted here, for EVERY stream and chunk:

    take the frames whose ctc_decode2 token could differ between two softmax arrays that agree within 1e-3
    (oracle.streaming.ambiguous_frames: winner within 1e-3 of the threshold, or runner-up within 2e-3 of an
    above-threshold winner) from the GPU; then the GPU's triggers and window labels equal the fp32 detector-loop
    oracle's bit for bit, and its probabilities / state stay within 1e-3 -- no stream is masked out.

The fraction of ambiguous frames is printed and bounded.
"""
import numpy as np
import pytest

from tests._util import make_config, synth_pcm16, to_product_weights

pytestmark = pytest.mark.gpu

TOL_CONTRACT = 1e-3
CHUNK = 4800


def _boost_fc(ow, gain=3.0):
    """Random-init posteriors never cross the decode thresholds; a 3x FC makes labels, triggers and resets frequent
    while the tensor-core path's deviation stays inside the 1e-3 contract (the FC amplifies state noise linearly:
    emulated 2.1e-4 at 1x, 6.4e-4 at 3x, 1.2e-3 at 6x -- oracle.model with operand_dtype=float16)."""
    import copy
    w2 = copy.deepcopy(ow)
    w2.fc_w = (w2.fc_w * gain).astype(np.float32)
    return w2


def _chunked_pcm_with_silences(rng, n, chunks, quiet_frac=0.3):
    pcm = synth_pcm16(rng, n, CHUNK * chunks, silent_frac=0.0)
    quiet = rng.random((n, chunks)) < quiet_frac
    for s, c in zip(*np.nonzero(quiet)):
        pcm[s, c * CHUNK:(c + 1) * CHUNK] = rng.integers(-2, 3, CHUNK)
    return pcm


class _MarginJudge:
    """Feeds the oracle's decision window with the oracle's own softmax except at ambiguous frames, which are
    taken from the GPU's softmax; counts them."""

    def __init__(self, thres, tol=TOL_CONTRACT):
        self.thres, self.tol = thres, tol
        self.frames = 0
        self.ambiguous = 0
        self.gpu = None

    def __call__(self, softmax):
        from oracle import streaming as ost
        amb = ost.ambiguous_frames(softmax, self.thres, self.tol)
        self.frames += amb.size
        self.ambiguous += int(amb.sum())
        out = softmax.copy()
        out[amb] = self.gpu[:, :softmax.shape[1]][amb]
        return out


def _check_chunk(want, trig, probs, labels, counts, state, where):
    n = want["softmax"].shape[1]
    perr = float(np.abs(probs[:, :n] - want["softmax"]).max())
    serr = float(np.abs(state - want["state"]).max())
    assert perr < TOL_CONTRACT and serr < TOL_CONTRACT, (where, perr, serr)
    np.testing.assert_array_equal(trig, want["trigger"], err_msg=str(where))
    for s in range(len(trig)):
        if want["trigger"][s]:
            assert counts[s] == 1, (where, s)              # the window was cleared
        else:
            assert counts[s] == len(want["labels"][s]), (where, s)
            k = min(int(counts[s]), labels.shape[1])
            np.testing.assert_array_equal(labels[s, :k], want["labels"][s][:k], err_msg=str((where, s)))
    return perr, serr


def test_tc_server_labels_and_triggers_bit_exact_outside_the_tolerance_margin(capsys):
    """tensor-core server, 256 streams x 24 chunks, boosted FC so that labels / triggers / resets all occur."""
    from keyword_spotting_b200 import DeployModel, StreamingDetector
    from oracle import model as om, streaming as ost
    ow = _boost_fc(om.init_weights(seed=1234, n_mel=40))
    dm = DeployModel(make_config(40), to_product_weights(ow))
    assert dm.precision == "tc"
    S, chunks = 256, 24
    rng = np.random.default_rng(5678)
    pcm = _chunked_pcm_with_silences(rng, S, chunks)
    det = StreamingDetector(dm, S, keyword="1")
    orc = ost.StreamOracle(ow, S, label="1")
    judge = _MarginJudge(0.4)
    n_trig = n_lab = n_sil = 0
    worst = [0.0, 0.0]
    for c in range(chunks):
        blk = pcm[:, c * CHUNK:(c + 1) * CHUNK]
        trig, probs, nfr = det.step(blk, want_probs=True)
        state_before_reset_unknown = det.state().cpu().numpy()
        labels, counts = det.window_labels()
        judge.gpu = probs
        want = orc.step(blk, decide_on=judge)
        assert (nfr == want["softmax"].shape[1]).all()
        pe, se = _check_chunk(want, trig, probs, labels, counts, state_before_reset_unknown, c)
        worst = [max(worst[0], pe), max(worst[1], se)]
        n_trig += int(trig.sum())
        n_sil += int((~want["speech"]).sum())
        n_lab += sum(len(l) // 2 for l in want["labels"])
    frac = judge.ambiguous / judge.frames
    with capsys.disabled():
        print("\n[tc margin parity] %d streams x %d chunks: %d triggers, %d labels, %d VAD resets; ambiguous frames "
              "%d / %d = %.4f%%; worst |dp| %.2e |dh| %.2e" % (S, chunks, n_trig, n_lab, n_sil, judge.ambiguous,
                                                             judge.frames, 100 * frac, worst[0], worst[1]))
    assert n_trig > 50 and n_lab > 200 and n_sil > 500, (n_trig, n_lab, n_sil)
    assert frac < 0.02, frac                               # decisions within 1e-3 of a threshold are rare
    det.close()
    dm.close()


def test_configs2_size_131072_streams_tc_against_oracle_sample(capsys):
    """BASELINE configs[2] at its own size: 131,072 streams, tensor-core precision, 20 chunks with VAD silences.
    The 131,072 streams replay 1,024 distinct recordings; the detector-loop oracle runs on a 512-stream sample that
    covers the first tile, streams beyond 65,536, and the last tile of the persistent schedule.  Every other stream
    is checked through batch independence: it must reproduce, bit for bit, the sampled stream that plays the same
    recording."""
    import torch
    from keyword_spotting_b200 import DeployModel, StreamingDetector
    from oracle import model as om, streaming as ost
    ow = _boost_fc(om.init_weights(seed=1234, n_mel=40))
    dm = DeployModel(make_config(40), to_product_weights(ow))
    S, B, chunks = 131072, 1024, 20
    rng = np.random.default_rng(2026)
    base = _chunked_pcm_with_silences(rng, B, chunks)
    # recording played by every stream; the sampled streams play recordings 0..511, one each
    idx = rng.integers(0, B, S)
    sample = np.concatenate([np.arange(0, 128), np.arange(65536, 65536 + 128), np.arange(S - 128, S),
                             65536 + 128 + rng.choice(S - 65536 - 256, 128, replace=False)])
    assert len(np.unique(sample)) == 512 and (sample >= 65536).sum() >= 384
    idx[sample] = np.arange(512)
    rep_of_rec = np.full(B, -1, np.int64)                  # the stream that represents each recording
    uniq, first = np.unique(idx, return_index=True)
    rep_of_rec[uniq] = first
    rep_of_rec[:512] = sample
    rep_np = rep_of_rec[idx]                               # for every stream: its representative
    assert (rep_np >= 0).all() and (idx[rep_np] == idx).all()
    rep = torch.from_numpy(rep_np).cuda()
    idx_d = torch.from_numpy(idx).cuda()
    sample_d = torch.from_numpy(sample).cuda()
    base_d = torch.from_numpy(base).cuda()
    det = StreamingDetector(dm, S, keyword="1")
    orc = ost.StreamOracle(ow, 512, label="1")
    judge = _MarginJudge(0.4)
    n_trig_all = n_trig = 0
    worst = [0.0, 0.0]
    for c in range(chunks):
        blk = base_d[:, c * CHUNK:(c + 1) * CHUNK][idx_d].contiguous()           # [131072, 4800] int16
        trig, probs, nfr = det.step(blk, want_probs=True)
        state = det.state()
        # batch independence over all 131,072 streams (tiles, SMs and schedule position must not matter)
        assert torch.equal(trig, trig[rep]), c
        assert torch.equal(probs, probs[rep]), c
        assert torch.equal(state, state[:, rep]), c
        labels, counts = det.window_labels(max_labels=128)
        assert np.array_equal(counts, counts[rep_np]) and np.array_equal(labels, labels[rep_np]), c
        judge.gpu = probs[sample_d].cpu().numpy()
        want = orc.step(base[:512, c * CHUNK:(c + 1) * CHUNK], decide_on=judge)
        pe, se = _check_chunk(want, trig[sample_d].cpu().numpy(), judge.gpu, labels[sample], counts[sample],
                              state[:, sample_d].cpu().numpy(), c)
        worst = [max(worst[0], pe), max(worst[1], se)]
        n_trig += int(want["trigger"].sum())
        n_trig_all += int(trig.sum())
        del blk, probs, trig, labels, counts
    frac = judge.ambiguous / judge.frames
    with capsys.disabled():
        print("\n[configs[2] size] 131072 streams x %d chunks: %d triggers (%d in the 512-stream oracle sample); "
              "ambiguous frames %.4f%%; worst |dp| %.2e |dh| %.2e" % (chunks, n_trig_all, n_trig, 100 * frac, *worst))
    assert n_trig > 100 and n_trig_all > 10000
    assert frac < 0.02
    det.close()
    dm.close()


def test_tc_drift_2000_chunks_without_reset(capsys):
    """60,000 recurrent steps with carried state and no reset (never silent, never triggered): the tensor-core
    server against the float64 graph.  The GRU contracts rounding noise instead of accumulating it."""
    from keyword_spotting_b200 import DeployModel, StreamingDetector
    from oracle import model as om, streaming as ost
    ow = om.init_weights(seed=1234, n_mel=40)
    dm = DeployModel(make_config(40), to_product_weights(ow))
    S, chunks = 4, 2000
    rng = np.random.default_rng(4242)
    det = StreamingDetector(dm, S, keyword="1233", decode_thres=0.999)          # never fires
    st64 = np.zeros((2, S, 128))
    res = np.zeros((S, 0), np.float32)
    worst_p = worst_s = 0.0
    t = np.arange(CHUNK) / 16000.0
    for c in range(chunks):
        sigma = np.exp(rng.uniform(np.log(300.0), np.log(3000.0), size=(S, 1)))
        x = rng.standard_normal((S, CHUNK)) * sigma
        x += rng.uniform(0, 8000.0, (S, 1)) * np.sin(2 * np.pi * rng.uniform(200, 4000, (S, 1)) * t)
        blk = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
        trig, probs, nfr = det.step(blk, want_probs=True)
        assert not trig.any()
        full = np.concatenate([res, om.pcm16_to_float(blk)], 1)
        res = full[:, -ost.residual_length(full.shape[1]):]
        p64, st64, _ = om.deploy_forward(full.astype(np.float64), st64, ow, np.float64)
        worst_p = max(worst_p, float(np.abs(probs[:, :p64.shape[1]] - p64).max()))
        if c % 50 == 49 or c == chunks - 1:
            worst_s = max(worst_s, float(np.abs(det.state().cpu().numpy() - st64).max()))
            assert worst_p < TOL_CONTRACT and worst_s < TOL_CONTRACT, (c, worst_p, worst_s)
    with capsys.disabled():
        print("\n[tc drift] %d chunks (%d steps) without reset: worst |dp| %.2e |dh| %.2e vs float64"
              % (chunks, chunks * 30, worst_p, worst_s))
    det.close()
    dm.close()


@pytest.mark.parametrize("C", [9, 12, 16])
def test_more_than_8_classes_run_on_the_exact_kernel(C):
    """The tensor-core kernel keeps 8 FC columns; wider models must not be truncated (they run on the fp32 kernel
    whatever the requested precision)."""
    from keyword_spotting_b200 import Config, DeployModel
    from oracle import model as om
    cfg = Config(n_mel=40, label_dict={"w%d" % i: i + 1 for i in range(C - 3)})
    assert cfg.num_classes == C
    ow = om.init_weights(seed=7, n_mel=40, num_classes=C)
    rng = np.random.default_rng(3)
    mel = np.abs(rng.standard_normal((70, 9, 40))).astype(np.float32)
    st = (rng.uniform(-1, 1, (2, 70, 128)) * 0.5).astype(np.float32)
    p_want, s_want, _ = om.mel_forward(mel, st, ow)
    pcm = synth_pcm16(rng, 5, 5120, silent_frac=0.0)
    pd_want, sd_want, _ = om.deploy_forward(om.pcm16_to_float(pcm), np.zeros((2, 5, 128), np.float32), ow)
    for precision in ("tc", "fp32"):
        dm = DeployModel(cfg, to_product_weights(ow), precision=precision)
        p, s = dm.run_mel(mel, st)
        assert p.shape == (70, 9, C)
        assert np.abs(p - p_want).max() < 1e-4 and np.abs(s - s_want).max() < 1e-4
        np.testing.assert_allclose(p.sum(-1), 1.0, atol=1e-5)
        pd, sd = dm(pcm, np.zeros((2, 5, 128), np.float32))
        assert np.abs(pd - pd_want).max() < 1e-4 and np.abs(sd - sd_want).max() < 1e-4
        dm.close()


@pytest.mark.parametrize("layers,n_mel", [(1, 40), (1, 24), (3, 40), (2, 64), (2, 80), (4, 16)])
def test_tensor_core_kernel_instances_other_depths_and_widths(layers, n_mel):
    """Every instance of the tensor-core recurrent kernel: a first layer that is also the last one (bulk-copied x
    operand AND the FC on the tensor core), middle layers, mel counts that take the run-time trip counts, a split and an
    unsplit x product -- from PCM (the front end writes the operand tiles) and from row-major mel (packed on the fly)."""
    from keyword_spotting_b200 import Config, DeployModel
    from oracle import model as om
    cfg = Config(n_mel=n_mel, num_layers=layers)
    ow = om.init_weights(seed=11 + layers, n_mel=n_mel, num_layers=layers)
    rng = np.random.default_rng(100 * layers + n_mel)
    dm = DeployModel(cfg, to_product_weights(ow), precision="tc")
    assert dm.precision == "tc"
    for S, n in [(3, 30), (130, 9), (257, 30)]:
        # (above 64 mels the x product is not split into hi/lo operands -- the wider weights do not fit next to the operand
        # tile -- and holds the contract for moderate levels only: DESIGN.md, numerics)
        mel = (np.abs(rng.standard_normal((S, n, n_mel))) * (rng.uniform(0.2, 5.0) if n_mel <= 64 else 0.5)).astype(np.float32)
        st = (rng.uniform(-1, 1, (layers, S, 128)) * 0.7).astype(np.float32)
        p_want, s_want, l_want = om.mel_forward(mel, st, ow, dtype=np.float32)
        p_emu, s_emu, _ = om.mel_forward(mel, st, ow, dtype=np.float32, operand_dtype=np.float16)
        p, s, lg = dm.run_mel(mel, st, want_logits=True)
        # the kernel's LOGIC, against the oracle run with the same operand rounding (measured: 2.3e-4 / 1.5e-4)
        if n_mel <= 64:                      # the emulation models the split x product (what layer 0 runs up to 64 mels)
            assert np.abs(p - p_emu).max() < 6e-4 and np.abs(s - s_emu).max() < 4e-4, (S, n)
        # the 1e-3 contract is stated for the deployment shape (2 layers, 40 mels); deeper / wider models accumulate more
        # fp16-operand rounding on these loud inputs (the emulation shows the same 1.2e-3): bounded, not contracted
        tol = 1e-3 if layers <= 2 and n_mel <= 40 else 3e-3
        assert np.abs(p - p_want).max() < tol and np.abs(s - s_want).max() < tol, (S, n)
        np.testing.assert_allclose(p.sum(-1), 1.0, atol=1e-5)
    pcm = synth_pcm16(rng, 131, 5120, silent_frac=0.1)
    st0 = np.zeros((layers, 131, 128), np.float32)
    pd_want, sd_want, _ = om.deploy_forward(om.pcm16_to_float(pcm), st0, ow)
    pd, sd = dm(pcm, st0)
    assert np.abs(pd - pd_want).max() < 3e-3 and np.abs(sd - sd_want).max() < 3e-3
    dm.close()
