"""GPU: attention_ctc deployment forward (BASELINE config 5) vs the numpy restatement (oracle/attention.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-3          # the contract's tolerance for floating-point outputs; the fp32 kernels land near 1e-5


@pytest.fixture(scope="module")
def att():
    from keyword_spotting_b200 import AttentionConfig, AttentionDeployModel, AttentionWeights
    from oracle import attention as oa
    ow = oa.init_weights(seed=4321, n_mel=60)
    m = AttentionDeployModel(AttentionConfig(), AttentionWeights.from_object(ow))
    yield ow, m
    m.close()


# T' = T // 2 + 1 keys.  The tensor-core attention core takes T' <= 400: 510 -> 256 keys (one score MMA), 512 -> 257 keys
# (padded to 272: a second MMA of 16 columns), 766 -> 384 (whole query tiles), 798 -> 400 (the full TMEM budget);
# 1000 -> 501 keys runs the FFMA attention core behind the tensor-core dense layers.
@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (2, 37), (3, 100), (1, 255), (2, 798), (5, 256), (1, 510), (2, 512), (3, 766),
                                 (2, 1000)])
def test_attention_mel_forward_matches_oracle(att, B, T):
    from oracle import attention as oa
    ow, m = att
    rng = np.random.default_rng(B * 1000 + T)
    mel = (np.abs(rng.standard_normal((B, T, 60))) * rng.uniform(0.2, 3.0)).astype(np.float32)
    want, lwant = oa.mel_forward(mel, ow, np.float32)
    want64, _ = oa.mel_forward(mel, ow, np.float64)
    got, lgot = m.run_mel(mel, want_logits=True)
    assert got.shape == want.shape == (B, T // 2 + 1, 6)
    assert np.abs(got - want64).max() < TOL, np.abs(got - want64).max()
    assert np.abs(got - want).max() < 1e-4
    assert np.abs(lgot - lwant).max() < 1e-3
    assert np.abs(got.sum(-1) - 1).max() < 1e-5


def test_attention_pcm_deploy_call_and_batch_independence(att):
    """PCM through the shared front end, the sess.run-style call, and utterances of a batch do not interact."""
    from oracle import attention as oa, model as om
    from tests._util import synth_pcm16
    ow, m = att
    rng = np.random.default_rng(9)
    pcm16 = synth_pcm16(rng, 3, 16000, silent_frac=0.0)
    pcm = om.pcm16_to_float(pcm16)
    want, _ = oa.deploy_forward(pcm, ow)
    got = m.run(["model/softmax:0"], {"model/inputX:0": pcm})[0]
    assert got.shape == want.shape
    assert np.abs(got - want).max() < TOL
    one = m(pcm[1])                                            # the reference's batch-1 form: [L] -> [1, T', C]
    assert one.shape == (1,) + want.shape[1:]
    assert np.abs(one[0] - got[1]).max() < 1e-5
    g16 = m(pcm16)                                             # int16 input
    assert np.abs(g16 - got).max() < 1e-4
    with pytest.raises(ValueError):
        m.run(["model/logit:0"], {"model/inputX:0": pcm})


def test_attention_config5_full_length_properties(att):
    """8 s utterances (T = 798 -> T' = 400): softmax rows sum to one, permuting the batch permutes the output,
    repeated utterances give identical rows."""
    import torch
    ow, m = att
    g = torch.Generator(device="cuda").manual_seed(5)
    mel = torch.rand((48, 798, 60), device="cuda", generator=g) * 2
    mel[7] = mel[3]
    out = m.run_mel(mel)
    assert out.shape == (48, 400, 6)
    assert float((out.sum(-1) - 1).abs().max()) < 1e-5
    assert torch.equal(out[7], out[3])
    perm = torch.randperm(48, device="cuda", generator=g)
    out_p = m.run_mel(mel[perm].contiguous())
    assert float((out_p - out[perm]).abs().max()) < 1e-6
