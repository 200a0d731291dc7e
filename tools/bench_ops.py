"""Device timings of the custom ops (K4 decode, K5 octbit, K6 posenc) at BASELINE sizes -> one JSON object.

    python tools/bench_ops.py            # needs a CUDA device
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from keyword_spotting_b200 import _lib
from keyword_spotting_b200.octbit.octbit_ops import octbit_mat_mul
from keyword_spotting_b200.octbit.octbit_graph import octize_weight_int8_signed
from keyword_spotting_b200.positional_encoding.positional_encoding_op import positional_encoding
from keyword_spotting_b200.utils import prediction


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


out = {}
rng = np.random.default_rng(0)
g = torch.Generator(device="cuda").manual_seed(1)
# ---- octbit: the three converted matmuls of the rnn_ctc graph at A = 131072 streams x 30 frames
A = 131072 * 30
for K, B in ((256, 256), (256, 128), (128, 6)):
    x = torch.randn((A, K), device="cuda", generator=g)
    # weights as the graph rewriter produces them: Xavier-normal fp32 -> octize_weight_int8_signed
    wf = torch.randn((K, B), device="cuda", generator=g) * (2.0 / (K + B)) ** 0.5
    w, scale, bias = octize_weight_int8_signed(wf)
    ms = timeit(lambda: octbit_mat_mul(x, w, scale=scale, bias=bias), iters=5)
    nbytes = A * K * 4 + B * K + A * B * 4
    out["octbit_A%d_K%d_B%d" % (A, K, B)] = dict(ms=ms, algorithmic_GBps=nbytes / ms / 1e6, int8_TOPs=2.0 * A * B * K / ms / 1e9)
    del x
# ---- posenc: config 5 table and a large one
for T, N in ((400, 128), (65536, 512)):
    ms = timeit(lambda: positional_encoding(T, N), iters=20)
    out["posenc_%dx%d" % (T, N)] = dict(ms=ms, GBps=T * N * 4 / ms / 1e6)
# ---- decode: config 2, 4096 utterances x 298 frames, and the streaming window size
lib = _lib.load()
for S, T in ((4096, 298), (131072, 450)):
    p = torch.rand((S, T, 6), device="cuda", generator=g)
    p = p / p.sum(-1, keepdim=True)
    for name, mode in (("ctc_decode", prediction.MODE_CTC_DECODE), ("ctc_decode2", prediction.MODE_CTC_DECODE2)):
        ms = timeit(lambda: prediction.decode_batch(p, mode=mode, want_labels=False), iters=5)
        out["%s_%dx%d" % (name, S, T)] = dict(ms=ms, GBps=S * T * 6 * 4 / ms / 1e6)
# ---- BASELINE configs[1]: offline batched inference, 4096 utterances x 3 s of int16 PCM already in HBM ->
# front end -> 2L GRU-128 + FC + softmax (n = 298 frames) -> ctc_decode labels + trigger (main.py:263-314)
from keyword_spotting_b200 import Config, DeployModel, ModelWeights
from keyword_spotting_b200.rnn_ctc import INPUT_X, INITIAL_STATES, SOFTMAX, RNN_STATES
cfg = Config(n_mel=40)
dm = DeployModel(cfg, ModelWeights.random_init(cfg, seed=1234))
for S in (4096, 32768):
    pcm16 = (torch.randn((S, 48000), device="cuda", generator=g) * 800).clamp_(-32768, 32767).to(torch.int16)
    st0 = torch.zeros((2, S, 128), device="cuda")

    def offline():
        probs, _ = dm.run([SOFTMAX, RNN_STATES], {INPUT_X: pcm16, INITIAL_STATES: st0})
        return prediction.decode_batch(probs, mode=prediction.MODE_CTC_DECODE, want_labels=True)

    ms = timeit(offline, iters=5)
    ms_fe = timeit(lambda: dm.frontend(pcm16), iters=5)
    out["offline_config2_S%d_3s" % S] = dict(ms=ms, ms_frontend=ms_fe, utterances_per_s=S / ms * 1e3, audio_s_per_s=3.0 * S / ms * 1e3,
                                             gru_tiles=S // 128)
    del pcm16, st0
dm.close()
# ---- attention_ctc forward, config 5: batched 8 s utterances (T = 798 mel frames of 60 bands -> T' = 400)
from keyword_spotting_b200 import AttentionConfig, AttentionDeployModel
am = AttentionDeployModel(AttentionConfig())
for B in (256, 1024):
    mel = torch.rand((B, 798, 60), device="cuda", generator=g) * 2
    ms = timeit(lambda: am.run_mel(mel), iters=3, warm=2)
    flop = B * (400 * 120 * 128 * 2 + 3 * (400 * 128 * 384 * 2 + 2 * 8 * 400 * 400 * 16 * 2 + 2 * 400 * 128 * 512 * 2) + 400 * 128 * 6 * 2)
    out["attention_mel_forward_B%d_8s" % B] = dict(ms=ms, utterances_per_s=B / ms * 1e3, audio_s_per_s=8.0 * B / ms * 1e3,
                                                   TFLOPs=flop / ms / 1e9)
pcm = (torch.randn((256, 128000), device="cuda", generator=g) * 0.05)
ms = timeit(lambda: am(pcm), iters=3, warm=2)
out["attention_pcm_forward_B256_8s"] = dict(ms=ms, audio_s_per_s=8.0 * 256 / ms * 1e3)
am.close()
print(json.dumps(out, indent=1))
