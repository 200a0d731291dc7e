import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import model as om
from tests._util import to_product_weights, synth_pcm16
from keyword_spotting_b200 import Config, DeployModel
for layers,n_mel in [(2,64),(4,16),(2,40),(1,24)]:
    cfg = Config(n_mel=n_mel, num_layers=layers)
    ow = om.init_weights(seed=11 + layers, n_mel=n_mel, num_layers=layers)
    rng = np.random.default_rng(100 * layers + n_mel)
    dm = DeployModel(cfg, to_product_weights(ow), precision="tc")
    for S, n in [(3, 30), (130, 9), (257, 30)]:
        mel = (np.abs(rng.standard_normal((S, n, n_mel))) * (rng.uniform(0.2, 5.0) if n_mel <= 64 else 0.5)).astype(np.float32)
        st = (rng.uniform(-1, 1, (layers, S, 128)) * 0.7).astype(np.float32)
        p_want, s_want, _ = om.mel_forward(mel, st, ow, dtype=np.float32)
        p_emu, s_emu, _ = om.mel_forward(mel, st, ow, dtype=np.float32, operand_dtype=np.float16)
        p, s = dm.run_mel(mel, st)
        print(layers, n_mel, S, n, "vs fp32 p %.2e s %.2e | emu-vs-fp32 p %.2e s %.2e | vs emu p %.2e s %.2e" % (
            np.abs(p-p_want).max(), np.abs(s-s_want).max(), np.abs(p_emu-p_want).max(), np.abs(s_emu-s_want).max(),
            np.abs(p-p_emu).max(), np.abs(s-s_emu).max()), flush=True)
    dm.close()
