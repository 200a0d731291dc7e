"""Per-case deviation of the tensor-core recurrent kernel from the oracle (fp32 graph, and the fp16-operand emulations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import model as om
from tests._util import make_config, to_product_weights
from keyword_spotting_b200 import DeployModel

for M in (40, 60):
    ow = om.init_weights(seed=1234, n_mel=M)
    dm = DeployModel(make_config(M), to_product_weights(ow), precision="tc")
    rng = np.random.default_rng(77)
    for S, n in [(1, 1), (1, 30), (5, 30), (64, 7), (65, 30), (130, 3), (3, 298), (300, 30)]:
        mel = (np.abs(rng.standard_normal((S, n, ow.n_mel))) * rng.uniform(0.1, 6.0)).astype(np.float32)
        st = (rng.uniform(-1, 1, (2, S, 128)) * 0.8).astype(np.float32)
        p32, s32, _ = om.mel_forward(mel, st, ow, dtype=np.float32)
        pe, se, _ = om.mel_forward(mel, st, ow, dtype=np.float32, operand_dtype=np.float16)
        pg, sg, lg = dm.run_mel(mel, st, want_logits=True)
        print("M=%d S=%d n=%d: vs fp32 p %.2e s %.2e | vs emu p %.2e s0 %.2e s1 %.2e" % (
            M, S, n, np.abs(pg - p32).max(), np.abs(sg - s32).max(), np.abs(pg - pe).max(),
            np.abs(sg[0] - se[0]).max(), np.abs(sg[1] - se[1]).max()))
    dm.close()
