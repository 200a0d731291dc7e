"""Probe of the tensor-core front end against the float64 oracle on a few shapes (run on a GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from keyword_spotting_b200 import Config, DeployModel, ModelWeights
from oracle import model as om
from tests._util import synth_pcm16, to_product_weights

ow = om.init_weights(seed=1234, n_mel=40)
dm = DeployModel(Config(n_mel=40), to_product_weights(ow), frontend="tc")
rng = np.random.default_rng(5678)
shapes = [(1, 400), (1, 5120), (3, 5120), (7, 559), (2, 48000), (300, 5120)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for S, L in shapes:
    pcm16 = synth_pcm16(rng, S, L, silent_frac=0.2)
    want = om.pcm_to_mel(om.pcm16_to_float(pcm16).astype(np.float64), ow, np.float64)
    t0 = time.time()
    got = dm.frontend(pcm16)
    torch.cuda.synchronize()
    err = np.abs(got - want).max() / max(1e-12, np.abs(want).max())
    print("S=%d L=%d frames=%d rel err %.3e (%.2f s)" % (S, L, want.shape[1], err, time.time() - t0), flush=True)
