"""Error of the tensor-core front end vs float64 by signal type (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from keyword_spotting_b200 import Config, DeployModel
from oracle import model as om
from tests._util import to_product_weights

ow = om.init_weights(seed=1234, n_mel=40)
dm = DeployModel(Config(n_mel=40), to_product_weights(ow), frontend="tc")
rng = np.random.default_rng(1)
L = 48000
t = np.arange(L) / 16000.0
sig = {
    "noise_3000": rng.standard_normal(L) * 3000,
    "noise_fullscale_clipped": rng.standard_normal(L) * 30000,
    "tone_1k_12000": 12000 * np.sin(2 * np.pi * 1000 * t),
    "tone_997_30000": 30000 * np.sin(2 * np.pi * 997.3 * t),
    "dc_20000": np.full(L, 20000.0),
    "dc_plus_noise": 15000 + rng.standard_normal(L) * 100,
    "square_32767": 32767 * np.sign(np.sin(2 * np.pi * 440 * t)),
    "quiet_30": rng.standard_normal(L) * 30,
}
for name, x in sig.items():
    pcm16 = np.clip(np.rint(x), -32768, 32767).astype(np.int16)[None, :]
    want = om.pcm_to_mel(om.pcm16_to_float(pcm16).astype(np.float64), ow, np.float64)
    for L_use in (5120, 48000):
        got = dm.frontend(pcm16[:, :L_use])
        w = want[:, :got.shape[1]]
        err = np.abs(got - w)
        print("%-26s L=%5d  max|mel| %.4g  abs err %.3e  rel-to-max %.3e" % (name, L_use, np.abs(w).max(), err.max(), err.max() / max(1e-12, np.abs(w).max())), flush=True)
os.environ["X"] = "1"
