import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from keyword_spotting_b200 import Config, DeployModel, ModelWeights
from keyword_spotting_b200.rnn_ctc import OctbitModelWeights
cfg = Config(n_mel=40)
w = ModelWeights.random_init(cfg, seed=1234)
dm = DeployModel(cfg, w, device=0)
dm.set_octbit(OctbitModelWeights.from_float(w))
S = 32768
mel = torch.rand((S, 30, 40), device="cuda") * 3
st = torch.zeros((2, S, 128), device="cuda")
for _ in range(2): dm.run_mel(mel, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): dm.run_mel(mel, st)
e1.record(); torch.cuda.synchronize()
print("X1 run_mel S=%d: %.2f ms" % (S, e0.elapsed_time(e1) / 3))
