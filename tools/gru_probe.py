"""Time the recurrent part alone (both GRU layers + FC + softmax: DeployModel.run_mel) at the benchmark size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keyword_spotting_b200 import Config, DeployModel, ModelWeights

S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
cfg = Config(n_mel=40)
dm = DeployModel(cfg, ModelWeights.random_init(cfg, seed=1234), device=0, precision="tc")
mel = torch.rand((S, 30, 40), device="cuda") * 3
st = torch.zeros((2, S, 128), device="cuda")
for _ in range(3):
    out = dm.run_mel(mel, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    out = dm.run_mel(mel, st)
e1.record()
torch.cuda.synchronize()
print("%s run_mel S=%d: %.3f ms" % (os.environ.get("KWS_B200_LIB", "default"), S, e0.elapsed_time(e1) / n))
