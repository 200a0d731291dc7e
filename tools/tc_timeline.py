"""Print the per-phase timeline of the tensor-core GRU kernel (first tile of CTA 0), in SM clock ticks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from keyword_spotting_b200 import Config, DeployModel, ModelWeights, _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072 // 4
LAYER = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # which layer's launch records the timeline
cfg = Config(n_mel=40)
dm = DeployModel(cfg, ModelWeights.random_init(cfg, seed=1234), device=0, precision="tc")
lib = _lib.load()
mel = torch.rand((S, 30, 40), device="cuda") * 3
st = torch.zeros((2, S, 128), device="cuda")
for _ in range(2):
    dm.run_mel(mel, st)
lib.kws_debug_tc_timeline(2 + LAYER, None, 0)
dm.run_mel(mel, st)
torch.cuda.synchronize()
buf = np.zeros(64 * 8, np.int64)
lib.kws_debug_tc_timeline(0, buf.ctypes.data, buf.size)
tl = buf.reshape(64, 8)[:30]
names = ["waitR", "epi_r", "waitU+st", "epi_u(+ldx)", "waitC", "epi_c", "out/FC"]
print("layer %d.  per-step phase durations (ticks):" % LAYER)
for t in range(30):
    row = tl[t]
    nxt = tl[t + 1][0] if t + 1 < 30 else tl[0][7]
    d = [row[1] - row[0], row[2] - row[1], row[3] - row[2], row[4] - row[3], row[5] - row[4], row[6] - row[5], nxt - row[6]]
    print(t, dict(zip(names, [int(v) for v in d])), "step", int(nxt - row[0]))
    f = buf[256 + 4 * t: 256 + 4 * t + 3]
    print("     fc: partial %d  wait-others %d  softmax+store %d" % (f[0] - row[6], f[1] - f[0], f[2] - f[1]))
