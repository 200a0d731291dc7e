import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keyword_spotting_b200 import AttentionConfig, AttentionDeployModel
am = AttentionDeployModel(AttentionConfig())
g = torch.Generator(device="cuda").manual_seed(1)
mel = torch.rand((256, 798, 60), device="cuda", generator=g) * 2
for _ in range(3):
    am.run_mel(mel)
torch.cuda.synchronize()
