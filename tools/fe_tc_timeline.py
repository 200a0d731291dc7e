import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from keyword_spotting_b200 import Config, DeployModel
dm = DeployModel(Config(n_mel=40), frontend="tc")
x = (torch.randn((148*60, 5120), device='cuda')*800).to(torch.int16)
for _ in range(2): dm.frontend(x); torch.cuda.synchronize()
