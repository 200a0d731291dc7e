import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keyword_spotting_b200.octbit.octbit_ops import octbit_mat_mul
from keyword_spotting_b200.octbit.octbit_graph import octize_weight_int8_signed
g = torch.Generator(device="cuda").manual_seed(1)
A, K, B = 131072 * 8, 256, 256
x = torch.randn((A, K), device="cuda", generator=g)
wf = torch.randn((K, B), device="cuda", generator=g) * (2.0 / (K + B)) ** 0.5
w, scale, bias = octize_weight_int8_signed(wf)
for _ in range(3):
    octbit_mat_mul(x, w, scale=scale, bias=bias)
torch.cuda.synchronize()
