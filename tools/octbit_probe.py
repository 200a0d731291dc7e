"""Octbit op timing at the model shapes (A = 131072 x 30 rows) + bit-exactness spot check (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from keyword_spotting_b200.octbit.octbit_ops import octbit_mat_mul
from keyword_spotting_b200.octbit.octbit_graph import octize_weight_int8_signed
from oracle import octbit as ooct

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

g = torch.Generator(device="cuda").manual_seed(1)
A = 131072 * 30
for K, B in ((256, 256), (256, 128), (128, 6)):
    x = torch.randn((A, K), device="cuda", generator=g)
    wf = torch.randn((K, B), device="cuda", generator=g) * (2.0 / (K + B)) ** 0.5
    w, scale, bias = octize_weight_int8_signed(wf)
    ms = timeit(lambda: octbit_mat_mul(x, w, scale=scale, bias=bias))
    nbytes = A * K * 4 + B * K + A * B * 4
    got = octbit_mat_mul(x[:4096], w, scale=scale, bias=bias).cpu().numpy()
    want = ooct.octbit_mat_mul(x[:4096].cpu().numpy(), w.cpu().numpy(), scale=np.float32(scale), bias=bias.cpu().numpy())
    print("K=%d B=%d: %.3f ms  %.0f GB/s algorithmic  bit-exact(4096 rows)=%s" % (K, B, ms, nbytes / ms / 1e6, got.tobytes() == want.tobytes()), flush=True)
    del x
