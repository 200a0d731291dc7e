"""GPU: pin the tcgen05 conventions (TMEM A operand, smem B descriptor, TMEM accumulator read-back)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ss_mode", [0, 1, 2, 3])   # 2: B operand MN-major; 3: A row-tiled in HBM, bulk-copied to smem (GRU x operand)
@pytest.mark.parametrize("N,K", [(16, 32), (32, 256), (128, 128), (256, 256), (128, 160), (256, 64)])
def test_tc_gemm_layouts(N, K, ss_mode):
    import torch
    from keyword_spotting_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(N * 1000 + K + ss_mode)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    # distinguishable structure: a wrong layout cannot pass by symmetry
    A[:, 0] += np.arange(128) * 0.25
    B[:, 1] += np.arange(N) * 0.125
    a = torch.from_numpy(A).cuda()
    d = torch.zeros((128, N), dtype=torch.float32, device="cuda")
    _lib.check(lib.kws_debug_tc_gemm(a.data_ptr(), B.ctypes.data, d.data_ptr(), N, K, ss_mode,
                                     torch.cuda.current_stream().cuda_stream))
    want = A.astype(np.float16).astype(np.float64) @ B.astype(np.float16).astype(np.float64).T
    got = d.cpu().numpy()
    err = np.abs(got - want).max()
    assert err < 2e-4 * max(1.0, np.abs(want).max()), (N, K, ss_mode, err)
