"""CPU: the C-ABI library loads and exports what include/kws_b200.h declares; device logic that is
written as host/device code (FFT team, decoders) is checked on the CPU; the product refuses to run
without a GPU instead of falling back."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from tests._util import golden, unpack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "kws_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kws_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from keyword_spotting_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), "libkws_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names        # the ctypes table binds exactly the header
    assert lib.kws_abi_version() == 2


def test_library_is_sm100a_only():
    from keyword_spotting_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import keyword_spotting_b200 as k
    from keyword_spotting_b200.octbit.octbit_ops import octbit_mat_mul
    from keyword_spotting_b200.positional_encoding.positional_encoding_op import positional_encoding
    from keyword_spotting_b200.utils import prediction
    with pytest.raises(RuntimeError):
        k.DeployModel()
    with pytest.raises(RuntimeError):
        octbit_mat_mul(np.zeros((1, 64), np.float32), np.zeros((1, 64), np.int8), scale=1.0, bias=[0])
    with pytest.raises(RuntimeError):
        positional_encoding(4, 8)
    with pytest.raises(RuntimeError):
        prediction.ctc_decode2(np.zeros((3, 6), np.float32), 6)
    lib = k._lib_mod.load()
    assert lib.kws_device_count() < 0 and "cudaGetDeviceCount" in k._lib_mod.last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "keyword_spotting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "/root/reference" not in src, f


def test_num_frames_abi_matches_stft():
    from keyword_spotting_b200 import _lib
    from oracle import model as om
    lib = _lib.load()
    for L in (0, 1, 239, 399, 400, 401, 559, 560, 4800, 5120, 48000, 128000):
        assert lib.kws_num_frames(None, L) == om.num_frames(L), L


@pytest.fixture(scope="module")
def host_libs(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostlibs")
    libs = {}
    for name in ("fft400_host", "decode_host"):
        so = str(d / (name + ".so"))
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host", name + ".cpp")])
        libs[name] = ctypes.CDLL(so)
    return libs


def test_fft_team_emulation_matches_numpy_rfft(host_libs):
    lib = host_libs["fft400_host"]
    rng = np.random.default_rng(0)
    for scale in (1e-3, 0.3, 1.0):
        win = (rng.standard_normal(560) * scale).astype(np.float32)
        ma = np.zeros(201, np.float32)
        mb = np.zeros(201, np.float32)
        lib.fft400_pair_mags(win.ctypes.data_as(ctypes.c_void_p), ma.ctypes.data_as(ctypes.c_void_p),
                             mb.ctypes.data_as(ctypes.c_void_p))
        ra = np.abs(np.fft.rfft(win[:400].astype(np.float64)))
        rb = np.abs(np.fft.rfft(win[160:560].astype(np.float64)))
        tol = 4e-7 * max(ra.max(), rb.max()) * 20           # ~ fp32 FFT noise: eps * sqrt(N) * max
        assert np.abs(ma - ra).max() < tol and np.abs(mb - rb).max() < tol


def test_device_decoder_logic_on_host_matches_reference_golden(host_libs):
    lib = host_libs["decode_host"]
    lib.decode_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                ctypes.c_double, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]

    def run(p, mode, lockout=3, thres=None, loose=0.2):
        p = np.ascontiguousarray(p, np.float32)
        T = p.shape[0]
        thres = {0: 0.5, 1: 0.4, 2: 0.5}[mode] if thres is None else thres
        out = np.full(2 * T + 3, -7, np.int32)
        trig = ctypes.c_int(0)
        n = lib.decode_host(p.ctypes.data, T, 6, mode, lockout, thres, loose, b"1233", out.ctypes.data, len(out), ctypes.byref(trig))
        assert (out[n:] == -1).all()
        return out[:n], trig.value

    g = golden("decode_golden.npz")
    po = g["probs_off"]
    for mode, k in ((0, "decode"), (1, "decode2"), (2, "strict")):
        for i in range(len(po) - 1):
            got, trig = run(unpack(g["probs"], po, i, 6), mode)
            np.testing.assert_array_equal(got, unpack(g[k], g[k + "_off"], i))
            assert trig == g[k + "_pred"][i]
    eo = g["extra_probs_off"]
    for i in range(len(eo) - 1):
        p = unpack(g["extra_probs"], eo, i, 6)
        lo, th, ls = int(g["extra_lockout"][i]), float(g["extra_thres"][i]), float(g["extra_loose"][i])
        for mode, k in ((0, "decode"), (1, "decode2"), (2, "strict")):
            got, _ = run(p, mode, lo, th, ls)
            np.testing.assert_array_equal(got, unpack(g["extra_" + k], g["extra_" + k + "_off"], i))


def test_mel_filterbank_equals_oracle_restatement():
    from keyword_spotting_b200.utils.mel import mel_filterbank
    from oracle.model import slaney_mel_basis
    for M in (40, 60):
        np.testing.assert_array_equal(mel_filterbank(n_mels=M), slaney_mel_basis(n_mels=M))


def test_mel_quad_lists_cover_the_basis_exactly():
    """The front end walks per-warp lists of 4-bin 'quads' instead of the dense [201, M] basis (csrc/frontend.cu,
    build_mel_quads; host code, no device needed).  Summing the quads back must reproduce every weight bit for bit,
    every band must be finished exactly once, quads start at even bins, and the 10 warps carry equal lengths."""
    import ctypes
    from keyword_spotting_b200 import _lib
    from oracle.model import slaney_mel_basis
    lib = _lib.load()
    rng = np.random.default_rng(11)
    bases = [slaney_mel_basis(n_mels=40).T, slaney_mel_basis(n_mels=60).T]
    dense = rng.standard_normal((201, 12)) * (rng.random((201, 12)) < 0.3)        # arbitrary sparse-ish basis
    dense[:, 5] = 0.0                                                              # an empty band still writes its 0
    bases.append(dense)
    for basis in bases:
        basis = np.ascontiguousarray(basis, np.float32)
        M = basis.shape[1]
        qpw = lib.kws_debug_mel_quads(basis.ctypes.data, M, None, 0)
        assert qpw >= 1 and qpw % 2 == 0
        rec = np.zeros((10 * qpw, 8), np.int32)
        assert lib.kws_debug_mel_quads(basis.ctypes.data, M, rec.ctypes.data, rec.shape[0]) == qpw
        w = rec[:, :4].copy().view(np.float32)
        rebuilt = np.zeros((204, M), np.float32)
        finished = np.zeros(M, int)
        for warp in range(10):
            band_rows = []
            for q in range(qpw):
                r = rec[warp * qpw + q]
                assert r[4] % (66 * 4) == 0                     # byte offset of a bin-pair row of 66 floats
                k0 = 2 * (r[4] // (66 * 4))
                band_rows.append((k0, w[warp * qpw + q]))
                if r[5] >= 0:
                    b = r[5] // 4
                    for k, wv in band_rows:
                        rebuilt[k:k + 4, b] += wv
                    finished[b] += 1
                    band_rows = []
            assert all((wv == 0).all() for _, wv in band_rows)  # trailing padding quads carry no weight
        np.testing.assert_array_equal(finished, 1)
        np.testing.assert_array_equal(rebuilt[:201], basis)
        assert (rebuilt[201:] == 0).all()


def test_weights_random_init_matches_oracle_recipe():
    from keyword_spotting_b200 import Config, ModelWeights
    from oracle import model as om
    for M in (40, 60):
        pw = ModelWeights.random_init(Config(n_mel=M), seed=1234)
        ow = om.init_weights(seed=1234, n_mel=M)
        np.testing.assert_array_equal(pw.mel_basis, ow.mel_basis)
        for l in range(2):
            np.testing.assert_array_equal(pw.gates_kernel[l], ow.gates_kernel[l])
            np.testing.assert_array_equal(pw.cand_kernel[l], ow.cand_kernel[l])
        np.testing.assert_array_equal(pw.fc_w, ow.fc_w)
        pw.validate(Config(n_mel=M))


def test_host_side_predict_and_evaluate_match_reference_golden():
    from keyword_spotting_b200.utils import prediction
    g = golden("decode_golden.npz")
    off = g["predict_cases_off"]
    for i in range(len(off) - 1):
        seq = unpack(g["predict_cases"], off, i)
        assert prediction.ctc_predict(seq, "1233") == g["predict_out"][i]
        assert prediction.ctc_predict(seq, "123") == g["predict_out_123"][i]
    assert tuple(prediction.evaluate(g["eval_result"].tolist(), g["eval_target"].tolist())) == tuple(int(v) for v in g["eval_out"])
