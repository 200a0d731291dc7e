"""Build-at-import of libkws_b200.so, the way the reference builds its ops
(octbit/op_compile.py:27-79: compile the sources next to the wrapper, then load
the .so) -- with nvcc for sm_100a instead of g++, and ctypes instead of
tf.load_op_library.  The library is built IN-TREE so it travels with the repo
snapshot; there is no CPU fallback: if it can neither be found nor built the
import fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libkws_b200.so")
SOURCES = ["api.cu", "posenc.cu", "octbit.cu", "octbit_tc.cu", "frontend.cu", "gru.cu", "gru_tc.cu", "decode.cu", "stream.cu", "tc_debug.cu", "attention.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def _inputs():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "kws_b200.h"))
    return files


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in _inputs() if os.path.exists(f))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into keyword_spotting_b200/libkws_b200.so."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH          # stale but usable (e.g. mtimes scrambled by a copy)
        raise ImportError("libkws_b200.so is not built and nvcc was not found; "
                          "keyword_spotting_b200 has no CPU fallback")
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise ImportError("nvcc failed building libkws_b200.so:\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH
