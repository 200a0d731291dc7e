"""Shared helpers for the test-suite (synthetic inputs, golden unpacking)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def unpack(flat, off, i, width=None):
    a = flat[off[i]:off[i + 1]]
    return a.reshape(-1, width) if width else a


def synth_pcm16(rng, n_streams, n_samples, silent_frac=0.3):
    """SURVEY.md 8d audio: per-stream noise with log-uniform sigma in [30, 3000] LSB plus up to three
    amplitude-modulated tone bursts; a fraction of the streams is near-silent (below the VAD threshold)."""
    t = np.arange(n_samples) / 16000.0
    sigma = np.exp(rng.uniform(np.log(30.0), np.log(3000.0), size=(n_streams, 1)))
    x = rng.standard_normal((n_streams, n_samples)) * sigma
    for s in range(n_streams):
        for _ in range(int(rng.integers(0, 4))):
            f = rng.uniform(200.0, 4000.0)
            a = rng.uniform(500.0, 12000.0)
            c = rng.uniform(0, t[-1] if n_samples > 1 else 0)
            w = rng.uniform(0.02, 0.2)
            x[s] += a * np.exp(-0.5 * ((t - c) / w) ** 2) * np.sin(2 * np.pi * f * t)
    silent = rng.random(n_streams) < silent_frac
    x[silent] = rng.integers(-3, 4, size=(int(silent.sum()), n_samples))
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def to_product_weights(ow):
    """oracle.model.Weights -> keyword_spotting_b200.ModelWeights (same arrays, same layout)."""
    from keyword_spotting_b200 import ModelWeights
    return ModelWeights.from_arrays(ow.mel_basis, ow.gates_kernel, ow.gates_bias, ow.cand_kernel,
                                    ow.cand_bias, ow.fc_w, ow.fc_b)


def make_config(n_mel=40):
    from keyword_spotting_b200 import Config
    return Config(n_mel=n_mel)
