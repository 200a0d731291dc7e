"""Mel filterbank for the deployment graph's constant (models/rnn_ctc.py:139-146).

The reference calls ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` (Slaney
scale, area-normalised triangles -- librosa's defaults of that era) once at
graph-build time; librosa is not a dependency here, so the same published
construction is done in numpy.  Plan-time host code: the basis is a weight that
is uploaded once, not part of the per-chunk path.
"""
from __future__ import annotations

import numpy as np

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(freq):
    freq = np.atleast_1d(np.asarray(freq, dtype=np.float64))
    mel = freq / _F_SP
    big = freq >= _MIN_LOG_HZ
    mel[big] = _MIN_LOG_MEL + np.log(freq[big] / _MIN_LOG_HZ) / _LOGSTEP
    return mel


def mel_to_hz(mel):
    mel = np.atleast_1d(np.asarray(mel, dtype=np.float64))
    freq = _F_SP * mel
    big = mel >= _MIN_LOG_MEL
    freq[big] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mel[big] - _MIN_LOG_MEL))
    return freq


def mel_filterbank(sr=16000, n_fft=400, n_mels=40, fmin=300.0, fmax=8000.0):
    """``[n_mels, 1 + n_fft//2]`` float64, equal to ``librosa.filters.mel`` (Slaney, norm=1)."""
    n_bins = 1 + n_fft // 2
    fft_freqs = np.linspace(0.0, sr / 2.0, n_bins)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin)[0], hz_to_mel(fmax)[0], n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - fft_freqs[None, :]
    bank = np.zeros((n_mels, n_bins))
    for b in range(n_mels):
        rising = -ramps[b] / width[b]
        falling = ramps[b + 2] / width[b + 1]
        bank[b] = np.clip(np.minimum(rising, falling), 0.0, None)
    bank *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return bank
