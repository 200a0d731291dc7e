"""Deployment constants of the rnn_ctc model -- the subset of config/rnn_config.py:20-103
that the inference path reads (the training/flag machinery is out of scope)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict


@dataclass
class Config:
    samplerate: int = 16000          # rnn_config.py:59
    fft_size: int = 400              # :57   25 ms
    hop_size: int = 160              # :58   10 ms
    fmin: float = 300.0              # :64
    fmax: float = 8000.0             # :65
    n_mel: int = 40                  # :63 ships 60; README.md:17-19 and BASELINE.json use 40
    num_layers: int = 2              # :76
    hidden_size: int = 128           # :84
    label_dict: Dict[str, int] = field(default_factory=lambda: {"ni3": 1, "hao3": 2, "le4": 3})  # :25
    label_seqs: str = "1233"         # :27
    # streaming loop constants (detector.py)
    chunk_samples: int = 4800        # 300 ms, README.md:88
    window_chunks: int = 15          # detector.py:122
    vad_threshold: int = 30          # detector.py:168
    decode_thres: float = 0.4        # utils/prediction.py:65

    @property
    def num_classes(self) -> int:    # rnn_config.py:87-91: words + space(0) + other(4) + blank(5)
        return len(self.label_dict) + 3

    @property
    def freq_size(self) -> int:      # rnn_config.py:97-99 with mfcc = False
        return self.n_mel


def get_config() -> Config:          # rnn_config.py:15-16
    return Config()
