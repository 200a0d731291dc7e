"""The HotwordDetector.start loop (detector.py:148-209) for S lock-step streams.

``StreamingDetector.step(chunk)`` is one pass of the reference's ``while True`` body
for every stream at once: VAD reset, tail carry, model call, 15-chunk window,
``ctc_decode2`` + ``ctc_predict``, trigger reset -- all on the device, all per-stream
state in HBM (kws_stream_* in include/kws_b200.h).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, _tensors
from .rnn_ctc import DeployModel


class StreamingDetector:
    def __init__(self, model: DeployModel, n_streams: int, max_chunk: int = None, window_chunks: int = None,
                 vad_threshold: int = None, decode_thres: float = None, keyword: str = None):
        cfg = model.config
        self.model = model
        self.device = model.device
        self.n_streams = int(n_streams)
        self.max_chunk = int(max_chunk if max_chunk is not None else cfg.chunk_samples)
        self._lib = _lib.load()
        sc = _lib.StreamConfig()
        sc.n_streams = self.n_streams
        sc.max_chunk = self.max_chunk
        sc.window_chunks = int(window_chunks if window_chunks is not None else cfg.window_chunks)
        sc.vad_threshold = int(vad_threshold if vad_threshold is not None else cfg.vad_threshold)
        sc.decode_thres = float(decode_thres if decode_thres is not None else cfg.decode_thres)
        sc.keyword = (keyword if keyword is not None else cfg.label_seqs).encode("ascii")
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kws_stream_create(model.handle, ctypes.byref(sc), ctypes.byref(handle)))
        self._handle = handle
        self.max_frames = int(self._lib.kws_stream_max_frames(handle))
        self._trigger = torch.zeros(self.n_streams, dtype=torch.int32, device=self.device)
        self._pinned_trigger = None

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.kws_stream_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kws_stream_reset(self._handle, _tensors.stream_ptr(self.device)))

    def step(self, chunk, want_probs=False):
        """chunk ``[S, n]`` int16 (numpy -> numpy results, CUDA tensor -> CUDA results).

        Returns ``trigger [S]`` int32, or ``(trigger, probs [S, max_frames, C], nframes [S])``.
        """
        host = _tensors.is_host(chunk)
        x = _tensors.to_device(chunk, torch.int16, self.device)
        if x.dim() != 2 or x.shape[0] != self.n_streams:
            raise _lib.InvalidArgumentError("chunk must be [%d, n] int16" % self.n_streams)
        C = self.model.config.num_classes
        trig = torch.empty(self.n_streams, dtype=torch.int32, device=self.device)
        probs = nfr = None
        if want_probs:
            probs = torch.zeros((self.n_streams, self.max_frames, C), dtype=torch.float32, device=self.device)
            nfr = torch.empty(self.n_streams, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kws_stream_step(self._handle, _tensors.ptr(x), x.shape[1], x.stride(0), _tensors.ptr(trig),
                                                 _tensors.ptr(probs), _tensors.ptr(nfr), _tensors.stream_ptr(self.device)))
        outs = (trig, probs, nfr) if want_probs else (trig,)
        if host:
            torch.cuda.current_stream(self.device).synchronize()
            outs = tuple(_tensors.to_host(o) for o in outs)
        return outs if want_probs else outs[0]

    def step_host(self, pinned_chunk: torch.Tensor, pinned_trigger: torch.Tensor = None):
        """Enqueue H2D(chunk) -> step -> D2H(trigger) without synchronising.
        ``pinned_chunk`` ``[S, n]`` int16 pinned host tensor; returns the pinned trigger tensor
        (valid after the current stream is synchronised)."""
        if pinned_chunk.dtype != torch.int16 or pinned_chunk.dim() != 2 or pinned_chunk.shape[0] != self.n_streams \
                or not pinned_chunk.is_contiguous() or pinned_chunk.is_cuda:
            raise _lib.InvalidArgumentError("pinned_chunk must be a contiguous host int16 [%d, n] tensor" % self.n_streams)
        if pinned_trigger is None:
            if self._pinned_trigger is None:
                self._pinned_trigger = torch.zeros(self.n_streams, dtype=torch.int32).pin_memory()
            pinned_trigger = self._pinned_trigger
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kws_stream_step_host(self._handle, pinned_chunk.data_ptr(), pinned_chunk.shape[1],
                                                      pinned_trigger.data_ptr(), _tensors.stream_ptr(self.device)))
        return pinned_trigger

    def state(self) -> torch.Tensor:
        """Copy of the carried GRU state ``[layers, S, H]`` (CUDA tensor)."""
        cfg = self.model.config
        out = torch.empty((cfg.num_layers, self.n_streams, cfg.hidden_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kws_stream_copy_state(self._handle, _tensors.ptr(out), _tensors.stream_ptr(self.device)))
        return out

    def window_labels(self, max_labels=None):
        """Window decode of every stream as of the last step -> (labels [S, max_labels], counts [S]) numpy."""
        if max_labels is None:
            max_labels = 2 * self.max_frames * self.model.config.window_chunks + 1
        labels = torch.empty((self.n_streams, max_labels), dtype=torch.int32, device=self.device)
        counts = torch.empty(self.n_streams, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kws_stream_labels(self._handle, _tensors.ptr(labels), max_labels, _tensors.ptr(counts),
                                                   _tensors.stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()
        return _tensors.to_host(labels), _tensors.to_host(counts)
