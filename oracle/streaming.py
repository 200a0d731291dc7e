"""Oracle (TEST INFRASTRUCTURE ONLY): the detector.py streaming loop, batched over streams.

Restates ``HotwordDetector.start`` (detector.py:148-209) for S independent
streams fed in lock-step:  VAD on the new samples (utils/basic_vad.py:17-18,
called at detector.py:168) -> on silence zero the GRU state and clear the
probability window (detector.py:171-177) -> prepend the carried tail and keep
``(len-400)%160+240`` samples for the next chunk (detector.py:179-183) -> model
call (detector.py:190-193) -> push into the 15-chunk window (utils/queue.py:16-38,
detector.py:122,195-197) -> ``ctc_decode2`` over the whole window
(detector.py:200) -> ``ctc_predict(.., '1233')`` and, on trigger, clear the
window and zero the state (detector.py:201-209).

``vad`` / ``SimpleQueue`` are PINNED against the reference's own modules via
tests/golden/make_golden.py; the loop itself is checked against the GPU server.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import model as om
from . import prediction as op

WINDOW_CHUNKS = 15      # detector.py:122  SimpleQueue(15)
VAD_THRESHOLD = 30      # detector.py:168
KEYWORD = "1233"        # detector.py:201, config/rnn_config.py:27


def vad(sig, thres=40):
    """utils/basic_vad.py:17-18, float form exactly as the reference runs it."""
    return bool(np.abs(np.asarray(sig)).sum() > thres)


def vad_pcm16(pcm_i16, thres=VAD_THRESHOLD):
    """The same predicate in exact integer arithmetic on int16 PCM:
    ``sum|x|/32768 > thres``  <=>  ``sum|x| > thres*32768``.  The fp32 pairwise
    sum of the reference can only disagree within rounding distance of the
    threshold (SURVEY.md 7 hard part 5); the GPU server uses this exact form."""
    s = np.abs(np.asarray(pcm_i16).astype(np.int64)).sum(axis=-1)
    return s > int(thres) * 32768


class SimpleQueue:
    """utils/queue.py:16-38 -- bounded FIFO keeping the newest ``maxLen`` items."""

    def __init__(self, maxLen):
        self.maxLen = maxLen
        self.content: List = []

    def clear(self):
        self.content = []

    def add(self, item):
        if len(self.content) == self.maxLen:
            self.content.pop(0)
        self.content.append(item)

    def full(self):
        return len(self.content) == self.maxLen

    def get_all(self):
        return self.content


def residual_length(total_len, fft=om.FFT_SIZE, hop=om.HOP_SIZE):
    """detector.py:181-182 -- samples carried to the next chunk."""
    return (total_len - fft) % hop + (fft - hop)


class StreamOracle:
    """S lock-step streams of the detector loop.  int16 PCM in, exact-integer VAD
    (``vad_rule='int'``) or the reference's float sum (``'float'``)."""

    def __init__(self, weights: om.Weights, n_streams: int, dtype=np.float32,
                 vad_rule="int", vad_thres=VAD_THRESHOLD, window=WINDOW_CHUNKS,
                 label=KEYWORD, decode_thres=0.4, forward=None):
        # forward(pcm_f32 [S, L], state) -> (softmax, state, logits); default: the float deployment graph
        self.forward = forward
        self.w = weights
        self.S = n_streams
        self.dtype = dtype
        self.vad_rule = vad_rule
        self.vad_thres = vad_thres
        self.label = label
        self.decode_thres = decode_thres
        self.state = np.zeros((weights.num_layers, n_streams, weights.hidden), dtype=dtype)
        self.res = np.zeros((n_streams, 0), dtype=np.float32)   # same length for all streams
        self.queues = [SimpleQueue(window) for _ in range(n_streams)]

    def step(self, chunk_i16: np.ndarray, decide_on=None):
        """One chunk ``[S, n_new]`` int16 for every stream.

        Returns dict: ``speech`` bool[S], ``softmax`` f[S, n, C], ``trigger``
        int32[S], ``labels`` list of int32 arrays (window decode per stream),
        ``state`` (after trigger reset).

        ``decide_on``: optional callable ``softmax -> softmax`` applied to what
        enters the decision window (the returned ``softmax`` stays the oracle's
        own).  The parity tests use it to take the frames whose decision lies
        within the float tolerance of a threshold (``ambiguous_frames``) from the
        implementation under test, so that everything else must match bit for bit.
        """
        chunk_i16 = np.asarray(chunk_i16, dtype=np.int16)
        assert chunk_i16.shape[0] == self.S
        data = om.pcm16_to_float(chunk_i16)
        if self.vad_rule == "int":
            speech = vad_pcm16(chunk_i16, self.vad_thres)
        else:
            speech = np.array([vad(d, self.vad_thres) for d in data])
        for s in np.nonzero(~speech)[0]:
            self.state[:, s, :] = 0
            self.queues[s].clear()
        full = np.concatenate([self.res, data], axis=1)
        keep = residual_length(full.shape[1])
        self.res = full[:, -keep:]
        if self.forward is not None:
            softmax, state, _ = self.forward(full, self.state)
        else:
            softmax, state, _ = om.deploy_forward(full, self.state, self.w, self.dtype)
        self.state = state
        trigger = np.zeros(self.S, dtype=np.int32)
        labels = []
        decided = softmax if decide_on is None else decide_on(softmax)
        for s in range(self.S):
            q = self.queues[s]
            q.add(decided[s])
            window = np.concatenate(q.get_all(), axis=0)
            seq = op.ctc_decode2(window, self.w.num_classes, self.decode_thres)
            labels.append(seq)
            if op.ctc_predict(seq, self.label):
                trigger[s] = 1
                q.clear()
                self.state[:, s, :] = 0
        return dict(speech=speech, softmax=softmax, trigger=trigger, labels=labels,
                    state=self.state.copy())


def ambiguous_frames(softmax: np.ndarray, thres: float, tol: float) -> np.ndarray:
    """bool ``[S, n]``: frames whose ``ctc_decode2`` token (utils/prediction.py:71-84: above-threshold winner of
    columns ``1..C-2``, else none) can differ between two softmax arrays that agree within ``tol`` (max-abs):
    the winner is within ``tol`` of the threshold, or it is above ``thres - tol`` and the runner-up is within
    ``2 tol`` of it.  For every other frame a ``tol``-close softmax yields the identical token."""
    lab = np.asarray(softmax, dtype=np.float64)[..., 1:-1]
    top = np.sort(lab, axis=-1)
    m1 = top[..., -1]
    m2 = top[..., -2] if lab.shape[-1] > 1 else np.full_like(m1, -np.inf)
    return (np.abs(m1 - thres) <= tol) | ((m1 > thres - tol) & (m1 - m2 <= 2 * tol))


def stream_vs_offline(pcm_f32: np.ndarray, weights: om.Weights, seg_len=3600, dtype=np.float32):
    """detector.py:254-289 ``test2``: feed one utterance in ``seg_len`` pieces
    with carried tail and state; returns the concatenated softmax, which must
    equal the whole-utterance softmax (the reference's only self-check)."""
    pcm_f32 = np.asarray(pcm_f32, dtype=np.float32)
    state = np.zeros((weights.num_layers, 1, weights.hidden), dtype=dtype)
    res = np.zeros(0, dtype=np.float32)
    outs = []
    seg_num = len(pcm_f32) // seg_len
    for i in range(seg_num + 1):
        feed = pcm_f32[i * seg_len:] if i == seg_num else pcm_f32[i * seg_len:(i + 1) * seg_len]
        data = np.concatenate([res, feed])
        if len(data) < om.FFT_SIZE:
            res = data
            continue
        res = data[-residual_length(len(data)):]
        used = len(data) - (len(data) - om.FFT_SIZE) % om.HOP_SIZE
        sm, state, _ = om.deploy_forward(data[:used], state, weights, dtype)
        outs.append(sm[0])
    return np.concatenate(outs, axis=0), state
