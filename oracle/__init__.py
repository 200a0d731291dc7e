"""CPU oracle for the keyword_spotting hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + plain C) of the reference algorithms
that the CUDA path in ``keyword_spotting_b200`` replaces.  It exists to *check*
the CUDA path; it is never the thing shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``keyword_spotting_b200/`` does.

Pinning status (see DESIGN.md "Oracle"):

* ``octbit``  -- PINNED.  Checked against the reference's two known-answer
  vectors (octbit/octbit_ops_test.py:24-34, :41-53) and against the
  *unmodified* reference kernel octbit/octbit_mat_mul_op.cc compiled from
  /root/reference behind a header shim (``oracle/_ref``).
* ``posenc``  -- PINNED against the unmodified positional_encoding_op.cc
  compiled the same way (the reference's own test asserts nothing).
* ``prediction`` / ``streaming`` (decoders, VAD, queue) -- PINNED against
  golden vectors produced by importing the reference's pure-numpy
  utils/prediction.py, utils/basic_vad.py and utils/queue.py in the build
  container (tests/golden/make_golden.py).
* ``model`` (framing, |rFFT|, mel, TF-GRUCell recurrence, FC, softmax) --
  PARITY UNPINNED.  TensorFlow 1.x and librosa are neither vendored in the
  reference nor installable here, and the reference has no fixture for this
  part.  The restatement follows models/rnn_ctc.py:113-166,202-284 and the
  documented TF 1.x GRUCell semantics, and is cross-checked float64 vs
  float32 and streaming vs offline (detector.py:254-289).
"""
