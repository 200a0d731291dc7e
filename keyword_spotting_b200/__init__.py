"""keyword_spotting_b200 -- B200-native hot path of colinsongf/keyword_spotting.

Importing the package builds (if stale) and loads ``libkws_b200.so`` -- hand-written
CUDA for sm_100a behind the C-ABI of ``include/kws_b200.h`` -- exactly as the
reference compiles its custom ops at import (octbit/op_compile.py).  There is no
CPU fallback anywhere in this package.

Reference entry point                         -> here
  models/rnn_ctc.py       DeployModel          rnn_ctc.DeployModel
  detector.py             HotwordDetector.start streaming.StreamingDetector (one batch), serving.WaveServer (a GPU's streams)
  utils/prediction.py     ctc_decode* / predict utils.prediction
  octbit/octbit_ops.py    octbit_mat_mul       octbit.octbit_ops.octbit_mat_mul
  octbit/octbit_graph.py  octize_weight_int8.. octbit.octbit_graph.octize_weight_int8_signed
  positional_encoding/positional_encoding_op.py positional_encoding.positional_encoding_op
  models/attention_ctc.py DeployModel          attention_ctc.AttentionDeployModel
"""
from . import _lib as _lib_mod

_lib_mod.load()          # fail loudly at import if the CUDA library cannot be built / loaded

from ._lib import InvalidArgumentError, KwsCudaError  # noqa: E402
from .config import Config, get_config  # noqa: E402
from .rnn_ctc import DeployModel, ModelWeights, OctbitMatrix, OctbitModelWeights  # noqa: E402
from .streaming import StreamingDetector  # noqa: E402
from .serving import WaveServer  # noqa: E402
from .attention_ctc import AttentionConfig, AttentionDeployModel, AttentionWeights  # noqa: E402

__all__ = ["Config", "get_config", "DeployModel", "ModelWeights", "OctbitMatrix", "OctbitModelWeights", "StreamingDetector", "WaveServer",
           "AttentionConfig", "AttentionDeployModel", "AttentionWeights",
           "InvalidArgumentError", "KwsCudaError"]
