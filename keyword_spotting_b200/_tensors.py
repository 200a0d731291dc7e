"""PyTorch as tensor glue only: device memory, streams, host<->device copies."""
from __future__ import annotations

import numpy as np
import torch

_NP_TO_TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.int16): torch.int16,
                np.dtype(np.int8): torch.int8, np.dtype(np.int32): torch.int32,
                np.dtype(np.uint8): torch.uint8}


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("keyword_spotting_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
    if dev.type != "cuda":
        raise RuntimeError("keyword_spotting_b200 runs on CUDA devices only, got %r" % (device,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def is_host(x) -> bool:
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def to_device(x, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """numpy / list / torch (any device) -> contiguous CUDA tensor of ``dtype``."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        arr = np.ascontiguousarray(x)
        if arr.dtype not in _NP_TO_TORCH:
            arr = arr.astype(np.float32)
        t = torch.from_numpy(arr)
    return t.to(device=device, dtype=dtype, non_blocking=False).contiguous()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()
