"""Stream sharding across the GPUs of one box.

Streams are independent (state is per stream, no cross-stream op anywhere in
models/rnn_ctc.py:113-166 or detector.py:148-209), so the hot path shards with NO
collective: rank r owns a contiguous block of streams for their whole lifetime
(state, tails and windows live in that GPU's HBM; the 657 KB of weights are
replicated).  The only communication is the gather of per-stream results
(trigger bits / counts) for reporting -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def partition(total_streams: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block ``(start, count)`` of rank ``rank``; blocks differ by at most one stream."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside [0, %d)" % (rank, world_size))
    base, extra = divmod(int(total_streams), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def owner_of(stream_id: int, total_streams: int, world_size: int) -> int:
    """Rank that owns ``stream_id`` under :func:`partition`."""
    base, extra = divmod(int(total_streams), int(world_size))
    boundary = extra * (base + 1)
    if stream_id < boundary:
        return stream_id // (base + 1)
    return extra + (stream_id - boundary) // max(base, 1)


def gather_stream_results(local: torch.Tensor, total_streams: int, group=None) -> torch.Tensor:
    """All-gather per-stream results (``[count, ...]`` on each rank, blocks from :func:`partition`)
    into the global ``[total_streams, ...]`` tensor on every rank.  Off the hot path."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    counts = [partition(total_streams, world, r)[1] for r in range(world)]
    width = max(counts)
    pad_shape = (width,) + tuple(local.shape[1:])
    padded = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def reduce_max_scalar(value: float, device=None, group=None) -> float:
    """MAX over ranks of a device-timed duration (every multi-GPU number is the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def reduce_sum_scalar(value: float, device=None, group=None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin this process to the CPU cores (NUMA node) that are local to GPU `device_index`, BEFORE any pinned host
    buffer is allocated: the host side of the streaming server is one process per GPU that feeds 55 GB/s of PCM
    through page-locked staging buffers, and Linux places those pages on the node the allocating thread runs on.
    With 8 ranks on a two-socket box an unbound rank lands on the far socket half of the time and its H2D copies
    cross the socket interconnect.  Returns what was done (for the bench record); never raises.
    """
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(device_index)       # honours CUDA_VISIBLE_DEVICES remapping
            h = pynvml.nvmlDeviceGetHandleByPciBusId(
                ("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"bound": True, "cpus": len(allowed), "first_cpu": allowed[0], "last_cpu": allowed[-1]}
    except Exception as e:                                     # no NVML / restricted container: run unbound
        info["why"] = "%s: %s" % (type(e).__name__, e)
    return info
