"""Multi-GPU plumbing: one process per GPU, pairs sharded by contiguous block, one NCCL broadcast of
the weight blob at start-up and no collective on the data path.

The reference is single-process, one pair at a time (pose_node.py:191-497); independent
(query-frame, map-tile) pairs are the natural unit to shard.  ``torch.distributed`` is plumbing
only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n_items`` owned by ``rank``; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frame_owner(frame_index: int, world: int) -> int:
    """Stream config (BASELINE config 4): frame i -> GPU i mod G; every GPU evaluates all
    candidate tiles of its frames, so no descriptors cross GPUs."""
    return frame_index % world


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def broadcast_weights(blob: Optional[bytes], nbytes: int, device=None, src: int = 0):
    """Rank ``src`` passes the blob, the others ``None``; returns a uint8 tensor (on ``device``)
    holding the blob on every rank.  With the NCCL backend this is a single ncclBroadcast over
    NVLink; with gloo (tests) it runs on the CPU."""
    import torch
    import torch.distributed as dist

    dev = device if device is not None else "cpu"
    if blob is not None:
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def gather_counts(local: np.ndarray):
    """Sum small per-rank counters on rank 0 (results only, never image data)."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(np.asarray(local, np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
