"""Multi-GPU plumbing of the ORB front-end.

The path shards by independent camera streams: stream s -> rank s mod world_size, one process per GPU, NO
data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) carries exactly two exchanges:
  * a one-off broadcast of the packed vocabulary blob from the rank that parsed/built it, and
  * an all-gather of a small per-stream counter record at report time.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

COUNTER_FIELDS = ("frames", "keypoints_left_last_batch", "keypoints_right_last_batch", "stereo_matches_last_batch", "device_us")


def streams_for_rank(n_streams: int, rank: int, world_size: int) -> List[int]:
    """Camera stream ids owned by `rank` (round robin: stream s lives on GPU s mod world_size)."""
    return [s for s in range(n_streams) if s % world_size == rank]


def broadcast_blob(blob: "torch.Tensor | None", src: int = 0, device: "torch.device | str" = "cpu") -> torch.Tensor:
    """Broadcasts a uint8 blob whose size only `src` knows.  Non-source ranks pass None and get a new tensor."""
    rank = dist.get_rank()
    size = torch.tensor([blob.numel() if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, src=src)
    if rank != src:
        blob = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, src=src)
    return blob


def gather_counters(counters: Sequence[int], device: "torch.device | str" = "cpu") -> np.ndarray:
    """All-gather of one counter record per rank -> (world_size, len(COUNTER_FIELDS)) int64 array on every rank."""
    t = torch.tensor(list(counters), dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


class DeviceBlobView:
    """Exposes memory owned by libborb (e.g. the packed vocabulary) to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def random_vocabulary_arrays(k: int = 10, L: int = 6, seed: int = 7):
    """Seeded random tree of ORBvoc's shape (k=10, L=6 -> 1,111,111 nodes), breadth-first ids; stands in for
    Vocabulary/ORBvoc.txt, which cannot travel to the GPU box."""
    rng = np.random.default_rng(seed)
    counts = [k ** d for d in range(L + 1)]
    offs = np.cumsum([0] + counts)
    n = int(offs[-1])
    parent = np.zeros(n, np.int32)
    for d in range(1, L + 1):
        parent[offs[d]:offs[d + 1]] = offs[d - 1] + np.arange(counts[d]) // k
    is_leaf = np.zeros(n, np.uint8)
    is_leaf[offs[L]:] = 1
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    weight = np.zeros(n, np.float64)
    weight[offs[L]:] = rng.uniform(0.5, 10.0, counts[L])
    return parent, is_leaf, desc, weight
