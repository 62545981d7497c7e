"""Multi-GPU plumbing for batched matching: one process per GPU (torch.distributed), the table blob is
broadcast once from rank 0, haystack batches are split into contiguous line ranges balanced by bytes, and
there is NO collective on the data path (every haystack is independent: the reference's Matcher instances
share nothing, DFAClassBuilder.java:669-694).  Works with the `nccl` backend on GPUs and with `gloo` on CPU
(the CPU tests run world_size 2 over gloo).
"""
from typing import List, Optional, Tuple

import numpy as np


def shard_ranges(offsets: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Split lines [0, n) into `world_size` contiguous ranges with (nearly) equal numbers of chars.
    `offsets` is the n+1 prefix array of the batch.  Returns [(lo, hi)] per rank; ranges tile [0, n)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if n <= 0:
        return [(0, 0)] * world_size
    base, total = int(offsets[0]), int(offsets[-1] - offsets[0])
    cuts = [0]
    for k in range(1, world_size):
        target = base + (total * k) // world_size
        idx = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        cuts.append(min(max(idx, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world_size)]


def shard_batch(data: np.ndarray, offsets: np.ndarray, char_width: int, rank: int, world_size: int):
    """This rank's slice of a batch: (data_view, offsets_rebased, (lo, hi))."""
    lo, hi = shard_ranges(offsets, world_size)[rank]
    o = np.asarray(offsets[lo:hi + 1], dtype=np.uint64)
    if hi > lo:
        b0, b1 = int(o[0]) * char_width, int(o[-1]) * char_width
        return np.asarray(data).view(np.uint8)[b0:b1], o - o[0], (lo, hi)
    return np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64), (lo, hi)


def broadcast_blob(blob: Optional[bytes], src: int = 0, device=None) -> bytes:
    """One broadcast of the compiled pattern's table blob from `src` to every rank (NCCL over NVLink on
    GPUs).  `blob` is only read on `src`.  Needs an initialised torch.distributed process group."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                              if dist.get_backend() == "nccl" else torch.device("cpu"))
    size = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(size, src)
    buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    if rank == src:
        buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def gather_results(matched: np.ndarray, start: Optional[np.ndarray], end: Optional[np.ndarray], dst: int = 0):
    """Optional: collect per-rank result slices on `dst` in rank order (not on the timed path)."""
    import torch.distributed as dist

    world = dist.get_world_size()
    parts = [None] * world if dist.get_rank() == dst else None
    dist.gather_object((matched, start, end), parts, dst=dst)
    if parts is None:
        return None
    cat = lambda k: None if parts[0][k] is None else np.concatenate([p[k] for p in parts])  # noqa: E731
    return cat(0), cat(1), cat(2)
