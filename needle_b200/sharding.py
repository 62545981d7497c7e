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


# ----------------------------------------------------------------------------------------------------------
# One haystack split across ranks (BASELINE config 4 on several GPUs; SURVEY.md section 8(e), second row).
#
# Rank r holds the contiguous chunk [base_r, base_r + n_r).  The forward scan of a chunk needs the automaton state
# at its first char, which depends on everything before it.  The protocol needs one small all-gather per round:
#
#   forward  round 0: every rank r > 0 GUESSES its entry state (the scan, from the root, of the `HALO` chars before
#            its chunk - the tail of rank r-1, exchanged once), rank 0 starts in the root; all ranks scan at once.
#            The (entry, end, exit) records are all-gathered and every rank resolves them the same way, walking
#            the ranks in order from the root: a record is VERIFIED when its entry equals the exit of the verified
#            record before it.  The walk stops at the first unverified record - that rank scans again from the now
#            known state and another round runs - or when the automaton has died / the ranks are exhausted: the
#            end of the match is the last accepting index seen on the verified path (indexForwards,
#            DFAClassBuilder.java:335-471).  A search automaton forgets (long8.cuh), so round 0 normally suffices;
#            in the worst case (`a.*c`) every round verifies one more rank, like a sequential hand-over.
#   reverse  fixed-length patterns: start = end - minLength (:640-646).  Otherwise indexBackwards(end - 1, 0)
#            (:529-614) starts on the rank that holds end - 1 and is handed to the rank before it for as long as
#            the BACKWARDS automaton is still alive at a chunk boundary (one all-gather per hand-over).
#
# `scan(entry_state) -> (end_local | -1, exit_state)` and `scan_back(index_local, entry_state, last_init_local) ->
# (start_local | NO_START, exit_state)` are the per-rank primitives: Pattern.find_long_from / find_long_back on a
# GPU, the oracle's scan_from / scan_back_from in the CPU tests.
# ----------------------------------------------------------------------------------------------------------
HALO = 16
NO_START = 2 ** 63 - 1


def resolve_forward(records, dead_state):
    """Pure step of the forward protocol.  records[r] = (entry, end_local, exit, base).  Returns
    ("done", end_global | -1) or ("rescan", rank, entry_state)."""
    state, end = 0, -1
    for r, (entry, end_local, exit_state, base) in enumerate(records):
        if state == dead_state:
            break
        if entry != state:
            return "rescan", r, state
        if end_local != -1:
            end = base + end_local
        state = exit_state
    return "done", end


def find_long_sharded(scan, scan_back, guess_entry, base, n_local, rank, world_size, allgather, fwd_dead, bwd_dead,
                      reverse_mode, min_length, bwd_root_accepting):
    """find() of one haystack whose chunk [base, base + n_local) lives on this rank.  Every rank returns the same
    (matched, start, end) in global indices.  `allgather(obj)` returns the list of every rank's obj."""
    entry = 0 if rank == 0 else guess_entry()
    rec = None
    end = -1
    for _ in range(world_size + 1):
        if rec is None or rec[0] != entry:
            # (rank 0 always scans: an accepting root matches the empty haystack, DFAClassBuilder.java:356)
            end_local, exit_state = scan(entry) if (n_local > 0 or rank == 0) else (-1, entry)
            rec = (int(entry), int(end_local), int(exit_state), int(base))
        verdict = resolve_forward(allgather(rec), fwd_dead)
        if verdict[0] == "done":
            end = verdict[1]
            break
        if verdict[1] == rank:
            entry = verdict[2]
    else:
        raise RuntimeError("find_long_sharded: forward phase did not converge")
    if end == -1:
        return False, -1, -1
    if reverse_mode == 2:
        return True, end - min_length, end

    # reverse pass: (state, smallest accepting index so far, next index to read), handed down rank by rank
    state, start, index = 0, (0 if (bwd_root_accepting and reverse_mode == 0) else NO_START), end - 1
    bases = allgather((int(base), int(n_local)))
    for r in range(world_size - 1, -1, -1):
        b, n = bases[r]
        if n == 0 or index < b or index >= b + n:
            continue  # this rank holds nothing of [.., index]
        if rank == r:
            li = start - b if start != NO_START else NO_START
            s_local, state = scan_back(index - b, state, li)
            start = s_local + b if s_local != NO_START else NO_START
        state, start = allgather((int(state), int(start)) if rank == r else None)[r]
        index = b - 1
        if state == bwd_dead or index < 0:
            break
    return True, start, end


def tensor_allgather(device=None):
    """An `allgather(obj)` for find_long_sharded over torch.distributed: obj is None or a tuple of up to 4 ints,
    moved as one int64[5] all-gather (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                              if dist.get_backend() == "nccl" else torch.device("cpu"))

    # one small buffer of each kind, reused by every call: pinned host memory on GPUs, so that the two copies are asynchronous DMAs
    cuda = dev.type == "cuda"
    mine_h = torch.zeros(5, dtype=torch.int64)
    out_h = torch.zeros(world * 5, dtype=torch.int64)
    if cuda:
        mine_h, out_h = mine_h.pin_memory(), out_h.pin_memory()
    mine_d = torch.zeros(5, dtype=torch.int64, device=dev) if cuda else mine_h
    out_d = torch.zeros(world * 5, dtype=torch.int64, device=dev) if cuda else out_h
    mine_np, out_np = mine_h.numpy(), out_h.numpy()

    def allgather(obj):
        mine_np[:] = 0
        if obj is not None:
            mine_np[0] = len(obj)
            mine_np[1:1 + len(obj)] = [int(v) for v in obj]
        if cuda:
            mine_d.copy_(mine_h, non_blocking=True)
        dist.all_gather_into_tensor(out_d, mine_d)
        if cuda:
            out_h.copy_(out_d, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        rows = out_np.reshape(world, 5).tolist()
        return [tuple(r[1:1 + r[0]]) if r[0] else None for r in rows]
    return allgather


def exchange_halo(tail, device=None):
    """Every rank contributes the last HALO bytes of its chunk (a uint8 tensor of exactly HALO entries, zero padded
    at the front if the chunk is shorter); returns the tail of the rank before this one as a numpy array."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    out = torch.empty(world * HALO, dtype=torch.uint8, device=tail.device)
    dist.all_gather_into_tensor(out, tail.contiguous())
    prev = out.view(world, HALO)[rank - 1] if rank > 0 else out.view(world, HALO)[0][:0]
    return prev.cpu().numpy()


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
