"""The N > 1 path on CPU: world_size-2 `gloo` process group.  Rank 0 compiles the regex, the table blob is
broadcast once, each rank takes its byte-balanced contiguous range of lines, and the concatenated
per-rank results equal the single-process result.  The matcher on each rank is the CPU oracle here (there
is no GPU in this container); on the GPU box bench.py runs the same plumbing over NCCL with the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from needle_b200.sharding import shard_batch, shard_ranges  # noqa: E402
from tests import workloads  # noqa: E402


def test_shard_ranges_tile_and_balance():
    data, offsets = workloads.c3_lines(20_000)
    for world in (1, 2, 3, 4, 8):
        r = shard_ranges(offsets, world)
        assert r[0][0] == 0 and r[-1][1] == len(offsets) - 1
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        chars = [int(offsets[hi] - offsets[lo]) for lo, hi in r]
        assert max(chars) - min(chars) <= 2 * 120  # within a couple of lines of each other


def test_shard_ranges_degenerate():
    assert shard_ranges(np.zeros(1, dtype=np.uint64), 4) == [(0, 0)] * 4
    one = np.array([0, 10], dtype=np.uint64)
    r = shard_ranges(one, 4)
    assert sum(hi - lo for lo, hi in r) == 1
    empty_lines = np.zeros(11, dtype=np.uint64)
    r = shard_ranges(empty_lines, 2)
    assert r[0][0] == 0 and r[-1][1] == 10 and r[0][1] == r[1][0]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import needle_b200 as nb
    from needle_b200.sharding import broadcast_blob, gather_results
    from tests.oracle_lib import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        regex = workloads.REGEX["c3"]
        blob = nb.compile_to_bytes(regex, 0) if rank == 0 else None
        blob = broadcast_blob(blob, src=0)
        data, offsets = workloads.c3_lines(30_000)  # same seed on every rank = the same global batch
        d, o, (lo, hi) = shard_batch(data, offsets, 1, rank, world)
        m, s, e = Oracle(blob).match_batch(2, d, o, 1)
        assert len(m) == hi - lo
        res = gather_results(m, s, e, dst=0)
        if rank == 0:
            em, es, ee = Oracle(blob).match_batch(2, data, offsets, 1)
            assert np.array_equal(res[0], em) and np.array_equal(res[1], es) and np.array_equal(res[2], ee)
            with open(os.path.join(out_dir, "ok"), "w") as f:
                f.write(f"{int(em.sum())}")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_batch(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
    assert int(open(tmp_path / "ok").read()) > 5000
