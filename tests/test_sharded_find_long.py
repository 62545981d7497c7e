"""One haystack split across ranks ("single 8 GiB haystack"): needle_b200.sharding.find_long_sharded.

CPU: the protocol is run by `world` threads (a barrier-based all-gather) and by a world_size-2 gloo group, with the
oracle's scan_from / scan_back_from as the per-rank primitives, against the oracle's plain sequential find().
GPU (-m gpu): the same protocol with Pattern.find_long_from / find_long_back on one device standing in for every rank."""
import ctypes
import os
import socket
import threading
import zlib

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200.blob import parse_blob
from needle_b200.sharding import HALO, NO_START, find_long_sharded, resolve_forward
from tests.oracle_lib import INT64_MAX, Oracle, lib as oracle_lib


def oracle_find_long(ora, data):
    st, en = ctypes.c_int64(), ctypes.c_int64()
    data = np.ascontiguousarray(data).view(np.uint8)
    m = oracle_lib().ndlo_find(ora._h, data.ctypes.data if data.size else None, data.size, 1, 0, INT64_MAX, ctypes.byref(st), ctypes.byref(en))
    return bool(m), st.value, en.value


class ThreadGather:
    """all-gather between `world` threads."""

    def __init__(self, world):
        self.world, self.slots, self.bar = world, [None] * world, threading.Barrier(world)

    def make(self, rank):
        def allgather(obj):
            self.slots[rank] = obj
            self.bar.wait()
            out = list(self.slots)
            self.bar.wait()
            return out
        return allgather


def cuts_for(n, world, rng=None):
    if rng is None:
        return [n * k // world for k in range(world + 1)]
    inner = sorted(int(x) for x in rng.integers(0, n + 1, size=world - 1))
    return [0] + inner + [n]


def run_ranks(blob, data, cuts, make_prims):
    """Run the protocol on len(cuts)-1 simulated ranks; returns the list of per-rank results."""
    world = len(cuts) - 1
    info = parse_blob(blob)
    tg = ThreadGather(world)
    results, errors = [None] * world, []

    def body(rank):
        try:
            lo, hi = cuts[rank], cuts[rank + 1]
            scan, scan_back, guess, fd, bd, bra = make_prims(rank, data, lo, hi)
            results[rank] = find_long_sharded(scan, scan_back, guess, lo, hi - lo, rank, world, tg.make(rank), fd, bd,
                                              info.reverse_mode, info.min_length, bra)
        except BaseException as e:  # noqa: BLE001 - surface it in the main thread
            errors.append(e)
            tg.bar.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def oracle_prims(ora):
    def make(rank, data, lo, hi):
        chunk = data[lo:hi]

        def scan(entry):
            return ora.scan_from(chunk, entry)

        def scan_back(index, entry, last_init):
            return ora.scan_back_from(chunk, index, entry, last_init)

        def guess():
            return ora.scan_from(data[max(0, lo - HALO):lo], 0)[1]
        return scan, scan_back, guess, ora.forwards_state_count(), ora.backwards_state_count(), ora.backwards_root_accepting()
    return make


CASES = [
    ("a[ab]{7}c", b"ab", [b"abababbac"]),            # BASELINE C4: fixed length, forgets within 9 chars
    ("[a-z0-9._%+-]+@[a-z0-9.-]+", b"ab .,", [b"x@y", b"abc.def@host.example"]),  # table-driven reverse pass
    ("a.*c", b"ab\n", [b"c"]),                        # remembers arbitrarily far back: guesses are wrong, matches span ranks
    ("a+b", b"ab ", []),                              # single-char reverse scan candidates
    ("(ab)*", b"abc", []),                            # accepting root
    ("x[ab]*y|abba", b"abxy", []),
]


@pytest.mark.parametrize("regex,alphabet,plants", CASES)
def test_protocol_matches_sequential_find(regex, alphabet, plants):
    blob = nb.compile_to_bytes(regex, 0)
    ora = Oracle(blob)
    rng = np.random.default_rng(zlib.crc32(regex.encode()))
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    n_hit = 0
    for trial in range(40):
        n = int(rng.integers(0, 400))
        data = alpha[rng.integers(0, len(alpha), size=n)].copy()
        for pl in plants:
            if n > len(pl) and rng.random() < 0.5:
                pos = int(rng.integers(0, n - len(pl)))
                data[pos:pos + len(pl)] = np.frombuffer(pl, dtype=np.uint8)
        want = oracle_find_long(ora, data)
        n_hit += want[0]
        for world in (1, 2, 3, 5):
            for cuts in (cuts_for(n, world), cuts_for(n, world, rng)):
                got = run_ranks(blob, data, cuts, oracle_prims(ora))
                assert all(g == want for g in got), (regex, trial, cuts, got, want)
    assert n_hit > 0 or not plants


def test_match_spanning_every_rank():
    blob = nb.compile_to_bytes("a.*c", 0)
    ora = Oracle(blob)
    data = np.full(1000, ord("b"), dtype=np.uint8)
    data[3], data[990] = ord("a"), ord("c")
    want = oracle_find_long(ora, data)
    assert want == (True, 3, 991)
    assert run_ranks(blob, data, cuts_for(1000, 4), oracle_prims(ora)) == [want] * 4
    # empty chunks in the middle and at the ends
    assert run_ranks(blob, data, [0, 0, 500, 500, 1000, 1000], oracle_prims(ora)) == [want] * 5


def test_empty_haystack_and_accepting_root():
    for regex in ("(ab)*", "a[ab]{7}c"):
        blob = nb.compile_to_bytes(regex, 0)
        ora = Oracle(blob)
        for data in (np.zeros(0, dtype=np.uint8), np.frombuffer(b"ab", dtype=np.uint8)):
            want = oracle_find_long(ora, data)
            for cuts in ([0, 0, len(data)], [0, len(data), len(data)], [0, 0, 0, len(data)]):
                assert run_ranks(blob, data, cuts, oracle_prims(ora)) == [want] * (len(cuts) - 1), (regex, cuts)


def test_resolve_forward_unit():
    dead = 9
    assert resolve_forward([(0, -1, 3, 0), (3, 5, dead, 100), (7, 1, 2, 200)], dead) == ("done", 105)
    assert resolve_forward([(0, -1, 3, 0), (4, 5, dead, 100)], dead) == ("rescan", 1, 3)
    assert resolve_forward([(0, -1, 3, 0), (3, -1, 4, 100)], dead) == ("done", -1)
    assert resolve_forward([(0, 7, 3, 0), (3, -1, 4, 100)], dead) == ("done", 7)


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        ok = 0
        for regex, alphabet, plant in (("a[ab]{7}c", b"ab", b"abbbbbbbc"), ("[a-z]+@[a-z]+", b"ab ", b"aa@bb")):
            blob = nb.compile_to_bytes(regex, 0)
            ora, info = Oracle(blob), parse_blob(blob)
            rng = np.random.default_rng(7)  # same data on both ranks
            alpha = np.frombuffer(alphabet, dtype=np.uint8)
            for pos in (10, 4990, 4996, 5000, 9000):
                data = alpha[rng.integers(0, len(alpha), size=10_000)].copy()
                if regex.startswith("a["):
                    data[data == ord("c")] = ord("b")
                data[pos:pos + len(plant)] = np.frombuffer(plant, dtype=np.uint8)
                lo, hi = (0, 5000) if rank == 0 else (5000, 10_000)
                scan, scan_back, guess, fd, bd, bra = oracle_prims(ora)(rank, data, lo, hi)
                got = find_long_sharded(scan, scan_back, guess, lo, hi - lo, rank, world, allgather, fd, bd, info.reverse_mode,
                                        info.min_length, bra)
                assert got == oracle_find_long(ora, data), (regex, pos, got)
                ok += 1
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write(str(ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_single_haystack(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == open(tmp_path / "ok1").read() == "10"


# ------------------------------------------------------------------------------------------------ GPU
def gpu_prims(pat):
    import torch

    def make(rank, data, lo, hi):
        chunk = torch.from_numpy(np.ascontiguousarray(data[lo:hi])).cuda()
        halo = np.ascontiguousarray(data[max(0, lo - HALO):lo])

        def scan(entry):
            return pat.find_long_from(chunk.data_ptr(), hi - lo, entry, mem_kind=nb.MEM_DEVICE)

        def scan_back(index, entry, last_init):
            return pat.find_long_back(chunk.data_ptr(), hi - lo, index, entry, last_init, mem_kind=nb.MEM_DEVICE)

        def guess():
            return pat.find_long_from(halo.ctypes.data if halo.size else 0, halo.size, 0, mem_kind=nb.MEM_HOST)[1]
        return scan, scan_back, guess, pat.forwards_state_count, pat.backwards_state_count, pat.backwards_root_accepting
    return make


@pytest.mark.gpu
@pytest.mark.parametrize("regex,alphabet,plants", CASES)
def test_gpu_chunks_match_sequential_find(regex, alphabet, plants):
    blob = nb.compile_to_bytes(regex, 0)
    pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    for n in (0, 37, 5000, 300_000):
        data = alpha[rng.integers(0, len(alpha), size=n)].copy()
        if regex.startswith("a[ab]"):
            data[data == ord("c")] = ord("b")
        for pl in plants:
            if n - len(pl) > n // 2:
                pos = int(rng.integers(n // 2, n - len(pl)))
                data[pos:pos + len(pl)] = np.frombuffer(pl, dtype=np.uint8)
        want = oracle_find_long(ora, data)
        for world in (1, 2, 4):
            for cuts in (cuts_for(n, world), cuts_for(n, world, rng)):
                got = run_ranks(blob, data, cuts, gpu_prims(pat))
                assert all(g == want for g in got), (regex, n, cuts, got, want)


@pytest.mark.gpu
def test_gpu_c4_match_at_the_end_of_the_last_rank():
    blob = nb.compile_to_bytes("a[ab]{7}c", 0)
    pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
    n = 8_000_000
    rng = np.random.default_rng(3)
    data = (rng.integers(0, 2, size=n, dtype=np.uint8) + ord("a")).astype(np.uint8)
    data[n - 9], data[n - 1] = ord("a"), ord("c")
    got = run_ranks(blob, data, cuts_for(n, 4), gpu_prims(pat))
    assert got == [(True, n - 9, n)] * 4
    assert NO_START == INT64_MAX
