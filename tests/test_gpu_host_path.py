"""GPU: the NDL_MEM_HOST path of the C ABI - pipelined chunks, the pinned bounce ring for pageable callers, library-owned
pinned buffers (ndl_host_alloc), offsets validation, and multi-device patterns (device = -1: one NCCL broadcast of the
table arena, batches sharded across the GPUs inside the library).  Results bit-exact against the CPU oracle."""
import ctypes

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200 import _lib
from tests import workloads
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu

SSN = workloads.REGEX["c2"]


def check(pat, ora, data, offsets, cw=1, modes=(0, 1, 2)):
    for mode in modes:
        got = pat.match_batch(mode, data, offsets, cw)
        exp = ora.match_batch(mode, data, offsets, cw, threads=8)
        for g, e in zip(got, exp):
            if e is not None:
                assert np.array_equal(g, e), f"mode {mode}"


def test_pageable_batch_larger_than_the_ring_and_several_chunks():
    """numpy arrays are pageable: 200 MB of lines go through the 3 x 32 MB pinned ring and 4 pipeline chunks."""
    blob = nb.compile_to_bytes(SSN, 0)
    pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
    data, offsets = workloads.c2_lines(3_200_000)
    check(pat, ora, data, offsets, modes=(2,))
    # ragged: offsets travel too
    data, offsets = workloads.c3_lines(1_500_000)
    blob3 = nb.compile_to_bytes(workloads.REGEX["c3"], 0)
    check(nb.Pattern(blob3, device=0), Oracle(blob3), data, offsets, modes=(2, 1))


def test_library_pinned_buffers():
    L = _lib.lib()
    blob = nb.compile_to_bytes(SSN, 0)
    pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
    data, offsets = workloads.c2_lines(500_000)
    n = len(offsets) - 1
    bufs = [L.ndl_host_alloc(sz) for sz in (data.nbytes, offsets.nbytes, n, 4 * n, 4 * n)]
    assert all(bufs)
    try:
        ctypes.memmove(bufs[0], data.ctypes.data, data.nbytes)
        ctypes.memmove(bufs[1], offsets.ctypes.data, offsets.nbytes)
        pat.match_batch_ptrs(2, bufs[0], bufs[1], n, 1, bufs[2], bufs[3], bufs[4], mem_kind=nb.MEM_HOST)
        m = np.ctypeslib.as_array(ctypes.cast(bufs[2], ctypes.POINTER(ctypes.c_uint8)), (n,))
        s = np.ctypeslib.as_array(ctypes.cast(bufs[3], ctypes.POINTER(ctypes.c_int32)), (n,))
        e = np.ctypeslib.as_array(ctypes.cast(bufs[4], ctypes.POINTER(ctypes.c_int32)), (n,))
        em, es, ee = ora.match_batch(2, data, offsets, 1, threads=8)
        assert np.array_equal(m, em) and np.array_equal(s, es) and np.array_equal(e, ee)
    finally:
        for b in bufs:
            L.ndl_host_free(b)


@pytest.mark.parametrize("lens", [[100] * 999 + [70 << 20], [100, 100, 70 << 20], [70 << 20] + [100] * 999])
def test_skewed_batches(lens):
    """One line that dwarfs a pipeline chunk, first or last: the chunk splitter used to run past the offsets array."""
    blob = nb.compile_to_bytes(SSN, 0)
    pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
    rng = np.random.default_rng(11)
    offsets = np.concatenate([[0], np.cumsum(np.array(lens, dtype=np.uint64))]).astype(np.uint64)
    alpha = np.frombuffer(b"0123456789abc -", dtype=np.uint8)
    data = alpha[rng.integers(0, len(alpha), size=int(offsets[-1]), dtype=np.uint8)]
    for i in range(0, len(lens), 7):  # some planted matches, one of them deep inside the long line
        o = int(offsets[i]) + max(0, min(int(lens[i]) - 11, int(lens[i]) // 2))
        data[o:o + 11] = np.frombuffer(b"123-45-6789", dtype=np.uint8)
    check(pat, ora, data, offsets, modes=(2,))


def test_decreasing_offsets_are_refused():
    pat = nb.Pattern(nb.compile_to_bytes("abc", 0), device=0)
    data = np.frombuffer(b"xxabcxxabcxxxxxxabc", dtype=np.uint8)
    for bad in ([0, 9, 5, 19], [5, 3, 19, 19]):
        with pytest.raises(RuntimeError, match="non-decreasing"):
            pat.match_batch(2, data, np.array(bad, dtype=np.uint64))


def test_current_device_is_restored():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    blob = nb.compile_to_bytes("abc", 0)
    pat = nb.Pattern(blob, device=1)
    pat.match_batch(2, np.frombuffer(b"xxabcxx", dtype=np.uint8), np.array([0, 7], dtype=np.uint64))
    assert torch.cuda.current_device() == 0


def test_multi_device_pattern_shards_the_batch():
    """device = -1: one replica per visible GPU (on a 1-GPU box: one), table arena NCCL-broadcast, same results."""
    L = _lib.lib()
    blob = nb.compile_to_bytes(SSN, 0)
    pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
    assert L.ndl_pattern_device_count(pat._h) == L.ndl_device_count()
    data, offsets = workloads.c2_lines(1_000_000)
    check(pat, ora, data, offsets)
    m, s, e = pat.match_lines(2, data, len(offsets) - 1, 64)
    em, es, ee = ora.match_batch(2, data, offsets, 1, threads=8)
    assert np.array_equal(m, em) and np.array_equal(s, es) and np.array_equal(e, ee)
    d3, o3 = workloads.c3_lines(300_000)
    blob3 = nb.compile_to_bytes(workloads.REGEX["c3"], 0)
    check(nb.Pattern(blob3, device=-1), Oracle(blob3), d3, o3)
    if L.ndl_device_count() > 1:  # device memory belongs to one GPU
        import torch
        t = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
        with pytest.raises(RuntimeError):
            pat.match_batch_ptrs(2, t.data_ptr(), t.data_ptr(), 1, 1, t.data_ptr(), t.data_ptr(), t.data_ptr())


def _oracle_find_long(ora, data, from_=0, cw=1):
    import ctypes
    from tests.oracle_lib import INT64_MAX, lib as oracle_lib
    st, en = ctypes.c_int64(), ctypes.c_int64()
    data = np.ascontiguousarray(data).view(np.uint8)
    m = oracle_lib().ndlo_find(ora._h, data.ctypes.data, data.size // cw, cw, from_, INT64_MAX, ctypes.byref(st), ctypes.byref(en))
    return bool(m), st.value, en.value


@pytest.fixture
def three_replicas():
    """ndl_pattern_create(device = -1) builds three replicas on GPU 0: the sharding code of the multi-device paths runs on a one-GPU box."""
    L = _lib.lib()
    L.ndl_debug_force_replicas(3)
    try:
        yield 3
    finally:
        L.ndl_debug_force_replicas(0)


def test_replicas_shard_a_batch(three_replicas):
    L = _lib.lib()
    blob = nb.compile_to_bytes(SSN, 0)
    pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
    assert L.ndl_pattern_device_count(pat._h) == 3
    data, offsets = workloads.c2_lines(300_000)
    check(pat, ora, data, offsets)
    d3, o3 = workloads.c3_lines(200_000)
    blob3 = nb.compile_to_bytes(workloads.REGEX["c3"], 0)
    check(nb.Pattern(blob3, device=-1), Oracle(blob3), d3, o3)


def test_replicas_split_one_long_haystack(three_replicas):
    """ndl_find_long on a multi-device pattern: chunk per replica, guessed entry states, in-order resolution with re-scans, reverse pass
    handed down across chunk boundaries - against the oracle's sequential walk."""
    n = 12 << 20
    rng = np.random.default_rng(11)
    ab = (rng.integers(0, 2, size=n, dtype=np.uint8) + ord("a")).astype(np.uint8)
    cut1, cut2 = (n // 3) & ~255, (2 * (n // 3)) & ~255  # the chunk boundaries of find_long_multi for from = 0

    # fixed-length pattern (BASELINE config 4): no match, match at the end, match across a boundary, from > 0
    blob = nb.compile_to_bytes(workloads.REGEX["c4"], 0)
    pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
    assert pat.find_long(ab) == _oracle_find_long(ora, ab) == (False, -1, -1)
    for pos in (n - 9, cut1 - 4, cut2 - 8, cut2, 5, cut1 + 100_000):
        data = ab.copy()
        data[pos], data[pos + 8] = ord("a"), ord("c")
        assert pat.find_long(data) == _oracle_find_long(ora, data) == (True, pos, pos + 9), pos
    data = ab.copy()
    for pos in (1000, cut2 + 77):
        data[pos], data[pos + 8] = ord("a"), ord("c")
    assert pat.find_long(data, from_=1001) == _oracle_find_long(ora, data, 1001) == (True, cut2 + 77, cut2 + 86)
    assert pat.find_long(data, from_=n - 5) == _oracle_find_long(ora, data, n - 5)

    # variable-length patterns: reverse pass on the tables, matches that straddle a boundary, automata that remember across it
    text = np.frombuffer(bytes(rng.choice(list(b"abcdefghij klmnop"), size=n).astype(np.uint8)), dtype=np.uint8).copy()
    for regex in (r"q[a-z ]*7", r"[0-9]+x", r"Sherlock|Street"):
        blob = nb.compile_to_bytes(regex, 0)
        pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
        assert pat.find_long(text) == _oracle_find_long(ora, text) == (False, -1, -1), regex
    blob = nb.compile_to_bytes(r"q[a-z ]*7", 0)
    pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
    for q, seven in ((cut1 - 50, cut1 + 50), (cut1 - 3_000_000, cut2 + 10), (100, n - 1), (cut2 - 1, cut2)):
        data = text.copy()
        data[q], data[seven] = ord("q"), ord("7")
        assert pat.find_long(data) == _oracle_find_long(ora, data) == (True, q, seven + 1), (q, seven)
    blob = nb.compile_to_bytes(r"[0-9]+x", 0)
    pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
    data = text.copy()
    data[cut1 - 20:cut1 + 30] = ord("5")
    data[cut1 + 30] = ord("x")
    assert pat.find_long(data) == _oracle_find_long(ora, data) == (True, cut1 - 20, cut1 + 31)

    # UTF-16 haystack, pageable memory (the pinned ring stages it)
    wide = text[: 4 << 20].astype(np.uint16)
    blob = nb.compile_to_bytes(r"q[a-z ]*7", 0)
    pat, ora = nb.Pattern(blob, device=-1), Oracle(blob)
    wide[(4 << 20) // 3 - 10], wide[(4 << 20) // 3 + 500] = ord("q"), ord("7")
    assert pat.find_long(wide, char_width=2) == _oracle_find_long(ora, wide, cw=2)


def test_concurrent_callers_on_shared_and_separate_patterns():
    """Host threads calling into the library at once: the same pattern from several threads (calls are serialised on its workspace),
    different patterns side by side, pageable and pinned buffers mixed (the copy pool serves all of them)."""
    import threading
    blob2, blob3 = nb.compile_to_bytes(SSN, 0), nb.compile_to_bytes(workloads.REGEX["c3"], 0)
    shared, own3 = nb.Pattern(blob2, device=0), nb.Pattern(blob3, device=0)
    ora2, ora3 = Oracle(blob2), Oracle(blob3)
    d2, o2 = workloads.c2_lines(400_000)
    d3, o3 = workloads.c3_lines(300_000)
    want2 = ora2.match_batch(2, d2, o2, threads=8)
    want3 = ora3.match_batch(2, d3, o3, threads=8)
    errors = []

    def worker(k):
        try:
            for it in range(4):
                if k % 2 == 0:
                    got = shared.match_batch(2, d2, o2)
                    want = want2
                else:
                    pat = own3 if k == 1 else nb.Pattern(blob3, device=0)
                    got = pat.match_batch(2, d3, o3)
                    want = want3
                for g, w in zip(got, want):
                    if not np.array_equal(g, w):
                        errors.append((k, it))
        except Exception as exc:  # noqa: BLE001
            errors.append((k, repr(exc)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_one_long_haystack_through_the_batch_call():
    """Matcher.find() on a document: a batch of ONE long haystack is routed to the chunk-parallel single-haystack path; same results
    as the oracle's find(), for byte and UTF-16 haystacks, and the short-haystack / other-mode paths are untouched."""
    text8, _ = workloads.c3_lines(60_000)  # ~3.8 MB
    for regex in (workloads.REGEX["c3"], workloads.REGEX["c2"], r"q[a-z ]*7", "Sherlock|Street"):
        blob = nb.compile_to_bytes(regex, 0)
        pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
        for data, cw in ((text8, 1), (text8[:1_500_000].astype(np.uint16).view(np.uint8), 2)):
            n_chars = data.size // cw
            for lo, hi in ((0, n_chars), (12_345, n_chars - 7), (0, 65_536), (100, 65_635), (0, 65_535)):
                off = np.array([lo, hi], dtype=np.uint64)
                for mode in (2, 1, 0):
                    got = pat.match_batch(mode, data, off, cw)
                    want = ora.match_batch(mode, data, off, cw)
                    for g, w in zip(got, want):
                        assert g is None or np.array_equal(g, w), (regex, cw, lo, hi, mode)
    # a handful of long haystacks, and one short one among them (then the batch kernels take the whole batch)
    blob = nb.compile_to_bytes(workloads.REGEX["c3"], 0)
    pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
    for cuts in ([0, 70_000, 1_000_000, 1_070_000, 3_000_000], [0, 70_000, 70_010, 1_500_000], [5, 100_000, 100_000 + 65_536]):
        off = np.array(cuts, dtype=np.uint64)
        got = pat.match_batch(2, text8, off, 1)
        want = ora.match_batch(2, text8, off, 1)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), cuts
    m = nb.DFACompiler.compile(workloads.REGEX["c2"], "Ssn").matcher("x" * 200_000 + "123-45-6789" + "y" * 1000)
    assert m.find() and (m.start(), m.end()) == (200_000, 200_011)
