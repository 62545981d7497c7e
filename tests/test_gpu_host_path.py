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
