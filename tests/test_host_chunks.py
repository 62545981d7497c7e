"""The host half of ndl_match_batch's NDL_MEM_HOST path: how a batch is cut into pipeline chunks and how offsets are
checked (needle_b200/csrc/host_chunks.h).  CPU only - ndl_debug_plan_chunks runs the same functions capi_device.cu calls.

Regression for the out-of-bounds reads a dominant first / last line used to cause (the chunk after the one that reached
the end of the batch started at line n and read offsets[n + 1])."""
import ctypes

import numpy as np
import pytest

from needle_b200 import _lib


def plan(offsets, n, line_chars=0, char_width=1, chunk_bytes=64 << 20, max_chunks=16):
    L = _lib.lib()
    L.ndl_debug_plan_chunks.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]
    L.ndl_debug_plan_chunks.restype = ctypes.c_int
    bounds = np.zeros(max_chunks + 1, dtype=np.uint64)
    flags = np.zeros(max_chunks, dtype=np.int32)
    if offsets is not None:
        # guard words after the array: the planner must never look at them
        guarded = np.concatenate([offsets.astype(np.uint64), np.full(4, 0xDEADBEEFDEADBEEF, dtype=np.uint64)])
        ptr = guarded.ctypes.data
    else:
        ptr = None
    k = L.ndl_debug_plan_chunks(ptr, line_chars, n, char_width, chunk_bytes, max_chunks, bounds.ctypes.data, flags.ctypes.data)
    return [int(b) for b in bounds[:k + 1]], [int(f) for f in flags[:k]]


def check_partition(bounds, n):
    assert bounds[0] == 0 and bounds[-1] == n
    assert all(a < b for a, b in zip(bounds, bounds[1:])), bounds  # every chunk is non-empty, none past n


@pytest.mark.parametrize("lens", [
    [100] * 999 + [200 << 20],             # one dominant LAST line (the advisor's first case)
    [100, 100, 70 << 20],                  # three lines, the last one larger than a chunk
    [200 << 20] + [100] * 999,             # one dominant FIRST line
    [100] * 10 + [300 << 20] + [100] * 10,  # dominant line in the middle
    [0] * 50 + [130 << 20],                # empty lines, then everything in the last
    [130 << 20],                           # a single line
    [1 << 20] * 200,                       # regular
])
def test_skewed_batches_never_leave_the_arrays(lens):
    offsets = np.concatenate([[0], np.cumsum(np.array(lens, dtype=np.uint64))]).astype(np.uint64)
    n = len(lens)
    bounds, flags = plan(offsets, n)
    check_partition(bounds, n)
    assert all(f & 2 for f in flags)


def test_chunks_balance_by_bytes():
    rng = np.random.default_rng(5)
    lens = rng.integers(8, 121, size=4_000_000)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bounds, flags = plan(offsets, len(lens))
    check_partition(bounds, len(lens))
    assert len(bounds) - 1 == int(offsets[-1]) // (64 << 20) + 1
    sizes = [int(offsets[b] - offsets[a]) for a, b in zip(bounds, bounds[1:])]
    assert max(sizes) - min(sizes) <= 240
    assert all(f == 2 for f in flags)  # ragged: monotonic, not uniform


def test_fixed_length_lines_and_uniform_offsets():
    n, L = 3_000_000, 64
    bounds, flags = plan(None, n, line_chars=L)
    check_partition(bounds, n)
    assert len(bounds) - 1 == n * L // (64 << 20) + 1
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(L) + np.uint64(12345)
    b2, f2 = plan(offsets, n)
    check_partition(b2, n)
    assert all(f == 3 for f in f2)  # equally spaced and monotonic: the offsets stay on the host
    offsets[n // 2] += 1
    b3, f3 = plan(offsets, n)
    assert sorted(set(f3)) == [2, 3] and f3.count(2) in (1, 2)


def test_decreasing_offsets_are_flagged():
    offsets = np.array([0, 100, 90, 300, 400], dtype=np.uint64)
    bounds, flags = plan(offsets, 4, chunk_bytes=64, max_chunks=16)
    check_partition(bounds, 4)
    assert any(not (f & 2) for f in flags)


def test_more_chunks_than_lines():
    offsets = np.array([0, 1 << 30, 2 << 30], dtype=np.uint64)
    bounds, _ = plan(offsets, 2)
    assert bounds == [0, 1, 2]


def test_copy_pool_serves_concurrent_callers():
    """The pinned-ring staging of pageable input (host_staging.h) is used by all replica threads of a multi-device pattern at once."""
    import ctypes
    L = _lib.lib()
    L.ndl_debug_parallel_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, size=(64 << 20) + 12345, dtype=np.uint8)
    for callers in (1, 2, 5, 8):
        dst = np.zeros_like(src)
        threads = L.ndl_debug_parallel_copy(dst.ctypes.data, src.ctypes.data, src.size, callers)
        assert threads >= 1
        assert np.array_equal(dst, src), callers
