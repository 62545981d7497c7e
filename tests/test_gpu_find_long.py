"""ndl_find_long: find() over one long haystack (BASELINE config 4), chunk-parallel on the GPU, against the
oracle's sequential walk with 64-bit offsets.  Needs a CUDA device."""
import ctypes

import numpy as np
import pytest

import needle_b200 as nb
from tests import workloads
from tests.oracle_lib import INT64_MAX, Oracle, lib as oracle_lib

pytestmark = pytest.mark.gpu

_P = {}


def pair(regex, flags=0):
    key = (regex, flags)
    if key not in _P:
        blob = nb.compile_to_bytes(regex, flags)
        _P[key] = (nb.Pattern(blob, device=0), Oracle(blob))
    return _P[key]


def oracle_find_long(ora, data, from_=0, cw=1):
    st, en = ctypes.c_int64(), ctypes.c_int64()
    data = np.ascontiguousarray(data).view(np.uint8)
    m = oracle_lib().ndlo_find(ora._h, data.ctypes.data, data.size // cw, cw, from_, INT64_MAX, ctypes.byref(st), ctypes.byref(en))
    return bool(m), st.value, en.value


def ab_buffer(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 2, size=n, dtype=np.uint8) + ord("a")).astype(np.uint8)


@pytest.mark.parametrize("n", [0, 1, 8, 9, 10, 100, 255, 256, 257, 271, 272, 273, 2047, 2048, 2049, 4096 + 9, 8191, 8192, 8193, 8192 + 255,
                               8192 + 256 + 17, 70_001, 3_000_000])
def test_c4_match_at_the_very_end(n):
    pat, ora = pair(workloads.REGEX["c4"])
    data = ab_buffer(n, n)
    if n >= 9:
        data[n - 9] = ord("a")
        data[n - 1] = ord("c")
    got = pat.find_long(data)
    assert got == oracle_find_long(ora, data)
    if n >= 9:
        assert got == (True, n - 9, n)


def test_c4_match_positions_across_segment_and_tile_boundaries():
    pat, ora = pair(workloads.REGEX["c4"])
    n = 300_000
    base = ab_buffer(n, 3)
    for pos in [0, 1, 55, 56, 63, 64, 240, 247, 248, 255, 256, 2040, 2047, 2048, 2050, 4090, 8184, 8190, 8192, 65_530, 131_072 - 4, 250_000,
                n - 9]:
        data = base.copy()
        data[pos] = ord("a")
        data[pos + 8] = ord("c")
        got = pat.find_long(data)
        assert got == oracle_find_long(ora, data) == (True, pos, pos + 9), pos
    # two matches: the leftmost wins; and find(from) skips the first
    data = base.copy()
    for pos in (10_000, 200_000):
        data[pos] = ord("a")
        data[pos + 8] = ord("c")
    assert pat.find_long(data) == (True, 10_000, 10_009)
    assert pat.find_long(data, from_=10_001) == oracle_find_long(ora, data, 10_001) == (True, 200_000, 200_009)
    assert pat.find_long(data, from_=200_001) == (False, -1, -1)


def test_unaligned_data_pointer_and_from():
    pat, ora = pair(workloads.REGEX["c4"])
    big = ab_buffer(500_000 + 37, 5)
    for shift in (1, 7, 15, 37):
        data = big[shift:]
        data2 = data.copy()
        data2[400_000] = ord("a")
        data2[400_008] = ord("c")
        for frm in (0, 3, 2048, 399_999):
            assert pat.find_long(data2, from_=frm) == oracle_find_long(ora, data2, frm)


@pytest.mark.parametrize("regex", [workloads.REGEX["c2"], workloads.REGEX["c3"], "Sherlock|Street", "[0-9]+x", "needle"])
def test_other_patterns_on_text(regex):
    pat, ora = pair(regex)
    data, offsets = workloads.c3_lines(40_000)  # ~2.5 MB of text with planted e-mail addresses
    data = data.copy()
    data[2_000_000:2_000_011] = np.frombuffer(b"123-45-6789", dtype=np.uint8)
    data[2_100_000:2_100_008] = np.frombuffer(b"Sherlock", dtype=np.uint8)
    data[2_200_000:2_200_006] = np.frombuffer(b"needle", dtype=np.uint8)
    for frm in (0, 1_000_000, 2_050_000, 2_150_000, 2_300_000):
        assert pat.find_long(data, from_=frm) == oracle_find_long(ora, data, frm), (regex, frm)


def test_long_memory_pattern_falls_back_to_the_exact_walk():
    # `a.*c` (no newline in the data): the state after an 'a' persists, so the 16-byte guess is wrong and
    # the call must notice and fall back.  Kept small: the fallback is a single-thread walk.
    pat, ora = pair("q[a-z ]*7")
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"abcdefghijklmnop rstuvwxyz", dtype=np.uint8)
    data = alpha[rng.integers(0, len(alpha), size=200_000)].copy()
    data[50_000] = ord("q")
    data[150_000] = ord("7")
    assert pat.find_long(data) == oracle_find_long(ora, data) == (True, 50_000, 150_001)
    from needle_b200 import _lib
    assert _lib.lib().ndl_debug_long_passes() == -1  # 100 000 chars of memory: refinement gives up after its bounded passes


def test_long_memory_pattern_is_refined_when_its_memory_is_bounded():
    """`q[a-z ]*7` over text whose runs of [a-z ] are a few hundred chars long: the 16-byte warm-up guesses wrong after every
    'q', but the automaton's memory ends at the next digit or punctuation mark, so a few refinement passes (every segment
    entering in its predecessor's exit state of the previous pass) make the chunk-parallel scan exact - no single-thread walk."""
    from needle_b200 import _lib
    pat, ora = pair("q[a-z ]*7")
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz      ", dtype=np.uint8)
    n = 48 << 20
    data = alpha[rng.integers(0, len(alpha), size=n)].copy()
    breaks = rng.integers(0, n, size=n // 300)
    data[breaks] = np.frombuffer(b".,;:!?0123456", dtype=np.uint8)[rng.integers(0, 13, size=len(breaks))]
    data[data == ord("7")] = ord("8")  # no match ...
    assert pat.find_long(data) == oracle_find_long(ora, data) == (False, -1, -1)
    assert 1 < _lib.lib().ndl_debug_long_passes() <= 10
    data[n - 5_000_000 + 17] = ord("q")  # ... then one, 40 chars long, deep in the buffer
    data[n - 5_000_000 + 18:n - 5_000_000 + 57] = ord("e")
    data[n - 5_000_000 + 57] = ord("7")
    got = pat.find_long(data)
    assert got == oracle_find_long(ora, data) and got[0] and got[2] == n - 5_000_000 + 58
    assert 1 < _lib.lib().ndl_debug_long_passes() <= 10


def test_accepting_root_and_utf16():
    pat, ora = pair("a*")
    data = np.frombuffer(b"baaaa" * 1000, dtype=np.uint8)
    assert pat.find_long(data) == oracle_find_long(ora, data) == (True, 0, 0)
    assert pat.find_long(data, from_=1) == oracle_find_long(ora, data, 1)
    pat, ora = pair(workloads.REGEX["c5"])
    d16, _ = workloads.c5_lines(2000)
    assert pat.find_long(d16, char_width=2) == oracle_find_long(ora, d16, 0, 2)


def test_device_pointer_entry_and_gigabyte_scale():
    """1 GiB of {a,b} generated on the device with the only match at the very end: the whole buffer must be
    scanned, and the answer is known by construction (no 'c' anywhere else)."""
    torch = pytest.importorskip("torch")
    pat, _ = pair(workloads.REGEX["c4"])
    n = 1 << 30
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0004)
    data = torch.randint(ord("a"), ord("b") + 1, (n,), dtype=torch.uint8, device="cuda", generator=g)
    tail = torch.tensor(list(b"abababbac"), dtype=torch.uint8, device="cuda")
    data[n - 9:] = tail
    assert pat.find_long_ptrs(data.data_ptr(), n) == (True, n - 9, n)
    data[n - 1] = ord("b")
    assert pat.find_long_ptrs(data.data_ptr(), n) == (False, -1, -1)


def test_device_haystack_with_device_or_host_results():
    """NDL_MEM_DEVICE writes the three results through device pointers, NDL_MEM_DEVICE_DATA through host pointers."""
    import torch
    from needle_b200 import _lib
    pat, ora = pair(workloads.REGEX["c4"])
    data = ab_buffer(1_000_000, 21)
    data[777_000], data[777_008] = ord("a"), ord("c")
    want = oracle_find_long(ora, data)
    assert want == (True, 777_000, 777_009)
    d = torch.from_numpy(data).cuda()
    assert pat.find_long_ptrs(d.data_ptr(), d.numel(), 1, 0, nb.MEM_DEVICE) == want  # (host results under the hood)
    out = torch.zeros(3, dtype=torch.int64, device="cuda")
    rc = _lib.lib().ndl_find_long(pat._h, d.data_ptr(), d.numel(), 1, 0, out.data_ptr() + 16, out.data_ptr(), out.data_ptr() + 8,
                                  _lib.MEM_DEVICE, None)
    assert rc == _lib.NDL_OK
    o = out.cpu().tolist()
    assert (bool(o[2] & 0xFF), o[0], o[1]) == want
    assert pat.find_long_from(d.data_ptr(), d.numel(), 0, mem_kind=nb.MEM_DEVICE)[0] == 777_009


def test_host_walk_gives_the_state_the_device_walk_gives():
    """ndl_forwards_walk_host (the entry-state guess of a sharded find) against ndl_find_long_from's exit state."""
    rng = np.random.default_rng(8)
    for regex in (workloads.REGEX["c4"], r"q[a-z ]*7", workloads.REGEX["c3"], r"Sherlock|Street"):
        pat, _ = pair(regex)
        for k in range(20):
            n = int(rng.integers(0, 40))
            halo = np.frombuffer(bytes(rng.choice(list(b"abcq7 @.Shtre"), size=n).astype(np.uint8)), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
            entry = int(rng.integers(0, pat.forwards_state_count))
            host = pat.walk_host(halo, entry)
            dev = pat.find_long_from(halo.ctypes.data if n else 0, n, entry, mem_kind=nb.MEM_HOST)[1]
            assert host == dev, (regex, halo.tobytes(), entry)


def test_utf16_haystacks_take_the_chunk_parallel_path():
    """char_width 2 (a java.lang.String's payload): the same segments / guesses / checks over UTF-16 code units - compares on 16-bit
    lanes for ASCII patterns, class from the high byte for BMP classes, class-map images otherwise."""
    from needle_b200 import _lib
    passes = _lib.lib().ndl_debug_long_passes
    n = 1_500_000
    ab = ab_buffer(n + 64, 31).astype(np.uint16)
    pat, ora = pair(workloads.REGEX["c4"])
    for shift in (0, 1, 3, 8, 21):  # 2-byte aligned, not 16
        base = ab[shift:shift + n]
        assert pat.find_long(base, char_width=2) == oracle_find_long(ora, base, 0, 2) == (False, -1, -1)
        assert passes() >= 1
        for pos in (0, 5, 127, 128, 120, 4090, 4096, 1_000_000, n - 9):
            data = base.copy()
            data[pos], data[pos + 8] = ord("a"), ord("c")
            assert pat.find_long(data, char_width=2) == oracle_find_long(ora, data, 0, 2) == (True, pos, pos + 9), (shift, pos)
        data = base.copy()
        for pos in (10_000, 900_000):
            data[pos], data[pos + 8] = ord("a"), ord("c")
        for frm in (0, 3, 10_001, 899_999, 900_001):
            assert pat.find_long(data, from_=frm, char_width=2) == oracle_find_long(ora, data, frm, 2), (shift, frm)
    # chars above 0xff must not alias ASCII ones: 0x0161 / 0x6100 are not 'a'
    data = ab[:n].copy()
    data[500_000:500_009] = np.array([0x0161, 0x62, 0x61, 0x62, 0x61, 0x62, 0x62, 0x61, 0x63], dtype=np.uint16)
    data[700_000:700_009] = np.array([0x61, 0x62, 0x6100, 0x62, 0x61, 0x62, 0x62, 0x61, 0x63], dtype=np.uint16)
    assert pat.find_long(data, char_width=2) == oracle_find_long(ora, data, 0, 2) == (False, -1, -1)

    # a BMP class (BASELINE config 5's regex) and class-map patterns over UTF-16 text
    text8, _ = workloads.c3_lines(30_000)
    text = text8.astype(np.uint16)
    for regex, plant in ((workloads.REGEX["c5"], [0x0627, 0x0644, 0x0639]), ("Sherlock|Street", [ord(c) for c in "Street"]),
                         (r"q[a-z ]*7", [ord(c) for c in "q the quick brown fox 7"]), (r"[Ѐ-ӿ]+[0-9]", [0x0416, 0x0417, ord("4")])):
        pat, ora = pair(regex)
        data = text.copy()
        assert pat.find_long(data, char_width=2) == oracle_find_long(ora, data, 0, 2), regex
        for pos in (1_200_000, 640_000 - 2, 77):
            d2 = data.copy()
            d2[pos:pos + len(plant)] = np.array(plant, dtype=np.uint16)
            for frm in (0, 100, pos + 1):
                assert pat.find_long(d2, from_=frm, char_width=2) == oracle_find_long(ora, d2, frm, 2), (regex, pos, frm)
    # which of them ran chunk-parallel is a property of the pattern's UTF-16 images; the two bench regexes must
    for regex in (workloads.REGEX["c4"], workloads.REGEX["c5"]):
        pat, ora = pair(regex)
        pat.find_long(text, char_width=2)
        assert passes() >= 1, regex
