"""SWAR table image (needle_b200/csrc/kernels/swar_plan.h, linesq_layout): built exactly as ndl_pattern_create
builds it and walked ON THE HOST with the kernel's integer arithmetic (packed compares, IDP.4A weights, entry
decoding) by the ndl_debug_swar_emulate test hook, against a plain walk of the device table.  No GPU needed:
this pins the plan solver and the image layout; tests/test_gpu_parity.py pins the kernel itself."""
import ctypes

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200 import _lib
from tests import workloads


def emulate(blob, mode, cw, backward, lane, data, n_chars):
    L = _lib.lib()
    L.ndl_debug_swar_emulate.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int32)]
    L.ndl_debug_swar_emulate.restype = ctypes.c_int
    info = (ctypes.c_int32 * 4)()
    rc = L.ndl_debug_swar_emulate(blob, len(blob), mode, cw, backward, lane, data.ctypes.data, n_chars, info)
    return rc, {"char_mode": info[0], "copies": info[1], "codes": info[2], "bytes": info[3]}


def cm_swar(k, planes, hi, u16=False, wide=False):
    return 16 | (64 if wide else 0) | (32 if u16 else 0) | (8 if k == 4 else 0) | (4 if hi else 0) | planes


def byte_soup(rng, n, hot):
    """All 256 byte values, biased towards `hot` (the pattern's own alphabet) so that walks leave the root."""
    data = rng.integers(0, 256, size=n, dtype=np.uint8)
    h = np.frombuffer(hot, dtype=np.uint8)
    pick = rng.random(n) < 0.7
    data[pick] = h[rng.integers(0, len(h), size=int(pick.sum()))]
    return data


BYTE_CASES = [
    # regex, hot alphabet, expected (k, planes, copies, codes, 16-bit entries) for mode find or None = "just has to be right"
    (workloads.REGEX["c2"], b"0123456789--- /:,.", (4, 2, 16, 3, False)),
    (workloads.REGEX["c4"], b"aaabbbc`d", (4, 2, 1, 4, True)),
    (r"[0-9]+", b"0123456789/: ", None),
    (r"a*", b"a`b", None),
    (r"[^a]+b", b"ab`c\x7f\x80", None),
    (r"[a-c]z|[b-d]z", b"abcdz`ey", None),
    (r"\d+-\d+", b"0123456789-,.", None),
    ("x[\x01-\x1f]y", b"xy\x00\x1f\x20\x01", None),
    ("[\x7f]+a", b"a\x7f\x7e\x80\xff", None),
    (r"(ab|a|b-)+", b"ab-,.`c", None),
    (r"a[ab]{7}c|b[ab]{4}d", b"aaabbbcd`e", (2, 3, 8, 5, True)),
    (r"a[ab]{5}c", b"aaabbbc`d", None),
    (r"a[ab]{4}c", b"aaabbbc`d", None),
]


@pytest.mark.parametrize("regex,hot,expect", BYTE_CASES, ids=[c[0][:20] for c in BYTE_CASES])
def test_byte_images_walk_like_the_table(regex, hot, expect):
    blob = nb.compile_to_bytes(regex, 0)
    rng = np.random.default_rng(abs(hash(regex)) % (1 << 32))
    seen = 0
    for mode in (0, 1, 2):
        for n in (0, 1, 3, 4, 5, 64, 4001):
            data = byte_soup(rng, max(n, 1), hot)
            for lane in (0, 5, 31):
                rc, info = emulate(blob, mode, 1, 0, lane, data, n)
                assert rc in (0, -1), (regex, mode, n, lane, rc, info)
                if rc == 0:
                    seen += 1
                    if expect and mode == 2:
                        k, planes, copies, codes, u16 = expect
                        assert info["char_mode"] == cm_swar(k, planes, False, u16) and info["copies"] == copies and info["codes"] == codes, info
    assert seen > 0, "no SWAR image for any mode"
    if expect:
        assert emulate(blob, 2, 1, 0, 0, np.zeros(16, dtype=np.uint8), 16)[0] == 0


def test_backward_rows():
    # variable-length patterns: BACKWARDS rows share the image and the classifier (joint classes)
    for regex, hot in ((r"[0-9]+x", b"0123456789x/:"), (r"a+b+", b"ab`c"), (r"\d+-\d+", b"0123456789-,.")):
        blob = nb.compile_to_bytes(regex, 0)
        rng = np.random.default_rng(5)
        for n in (1, 4, 7, 64, 1001):
            data = byte_soup(rng, n, hot)
            for lane in (0, 17):
                assert emulate(blob, 2, 1, 1, lane, data, n)[0] == 0, (regex, n, lane)
                assert emulate(blob, 2, 1, 0, lane, data, n)[0] == 0, (regex, n, lane)


def test_utf16_high_byte_images():
    blob = nb.compile_to_bytes(workloads.REGEX["c5"], 0)
    rng = np.random.default_rng(9)
    for n in (0, 1, 4, 6, 32, 3001):
        chars = rng.integers(0, 0x10000, size=max(n, 1)).astype(np.uint16)
        hot = rng.random(len(chars)) < 0.5
        chars[hot] = rng.integers(0x5F0, 0x710, size=int(hot.sum()))
        chars[::97] = 0xFFFF
        data = chars.view(np.uint8)
        for mode in (0, 1, 2):
            for backward in ((0, 1) if mode == 2 else (0,)):
                rc, info = emulate(blob, mode, 2, backward, 3, data, n)
                assert rc == 0, (mode, backward, n, rc, info)
                assert info["char_mode"] == cm_swar(4, 1, True) and info["copies"] == 16


WIDE_CASES = [
    # ASCII / BMP-range patterns over UTF-16 text (a java.lang.String): compares on 16-bit lanes
    (workloads.REGEX["c2"], "0123456789--- /:,.\u0130\u0660", (4, 2, 3)),
    (workloads.REGEX["c4"], "aaabbbc`d\u0161", (2, 2, 4)),
    ("[0-9]+x", "0123456789x/:\u0439", None),
    ("[\u03b1-\u03c9]+", "\u03b0\u03b1\u03c9\u03ca ab", None),
    ("[a-c\u0410-\u042f]x", "abcdx\u040f\u0410\u042f\u0430`", None),
]


@pytest.mark.parametrize("regex,hot,expect", WIDE_CASES, ids=[c[0][:16] for c in WIDE_CASES])
def test_utf16_16bit_lane_images(regex, hot, expect):
    blob = nb.compile_to_bytes(regex, 0)
    rng = np.random.default_rng(abs(hash(regex)) % (1 << 32))
    hot_units = np.array([ord(ch) for ch in hot], dtype=np.uint16)
    seen = 0
    for n in (0, 1, 3, 4, 5, 8, 64, 3001):
        chars = rng.integers(0, 0x10000, size=max(n, 1)).astype(np.uint16)
        pick = rng.random(len(chars)) < 0.7
        chars[pick] = hot_units[rng.integers(0, len(hot_units), size=int(pick.sum()))]
        chars[::53] = 0xFFFF
        chars[7::61] = 0x8000
        chars[3::67] = 0x7FFF
        data = chars.view(np.uint8)
        for mode in (0, 1, 2):
            for backward in (0, 1):
                for lane in (0, 9):
                    rc, info = emulate(blob, mode, 2, backward, lane, data, n)
                    if backward and rc == -2:
                        continue  # no table-driven reverse pass for this pattern / mode
                    assert rc in (0, -1), (regex, mode, backward, n, rc, info)
                    if rc == 0:
                        seen += 1
                        assert info["char_mode"] & 64, info
                        if expect and mode == 2:
                            k, planes, codes = expect
                            assert info["char_mode"] == cm_swar(k, planes, False, wide=True) and info["codes"] == codes, info
    assert seen > 0


def test_class_maps_without_a_plan_are_refused():
    for regex, cw in ((workloads.REGEX["c3"], 1), (workloads.REGEX["c1"], 1), ("[Ss]herlock", 1), (workloads.REGEX["c3"], 2), ("é+", 1)):
        blob = nb.compile_to_bytes(regex, 0)
        assert emulate(blob, 2, cw, 0, 0, np.zeros(16, dtype=np.uint8), 8)[0] == -1, regex
