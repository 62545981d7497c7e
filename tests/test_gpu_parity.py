"""GPU parity: the CUDA kernels, called through the C ABI (ndl_match_batch), against the CPU oracle and
the reference's golden vectors - bit exact on (matched, start, end).  Needs a CUDA device."""
import json
import os

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200 import _lib
from tests import kats, workloads
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu

_P = {}


def pair(regex, flags=0):
    """(GPU pattern, oracle) for a regex, cached."""
    key = (regex, flags)
    if key not in _P:
        blob = nb.compile_to_bytes(regex, flags)
        _P[key] = (nb.Pattern(blob, device=0), Oracle(blob))
    return _P[key]


def assert_batch_equal(regex, flags, data, offsets, cw=1, from_=None, modes=(0, 1, 2)):
    pat, ora = pair(regex, flags)
    for mode in modes:
        got = pat.match_batch(mode, data, offsets, cw, from_ if mode == 2 else None)
        exp = ora.match_batch(mode, data, offsets, cw, from_ if mode == 2 else None, threads=8)
        for name, g, e in zip(("matched", "start", "end"), got, exp):
            if e is None:
                continue
            if not np.array_equal(g, e):
                bad = np.nonzero(g != e)[0]
                i = int(bad[0])
                o0, o1 = int(offsets[i]), int(offsets[i + 1])
                raise AssertionError(f"{regex!r} flags={flags:#x} mode={mode} {name}: {len(bad)} of {len(e)} differ; first i={i} "
                                     f"gpu={g[i]} oracle={e[i]} haystack={bytes(data[o0 * cw:o1 * cw])!r}")


def test_native_library_is_the_one_loaded():
    assert os.path.exists(_lib.LIB_PATH)
    assert _lib.lib().ndl_device_count() >= 1
    before = _lib.lib().ndl_kernel_launches()
    pat, _ = pair("abc")
    pat.match_batch(2, np.frombuffer(b"xxabcxx", dtype=np.uint8), np.array([0, 7], dtype=np.uint64))
    assert _lib.lib().ndl_kernel_launches() > before


def test_matches_txt_all_rows_on_gpu(golden_dir):
    with open(os.path.join(golden_dir, "matches.json"), encoding="utf-8") as f:
        rows = json.load(f)
    bad = []
    for r in rows:
        fl = r["flags"] if r["flags"] is not None else r["java_random_flags"]
        pat, ora = pair(r["pattern"], fl)
        m = pat.matcher(r["haystack"])
        found = m.find()
        got = (found, m.start(), m.end())
        exp = (r["matched"], r["start"], r["end"])
        if got != exp:
            bad.append((r["line"], r["pattern"], r["haystack"], hex(fl), got, exp))
        # and the other two entry points agree with the oracle
        assert pat.matcher(r["haystack"]).matches() == ora.matches(r["haystack"]), r
        assert pat.matcher(r["haystack"]).containedIn() == ora.contained_in(r["haystack"]), r
    assert not bad, bad


@pytest.mark.parametrize("regex,flags,match,fail,contained", kats.MATCH_FAIL, ids=[k[0][:30] for k in kats.MATCH_FAIL])
def test_inline_match_fail_on_gpu(regex, flags, match, fail, contained):
    pat, _ = pair(regex, flags)
    for s in match:
        m = pat.matcher(s)
        assert m.matches() and m.containedIn(), s
        assert m.find() and (m.start(), m.end()) == (0, len(s)), s
    for s in fail:
        assert not pat.matcher(s).containedIn() and not pat.matcher(s).matches(), s
    for s in contained:
        assert not pat.matcher(s).matches() and pat.matcher(s).containedIn(), s


@pytest.mark.parametrize("regex,flags,hay,frm,exp", kats.FIND)
def test_inline_find_on_gpu(regex, flags, hay, frm, exp):
    pat, _ = pair(regex, flags)
    m = pat.matcher(hay)
    found = m.find(frm, len(hay))
    assert (found, m.start(), m.end()) == exp


@pytest.mark.parametrize("regex,flags,hay,exp", kats.FIND_ALL)
def test_iterated_find_on_gpu(regex, flags, hay, exp):
    pat, _ = pair(regex, flags)
    assert list(nb.iter_find(pat, hay)) == exp
    m = pat.matcher(hay)
    while m.find():
        pass
    assert not m.find() and not m.find()  # findDoesntRollOver (DFACompilerTest.java:816-825)


def test_c1_url_strings():
    strings = workloads.c1_strings(1000)
    data, offsets, cw = nb.pack_haystacks(strings)
    assert_batch_equal(workloads.REGEX["c1"], 0, data, offsets, cw)
    pat, _ = pair(workloads.REGEX["c1"])
    m, s, e = pat.match_batch(2, data, offsets, cw)
    for i, st in enumerate(strings):
        assert bool(m[i]) == ("http://" in st and len(st) > st.index("http://") + 7)


@pytest.mark.parametrize("n", [1, 2, 31, 1007, 1008, 1009, 4096, 200_003])
def test_c2_ssn_fixed_lines(n):
    data, offsets = workloads.c2_lines(n)
    assert_batch_equal(workloads.REGEX["c2"], 0, data, offsets)
    pat, _ = pair(workloads.REGEX["c2"])
    m, s, e = pat.match_batch(2, data, offsets)
    assert np.all((e - s)[m == 1] == 11)  # fixed-length pattern: start = end - 11


@pytest.mark.parametrize("line_len", [16, 32, 64, 128, 48, 80, 96, 112, 144, 256, 11, 1])
def test_fixed_lines_all_lengths(line_len):
    rng = np.random.default_rng(line_len)
    n = 5000
    alpha = np.frombuffer(b"0123456789-ab ", dtype=np.uint8)
    data = alpha[rng.integers(0, len(alpha), size=n * line_len)]
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_len)
    for regex in (workloads.REGEX["c2"], r"[0-9]+", r"a*", r"(ab|a|b-)+", r"\d+-\d+"):
        assert_batch_equal(regex, 0, data, offsets)
    # a line count that is not a multiple of the tile, one irregular line in the middle, a base that is not 16-byte aligned
    off2 = offsets[:4001].copy()
    off2[2000:] += 3
    assert_batch_equal(r"\d+-\d+", 0, data, off2)
    assert_batch_equal(workloads.REGEX["c3"], 0, data[5:], offsets[:4000])


@pytest.mark.parametrize("line_len", [80, 96, 112, 128, 160, 256, 272, 512, 1040, 4096, 100, 40, 24, 17, 31, 200, 333, 447, 16384, 65536, 65552, 5000])
def test_fixed_lines_in_rounds(line_len):
    """Records whose 32-line tile does not fit a buffer are walked in rounds of 64 bytes (every lane busy): fixed-length and
    variable-length patterns, UTF-16, a batch that is not a multiple of the tile, irregular tiles in between."""
    rng = np.random.default_rng(line_len)
    n = 32 * 57 + 19 if line_len <= 4096 else 32 * 3 + 5  # (65552 bytes is beyond the rounds walk: the ragged walk streams it)
    alpha = np.frombuffer(b"0123456789-ab @.", dtype=np.uint8)
    data = alpha[rng.integers(0, len(alpha), size=n * line_len)]
    # matches at the very start / end of a record and across round boundaries
    for i in range(0, n, 7):
        pos = [0, line_len - 11, 60, 120, 64][i % 5] % (line_len - 10) if line_len > 11 else 0
        data[i * line_len + pos:i * line_len + pos + 11] = np.frombuffer(b"123-45-6789", dtype=np.uint8)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_len)
    for regex in (workloads.REGEX["c2"], workloads.REGEX["c4"], r"[0-9]+", workloads.REGEX["c3"], r"(ab|a|b-)+", "9"):
        assert_batch_equal(regex, 0, data, offsets)
    off2 = offsets.copy()
    off2[min(1000, n // 2):] += 16  # one irregular line (tile), the rest regular again
    assert_batch_equal(workloads.REGEX["c2"], 0, np.concatenate([data, data[:16]]), off2)
    assert_batch_equal(r"[0-9]+", 0, data[16:], offsets[:n - 1])  # base moved by a whole chunk
    assert_batch_equal(r"[0-9]+", 0, data[3:], offsets[:n - 1])   # unaligned base: every tile irregular
    if line_len <= 512:
        wide = data.astype(np.uint16)
        o16 = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_len)
        for regex in (workloads.REGEX["c2"], r"[0-9]+", workloads.REGEX["c5"]):
            assert_batch_equal(regex, 0, wide.view(np.uint8), o16, cw=2)
    pat, ora = pair(workloads.REGEX["c2"])
    m, s_, e = pat.match_lines(2, data, n, line_len)
    em, es, ee = ora.match_batch(2, data, offsets, 1, threads=4)
    assert np.array_equal(m, em) and np.array_equal(s_, es) and np.array_equal(e, ee)


def test_unaligned_base_and_sub_batches():
    data, offsets = workloads.c2_lines(3000)
    pat, ora = pair(workloads.REGEX["c2"])
    # a batch that starts in the middle of the buffer (offsets[0] != 0)
    sub = offsets[1000:2501]
    got = pat.match_batch(2, data, sub)
    exp = ora.match_batch(2, data, sub)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)
    # a data pointer that is not 16-byte aligned
    shifted = np.empty(len(data) + 5, dtype=np.uint8)
    shifted[5:] = data
    got = pat.match_batch(2, shifted[5:], offsets)
    exp = ora.match_batch(2, data, offsets)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)


@pytest.mark.parametrize("n", [1, 77, 50_000])
def test_c3_email_ragged_lines(n):
    data, offsets = workloads.c3_lines(n)
    assert_batch_equal(workloads.REGEX["c3"], 0, data, offsets)


def test_c4_256_state_dfa_batched():
    data, offsets = workloads.c4_lines(20_000)
    assert_batch_equal(workloads.REGEX["c4"], 0, data, offsets)
    pat, _ = pair(workloads.REGEX["c4"])
    assert pat.info.n_states[2] >= 256


def test_c5_bmp_utf16_lines():
    data, offsets = workloads.c5_lines(20_000)
    assert_batch_equal(workloads.REGEX["c5"], 0, data, offsets, cw=2)


def test_find_with_from_offsets():
    data, offsets = workloads.c3_lines(5000)
    rng = np.random.default_rng(5)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    from_ = (rng.random(len(lens)) * (lens + 1)).astype(np.int32)
    for regex in (workloads.REGEX["c3"], "a*", "[a-z]+", r"\d{2}"):
        assert_batch_equal(regex, 0, data, offsets, from_=from_, modes=(2,))
    # from beyond the end, every kernel family (packed compares, class map in shared memory, single-char reverse scan), fixed lines
    from2 = np.minimum(from_ + rng.integers(0, 3, size=len(lens)).astype(np.int32) * (rng.random(len(lens)) < 0.1), (lens + 2).astype(np.int32))
    for regex in (workloads.REGEX["c2"], workloads.REGEX["c3"], "[0-9]+x", "Sherlock|Street", "a*", "b|a*"):
        assert_batch_equal(regex, 0, data, offsets, from_=from2.astype(np.int32), modes=(2,))
    d2, o2 = workloads.c2_lines(4000)
    f2 = rng.integers(0, 70, size=4000).astype(np.int32)
    assert_batch_equal(workloads.REGEX["c2"], 0, d2, o2, from_=f2, modes=(2,))
    d16, o16 = workloads.c2_lines_utf16(3000)
    assert_batch_equal(workloads.REGEX["c2"], 0, d16, o16, cw=2, from_=rng.integers(0, 34, size=3000).astype(np.int32), modes=(2,))


def test_empty_and_degenerate_batches():
    pat, ora = pair("a*")
    m, s, e = pat.match_batch(2, np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert len(m) == 0
    strings = ["", "", "a", "", "b", ""]
    data, offsets, cw = nb.pack_haystacks(strings)
    assert_batch_equal("a*", 0, data, offsets, cw)
    assert_batch_equal("a+", 0, data, offsets, cw)
    assert_batch_equal("", 0, data, offsets, cw)


def test_reverse_scan_variants():
    # table-driven reverse pass, single-char reverse scan (SherlockStreet snapshot), fixed length
    rng = np.random.default_rng(9)
    words = ["Sherlock", "Street", "Holmes", "Watson", " ", "x", "S", "Sh", "anywhere", "somewhere", "where", "\n"]
    strings = ["".join(words[j] for j in rng.integers(0, len(words), int(rng.integers(0, 12)))) for _ in range(4000)]
    data, offsets, cw = nb.pack_haystacks(strings)
    for regex in ("Sherlock|Street", "[Ss]herlock", "anywhere|somewhere", "Holmes.{1,10}Watson|Watson.{1,10}Holmes",
                  "([Ss]herlock)|([Hh]olmes)", "Sherlock|Holmes|Watson|Irene|Adler|John|Baker", "S.*e", "d|[a-c]x"):
        assert_batch_equal(regex, 0, data, offsets, cw)
    # same data as fixed 64-byte lines so the shared-memory kernel's reverse paths run too
    blob = "".join(strings).encode("latin-1")
    n = len(blob) // 64
    fixed = np.frombuffer(blob[:n * 64], dtype=np.uint8)
    off64 = np.arange(n + 1, dtype=np.uint64) * np.uint64(64)
    for regex in ("Sherlock|Street", "anywhere|somewhere", "([Ss]herlock)|([Hh]olmes)", "S.*e", "[a-z]+"):
        assert_batch_equal(regex, 0, fixed, off64)
    # UTF-16 (16-bit packed compare in the single-char scan) and find(from, to): the scan stops at `from`
    d16, o16, cw16 = nb.pack_haystacks(strings, char_width=2)
    lens = (o16[1:] - o16[:-1]).astype(np.int64)
    frm = (rng.random(len(lens)) * (lens + 1)).astype(np.int32)
    for regex in ("Sherlock|Street", "anywhere|somewhere", "d|[a-c]x"):
        assert_batch_equal(regex, 0, d16, o16, cw16)
        assert_batch_equal(regex, 0, d16, o16, cw16, from_=frm, modes=(2,))
        assert_batch_equal(regex, 0, data, offsets, cw, from_=frm, modes=(2,))


def test_flags_on_gpu():
    strings = ["Sam", "SAMWISE", "samwise", "abc\nabc", "ΓΔΘ γδθ", "x"]
    data, offsets, cw = nb.pack_haystacks(strings)
    for regex, fl in (("sam|samwise", nb.CASE_INSENSITIVE), ("sam|samwise", nb.LEFTMOST_LONGEST | nb.CASE_INSENSITIVE),
                      (".{5}", nb.DOTALL), (".{5}", 0), ("[Γ-Θ]+", nb.CASE_INSENSITIVE | nb.UNICODE_CASE),
                      (r"\w+", nb.UNICODE_CHARACTER_CLASS), (r"\w+", 0)):
        assert_batch_equal(regex, fl, data, offsets, cw)


def test_large_table_patterns():
    # 309-state search DFA (HolmesNearWatson snapshot) and its 2467-state sibling: too big for the replicated shared-memory
    # images, one plain copy of the stride-1 table fits (with the BACKWARDS rows for the first); a[ab]{11}c: 4097 states, 16-bit
    # 4-char table too large, plain stride-1 table fits, as does the one of a[ab]{12}c (8193 states x 4 columns)
    data, offsets = workloads.c2_lines(2000)
    text = np.frombuffer(("Holmes and then Watson said to Mr. Sherlock Holmes, my dear Watson " * 2000).encode()[:2000 * 64], dtype=np.uint8)
    rng = np.random.default_rng(3)
    ragged = np.zeros(2001, dtype=np.uint64)
    ragged[1:] = np.cumsum(rng.integers(0, 129, size=2000))
    ragged = np.minimum(ragged, len(text)).astype(np.uint64)
    for regex in ("Holmes.{1,10}Watson|Watson.{1,10}Holmes", "Holmes.{0,25}Watson|Watson.{0,25}Holmes"):
        fp = fast_path(pair(regex)[0], 2, 1)
        assert fp is not None and fp["char_mode"] == 3 and fp["replicated"] == 1, fp
        assert_batch_equal(regex, 0, text, offsets)
        assert_batch_equal(regex, 0, text, ragged)
    assert fast_path(pair("Holmes.{1,10}Watson|Watson.{1,10}Holmes")[0], 2, 1)["has_bwd"] == 1
    ab = (rng.integers(0, 2, size=2000 * 64, dtype=np.uint8) + ord("a")).astype(np.uint8)
    ab[rng.integers(0, len(ab), size=3000)] = ord("c")
    for regex in ("a[ab]{11}c", "a[ab]{12}c"):
        assert_batch_equal(regex, 0, ab, offsets)
    assert fast_path(pair("a[ab]{11}c")[0], 2, 1)["replicated"] == 1 and fast_path(pair("a[ab]{12}c")[0], 2, 1)["replicated"] == 1


def test_device_memory_entry_point():
    torch = pytest.importorskip("torch")
    data, offsets = workloads.c2_lines(100_000)
    pat, ora = pair(workloads.REGEX["c2"])
    d = torch.from_numpy(data).cuda()
    o = torch.from_numpy(offsets.view(np.int64)).cuda()
    n = len(offsets) - 1
    matched = torch.zeros(n, dtype=torch.uint8, device="cuda")
    start = torch.zeros(n, dtype=torch.int32, device="cuda")
    end = torch.zeros(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    pat.match_batch_ptrs(2, d.data_ptr(), o.data_ptr(), n, 1, matched.data_ptr(), start.data_ptr(), end.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    em, es, ee = ora.match_batch(2, data, offsets, threads=8)
    assert np.array_equal(matched.cpu().numpy(), em)
    assert np.array_equal(start.cpu().numpy(), es)
    assert np.array_equal(end.cpu().numpy(), ee)


def test_full_size_properties_c2():
    """BASELINE config 2 at full size (10 M x 64 B): too big for the oracle in seconds, so check
    size-independent properties: every planted SSN is found, end - start == 11, the matched span
    re-matches, and an oracle spot-check on a random sample of lines."""
    n = 10_000_000
    data, offsets = workloads.c2_lines(n)
    pat, ora = pair(workloads.REGEX["c2"])
    m, s, e = pat.match_batch(2, data, offsets)
    assert 0.25 * n * 0.98 < m.sum()  # >= the planted fraction (random text adds a few more)
    assert np.all((e - s)[m == 1] == 11) and np.all(s[m == 0] == -1) and np.all(e[m == 0] == -1)
    hit = np.nonzero(m)[0]
    starts = offsets[hit].astype(np.int64) + s[hit]
    span = data[starts[:, None] + np.arange(11)[None, :]]
    assert np.all(span[:, 3] == ord("-")) and np.all(span[:, 6] == ord("-"))
    digits = np.delete(span, [3, 6], axis=1)
    assert np.all((digits >= ord("0")) & (digits <= ord("9")))
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(n, size=200_000, replace=False))
    sub_off = np.zeros(len(sample) + 1, dtype=np.uint64)
    sub_off[1:] = np.cumsum(np.full(len(sample), 64, dtype=np.uint64))
    sub = data.reshape(n, 64)[sample].reshape(-1)
    em, es, ee = ora.match_batch(2, sub, sub_off, threads=8)
    assert np.array_equal(m[sample], em) and np.array_equal(s[sample], es) and np.array_equal(e[sample], ee)
    # matches()/containedIn() agree with find() where they must
    mc, _, _ = pat.match_batch(1, data, offsets)
    assert np.array_equal(mc, m)


def utf16_lines(n, line_chars, seed, alphabet):
    rng = np.random.default_rng(seed)
    alpha = np.array([ord(ch) for ch in alphabet], dtype=np.uint16)
    chars = alpha[rng.integers(0, len(alpha), size=n * line_chars)]
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_chars)
    return chars, offsets


@pytest.mark.parametrize("line_chars", [8, 16, 32, 64, 128, 24, 5])
def test_utf16_hi_byte_mode_fixed_lines(line_chars):
    # class depends on the high byte only: [U+0600-06FF]+ (kCmHi)
    chars, offsets = utf16_lines(6000, line_chars, line_chars, "abc xyz؀؁ۿ܀Ԁ✓")
    assert_batch_equal(workloads.REGEX["c5"], 0, chars.view(np.uint8), offsets, cw=2)


@pytest.mark.parametrize("line_chars", [8, 32, 64, 21])
def test_utf16_mixed_page_mode(line_chars):
    # ASCII patterns over UTF-16 text: page 0 is mixed, every other page is one class (kCmMixed)
    chars, offsets = utf16_lines(6000, line_chars, 100 + line_chars, "0123456789-ab @.εΩд中İ1ı")
    for regex in (workloads.REGEX["c2"], workloads.REGEX["c3"], "[0-9]+", "a.c", "b+@", r"\d+-\d+"):
        assert_batch_equal(regex, 0, chars.view(np.uint8), offsets, cw=2)


def test_utf16_ragged_and_unsupported_class_maps():
    rng = np.random.default_rng(77)
    alphabet = "0123456789-ab @.εΩλд中؀ۿ"
    strings = ["".join(alphabet[j] for j in rng.integers(0, len(alphabet), int(rng.integers(0, 90)))) for _ in range(5000)]
    data, offsets, cw = nb.pack_haystacks(strings, char_width=2)
    for regex in (workloads.REGEX["c5"], workloads.REGEX["c3"], "[0-9]+", "ε|λ", "[a-bα-ω]+", "[؀-ۿ]+|[0-9]+", "a*"):
        assert_batch_equal(regex, 0, data, offsets, cw)
    for regex, fl in ((r"\w+", nb.UNICODE_CHARACTER_CLASS), ("[Γ-Θ]+", nb.CASE_INSENSITIVE | nb.UNICODE_CASE)):
        assert_batch_equal(regex, fl, data, offsets, cw)


FIND_ALL_CASES = [
    (workloads.REGEX["c2"], b"0123456789-- x", 1),
    (workloads.REGEX["c3"], b"abc019._%+-@@ ,;", 1),
    (r"[0-9]+", b"0123456789ab ", 1),
    (r"a*", b"aab", 1),             # empty matches: reported once, then the haystack's list ends
    (r"a|b*", b"ab c", 1),
    (r"(ab|a|b-)+", b"ab- ", 1),
    (workloads.REGEX["c5"], None, 2),
]


@pytest.mark.parametrize("regex,alphabet,cw", FIND_ALL_CASES, ids=[c[0][:16] for c in FIND_ALL_CASES])
def test_find_all_batch_matches_the_iterated_find_of_the_oracle(regex, alphabet, cw):
    """ndl_find_all_batch: `while (m.find())` per haystack, CSR output, against the oracle's find(from) loop."""
    rng = np.random.default_rng(len(regex))
    n = 3000
    lens = rng.integers(0, 90, size=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    if cw == 1:
        a = np.frombuffer(alphabet, dtype=np.uint8)
        data = a[rng.integers(0, len(a), size=total)].copy()
    else:
        chars = rng.integers(0x5F0, 0x710, size=total).astype(np.uint16)
        chars[rng.random(total) < 0.4] = 0x20
        data = chars.view(np.uint8)
    pat, ora = pair(regex)
    got = pat.find_all_batch(data, offsets, cw)
    exp = ora.find_all_batch(data, offsets, cw, threads=8)
    for name, g, e in zip(("counts", "match_offsets", "starts", "ends"), got, exp):
        assert np.array_equal(g, e), (regex, name, g[:10], e[:10])
    assert int(got[0].sum()) > 0
    # the per-string loop of the Python mirror agrees on a few haystacks
    for i in range(0, 60, 7):
        o0, o1 = int(offsets[i]), int(offsets[i + 1])
        hay = bytes(data[o0 * cw:o1 * cw]).decode("latin-1" if cw == 1 else "utf-16-le")
        k0, k1 = int(got[1][i]), int(got[1][i + 1])
        assert list(nb.iter_find(pat, hay)) == list(zip(got[2][k0:k1].tolist(), got[3][k0:k1].tolist())) == ora.find_iter(hay)


def test_find_all_batch_count_only_and_short_capacity():
    pat, ora = pair(r"[0-9]+")
    data = np.frombuffer(b"1 22 333 4444" + b"no digits" + b"7", dtype=np.uint8)
    offsets = np.array([0, 13, 22, 23], dtype=np.uint64)
    counts, moff, starts, ends = pat.find_all_batch(data, offsets, 1)
    assert counts.tolist() == [4, 0, 1] and moff.tolist() == [0, 4, 4, 5]
    assert list(zip(starts.tolist(), ends.tolist())) == [(0, 1), (2, 4), (5, 8), (9, 13), (0, 1)]


@pytest.mark.parametrize("line_chars", [64, 16, 256, 40, 48, 80, 96, 112, 11, 0])
def test_match_lines_equals_match_batch_with_computed_offsets(line_chars):
    """ndl_match_lines (fixed-length records, no offsets array) against ndl_match_batch and the oracle."""
    rng = np.random.default_rng(line_chars)
    n = 5000
    alpha = np.frombuffer(b"0123456789-ab @.", dtype=np.uint8)
    data = alpha[rng.integers(0, len(alpha), size=max(1, n * line_chars))].copy()
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_chars)
    for regex, cw in ((workloads.REGEX["c2"], 1), (workloads.REGEX["c3"], 1), (r"[0-9]+", 1), (workloads.REGEX["c2"], 2)):
        if cw == 2 and line_chars % 2:
            continue
        pat, ora = pair(regex)
        lc = line_chars // cw
        offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(lc)
        for mode in (0, 1, 2):
            got = pat.match_lines(mode, data, n, lc, cw)
            exp = ora.match_batch(mode, data, offs, cw, threads=8)
            for name, g, e in zip(("matched", "start", "end"), got, exp):
                if e is not None:
                    assert np.array_equal(g, e), (regex, cw, mode, name, line_chars)
    # device pointers
    torch = pytest.importorskip("torch")
    pat, ora = pair(workloads.REGEX["c2"])
    d = torch.from_numpy(data).cuda()
    m = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s_ = torch.zeros(n, dtype=torch.int32, device="cuda")
    e_ = torch.zeros(n, dtype=torch.int32, device="cuda")
    pat.match_lines_ptrs(2, d.data_ptr(), n, line_chars, 1, m.data_ptr(), s_.data_ptr(), e_.data_ptr())
    torch.cuda.synchronize()
    em, es, ee = ora.match_batch(2, data, offsets, 1, threads=8)
    assert np.array_equal(m.cpu().numpy(), em) and np.array_equal(s_.cpu().numpy(), es) and np.array_equal(e_.cpu().numpy(), ee)


@pytest.mark.parametrize("shape", ["fixed512", "fixed1000", "fixed4096", "ragged_long", "mixed", "huge"])
def test_long_lines_are_streamed(shape):
    """Lines longer than a tile buffer holds eight of (down to lines of many KB): one line per lane, streamed 64 bytes at a time."""
    rng = np.random.default_rng(len(shape))
    if shape.startswith("fixed"):
        L = int(shape[5:])
        n = 700
        lens = np.full(n, L)
    elif shape == "ragged_long":
        n = 900
        lens = rng.integers(200, 3000, size=n)
    elif shape == "mixed":
        n = 3000
        lens = np.where(rng.random(n) < 0.2, rng.integers(300, 6000, size=n), rng.integers(0, 100, size=n))
    else:
        n = 40
        lens = rng.integers(50_000, 200_000, size=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    words = [b"Sherlock", b"Street", b"Holmes and Watson ", b" 123-45-6789 ", b"bob@example.org", b" ", b"x", b"0", b"-", b"\n", b"abababababc"]
    blob = b"".join(words[j] for j in rng.integers(0, len(words), total // 6 + 16))
    data = np.frombuffer(blob[:total], dtype=np.uint8).copy()
    # sparse variant of the same batch, so that most lines are walked to their very end
    sparse = data.copy()
    sparse[rng.random(total) < 0.97] = ord("q")
    frm = (rng.random(n) * (lens + 1)).astype(np.int32)
    for regex in (workloads.REGEX["c2"], workloads.REGEX["c3"], workloads.REGEX["c4"], "Sherlock|Street", "[Ss]herlock",
                  "Holmes.{1,10}Watson|Watson.{1,10}Holmes", "[0-9]+x", "q*"):
        for d in (data, sparse):
            assert_batch_equal(regex, 0, d, offsets)
        assert_batch_equal(regex, 0, data, offsets, from_=frm, modes=(2,))
    if shape in ("fixed512", "ragged_long", "mixed"):
        wide = data[:total - total % 2].astype(np.uint16)
        o16 = offsets.copy()
        o16[-1] = min(int(o16[-1]), len(wide))
        for regex in (workloads.REGEX["c2"], workloads.REGEX["c3"], "Sherlock|Street"):
            assert_batch_equal(regex, 0, wide.view(np.uint8), o16, cw=2)


@pytest.mark.parametrize("shape", ["loglike", "u0_400", "u100_300", "mixed_long", "u96_97"])
def test_ragged_longer_lines_take_the_sorted_streaming_walk(shape):
    """Ragged batches whose mean line length is 96 bytes or more (log lines): windows of 128 lines sorted by length in registers,
    batches of 32 similar lines streamed; reverse passes queued.  Against the oracle, all modes, byte and UTF-16 haystacks."""
    rng = np.random.default_rng(len(shape) + 11)
    n = 4096 * 3 + 77
    if shape == "loglike":
        lens = np.clip(rng.normal(150, 60, size=n), 20, 400).astype(np.int64)
    elif shape == "u0_400":
        lens = rng.integers(0, 401, size=n)
    elif shape == "u100_300":
        lens = rng.integers(100, 301, size=n)
    elif shape == "u96_97":
        lens = rng.integers(96, 98, size=n)
    else:
        lens = np.where(rng.random(n) < 0.03, rng.integers(2000, 20000, size=n), rng.integers(30, 260, size=n))
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    words = [b"Sherlock", b"Street", b"Holmes and Watson ", b" 123-45-6789 ", b"bob@example.org", b" ", b"x", b"0", b"-", b"\n", b"abababababc",
             b"the quick brown fox ", b"GET /index.html 200 "]
    blob = b"".join(words[j] for j in rng.integers(0, len(words), total // 5 + 16))
    data = np.frombuffer(blob[:total], dtype=np.uint8).copy()
    sparse = data.copy()
    sparse[rng.random(total) < 0.97] = ord("q")
    for regex in (workloads.REGEX["c2"], workloads.REGEX["c3"], workloads.REGEX["c4"], "Sherlock|Street", "[0-9]+x", "q*",
                  "Holmes.{1,10}Watson|Watson.{1,10}Holmes"):
        for d in (data, sparse):
            assert_batch_equal(regex, 0, d, offsets)
    # a sub-batch that starts in the middle, an unaligned base
    assert_batch_equal(workloads.REGEX["c3"], 0, data, offsets[1000:9000])
    assert_batch_equal(workloads.REGEX["c2"], 0, data[3:], (offsets[5:8000] - np.uint64(3)))
    if shape in ("loglike", "u100_300"):
        wide = data[:total - total % 2].astype(np.uint16)
        o16 = offsets.copy()
        o16[-1] = min(int(o16[-1]), len(wide))
        for regex in (workloads.REGEX["c2"], workloads.REGEX["c3"], "Sherlock|Street", workloads.REGEX["c5"]):
            assert_batch_equal(regex, 0, wide.view(np.uint8), o16, cw=2)


def fast_path(pat, mode, cw):
    L = _lib.lib()
    L.ndl_debug_fast_path.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int, __import__("ctypes").c_int]
    v = L.ndl_debug_fast_path(pat._h, mode, cw)
    if v < 0:
        return None
    return {"char_mode": v & 0xFF, "replicated": (v >> 8) & 0xFF, "has_bwd": (v >> 16) & 1, "n_cols": v >> 24}


def test_expected_kernels_are_selected():
    """Guards against silently falling back to a slower kernel on the BASELINE configs.  char_mode >= 16 are the
    SWAR modes of linesq_kernel: 16 | 16-bit entries << 5 | (K == 4) << 3 | high-byte << 2 | planes."""
    fp = fast_path(pair(workloads.REGEX["c2"])[0], 2, 1)
    assert fp == {"char_mode": 16 | 8 | 2, "replicated": 16, "has_bwd": 0, "n_cols": 3}  # 4 chars per lookup, 2 compare planes, 16 table copies
    fp = fast_path(pair(workloads.REGEX["c3"])[0], 2, 1)
    assert fp["char_mode"] == 0 and fp["replicated"] == 32 and fp["has_bwd"] == 1  # 15 class changes: class map in shared memory
    fp = fast_path(pair(workloads.REGEX["c4"])[0], 2, 1)
    assert fp == {"char_mode": 16 | 32 | 8 | 2, "replicated": 1, "has_bwd": 0, "n_cols": 4}  # 258 rows: 4 chars per lookup, one copy of 16-bit entries
    fp = fast_path(pair("a[ab]{7}c|b[ab]{4}d")[0], 2, 1)
    assert fp["char_mode"] == 16 | 32 | 3 and fp["replicated"] == 8  # 289 rows x 25 columns: 16-bit entries, 3 planes
    fp = fast_path(pair("a[ab]{8}c|b[ab]{6}d")[0], 2, 1)
    assert fp is None or fp["replicated"] in (1, 32)
    fp = fast_path(pair(workloads.REGEX["c5"])[0], 2, 2)
    assert fp == {"char_mode": 16 | 8 | 4 | 1, "replicated": 16, "has_bwd": 1, "n_cols": 2}  # class from the high byte, 1 plane
    fp = fast_path(pair(workloads.REGEX["c2"])[0], 2, 2)
    assert fp == {"char_mode": 64 | 16 | 8 | 2, "replicated": 16, "has_bwd": 0, "n_cols": 3}  # ASCII pattern over UTF-16: 16-bit lanes
    fp = fast_path(pair(workloads.REGEX["c3"])[0], 2, 2)
    assert fp["char_mode"] == 2  # e-mail regex over UTF-16: no compare plan, one mixed page (lines8)
    fp = fast_path(pair("Sherlock|Holmes|Watson|Irene|Adler|John|Baker")[0], 1, 1)
    assert fp["char_mode"] == 4 and fp["replicated"] == 1  # keyword list, containedIn: one plain copy of the pair table in 16-bit entries
    assert fast_path(pair("[a-bα-ω]+")[0], 2, 2)["char_mode"] & 64  # two mixed pages: no lines8 mode, but two ranges on 16-bit lanes
    assert fast_path(pair("[a-bα-ωа-я一-龥]+@")[0], 2, 2) is None  # four mixed pages, five ranges: generic kernel
    fp = fast_path(pair("Holmes.{1,10}Watson|Watson.{1,10}Holmes")[0], 2, 1)
    assert fp["char_mode"] == 3 and fp["replicated"] == 1 and fp["has_bwd"] == 1  # 309 states x 12 classes: one plain stride-1 table


SWAR_CASES = [
    (workloads.REGEX["c2"], b"0123456789--- /:,."),
    (workloads.REGEX["c4"], b"aaabbbc`d"),
    (r"[0-9]+", b"0123456789/: "),
    (r"[^a]+b", b"ab`c\x7f\x80"),
    (r"[a-c]z|[b-d]z", b"abcdz`ey"),
    (r"\d+-\d+", b"0123456789-,."),
    ("x[\x01-\x1f]y", b"xy\x00\x1f\x20\x01"),
    ("[\x7f]+a", b"a\x7f\x7e\x80\xff"),
    (r"(ab|a|b-)+", b"ab-,.`c"),
    (r"a+b+", b"ab`c"),
    (r"a[ab]{7}c|b[ab]{4}d", b"aaabbbcd`e"),
    (r"a[ab]{5}c", b"aaabbbc`d"),      # 4 chars per lookup, 4 copies of 16-bit entries
    (r"a[ab]{3}c|b[ab]{2}d", b"aaabbbcd`e"),
]


@pytest.mark.parametrize("regex,hot", SWAR_CASES, ids=[c[0][:16] for c in SWAR_CASES])
@pytest.mark.parametrize("line_len", [64, 16, 256, 40, 0])
def test_swar_modes_all_byte_values(regex, hot, line_len):
    """linesq_kernel (packed-compare classifier): every byte value, around every range bound, fixed / ragged."""
    rng = np.random.default_rng(line_len + len(regex))
    n = 4000
    if line_len:
        lens = np.full(n, line_len)
    else:
        lens = rng.integers(0, 150, size=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    data = rng.integers(0, 256, size=total, dtype=np.uint8)
    h = np.frombuffer(hot, dtype=np.uint8)
    pick = rng.random(total) < 0.75
    data[pick] = h[rng.integers(0, len(h), size=int(pick.sum()))]
    assert fast_path(pair(regex)[0], 2, 1)["char_mode"] >= 16
    assert_batch_equal(regex, 0, data, offsets)


@pytest.mark.parametrize("line_chars", [32, 8, 128, 21, 0])
def test_swar_utf16_16bit_lanes(line_chars):
    """ASCII / BMP-range patterns over UTF-16 text (what a java.lang.String is): compares on 16-bit lanes."""
    rng = np.random.default_rng(40 + line_chars)
    n = 4000
    lens = np.full(n, line_chars) if line_chars else rng.integers(0, 70, size=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    hot = np.array([ord(ch) for ch in "0123456789--- /:,.abc`dx\u0130\u0660\u03b1\u03c9\u03ca\u0410\u042f\u0430"], dtype=np.uint16)
    chars = rng.integers(0, 0x10000, size=total).astype(np.uint16)
    pick = rng.random(total) < 0.75
    chars[pick] = hot[rng.integers(0, len(hot), size=int(pick.sum()))]
    chars[::53] = 0xFFFF
    chars[7::61] = 0x8000
    chars[3::67] = 0x7FFF
    for regex in (workloads.REGEX["c2"], workloads.REGEX["c4"], "[0-9]+x", "[\u03b1-\u03c9]+", "[a-c\u0410-\u042f]x", r"\d+-\d+"):
        assert fast_path(pair(regex)[0], 2, 2)["char_mode"] & 64, regex
        assert_batch_equal(regex, 0, chars.view(np.uint8), offsets, cw=2)


def test_utf16_high_byte_mode_with_from_offsets_and_long_lines():
    """C5's regex (class from the high byte, table-driven reverse pass) with find(from, to) and with streamed long lines."""
    rng = np.random.default_rng(77)
    for lens in (rng.integers(0, 60, size=3000), rng.integers(150, 2500, size=300), np.full(500, 512)):
        n = len(lens)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(lens)
        total = int(offsets[-1])
        chars = rng.integers(0x20, 0x400, size=total).astype(np.uint16)
        hot = rng.random(total) < 0.3
        chars[hot] = rng.integers(0x600, 0x700, size=int(hot.sum()))
        frm = (rng.random(n) * (lens + 1)).astype(np.int32)
        assert_batch_equal(workloads.REGEX["c5"], 0, chars.view(np.uint8), offsets, cw=2)
        assert_batch_equal(workloads.REGEX["c5"], 0, chars.view(np.uint8), offsets, cw=2, from_=frm, modes=(2,))


def test_swar_utf16_all_high_bytes():
    rng = np.random.default_rng(12)
    for line_chars in (32, 8, 128, 19):
        n = 5000
        chars = rng.integers(0, 0x10000, size=n * line_chars).astype(np.uint16)
        hotm = rng.random(len(chars)) < 0.6
        chars[hotm] = rng.integers(0x5F0, 0x710, size=int(hotm.sum()))
        chars[::101] = 0xFFFF
        offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_chars)
        assert_batch_equal(workloads.REGEX["c5"], 0, chars.view(np.uint8), offsets, cw=2)


def _tiled_on_device(data_b, off_b, reps):
    """A block of lines repeated `reps` times in device memory (the bytes and an offsets array that keeps counting)."""
    import torch
    n_b, block_chars = len(off_b) - 1, int(off_b[-1])
    data_d = torch.from_numpy(np.ascontiguousarray(data_b).view(np.uint8)).cuda().repeat(reps)
    off = torch.from_numpy(off_b.astype(np.int64)).cuda()
    offs = (off[:-1].unsqueeze(0) + (torch.arange(reps, device="cuda", dtype=torch.int64) * block_chars).unsqueeze(1)).reshape(-1)
    offs = torch.cat([offs, torch.tensor([reps * block_chars], device="cuda", dtype=torch.int64)])
    return data_d, offs, n_b


@pytest.mark.parametrize("key,n_block,reps,cw", [("c3", 1_000_000, 100, 1), ("c5", 500_000, 100, 2)])
def test_full_size_block_repetition(key, n_block, reps, cw):
    """BASELINE configs 3 and 5 at their full sizes (100 M ragged lines / 50 M UTF-16 lines): a block of lines the oracle can do
    in seconds, repeated in device memory.  Every repetition must give the block's results (the blocks sit at different
    alignments and tile positions), and the first block must equal the oracle."""
    import torch
    gen = workloads.c3_lines if key == "c3" else workloads.c5_lines
    data_b, off_b = gen(n_block)
    pat, ora = pair(workloads.REGEX[key])
    data_d, offs, n_b = _tiled_on_device(data_b, off_b, reps)
    n = n_b * reps
    m = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s = torch.zeros(n, dtype=torch.int32, device="cuda")
    e = torch.zeros(n, dtype=torch.int32, device="cuda")
    for mode in (2, 1):
        m.fill_(7), s.fill_(-7), e.fill_(-7)
        pat.match_batch_ptrs(mode, data_d.data_ptr(), offs.data_ptr(), n, cw, m.data_ptr(), s.data_ptr(), e.data_ptr())
        torch.cuda.synchronize()
        em, es, ee = ora.match_batch(mode, data_b, off_b, cw, threads=8)
        assert np.array_equal(m[:n_b].cpu().numpy(), em)
        assert bool((m.view(reps, n_b) == m[:n_b]).all())
        if mode == 2:
            assert np.array_equal(s[:n_b].cpu().numpy(), es) and np.array_equal(e[:n_b].cpu().numpy(), ee)
            assert bool((s.view(reps, n_b) == s[:n_b]).all()) and bool((e.view(reps, n_b) == e[:n_b]).all())
