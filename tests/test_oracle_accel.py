"""The accelerated indexForwards of the CPU baseline (oracle/needle_oracle.c ndlo_index_forwards_accel) and the host
compiler's restatement of Factorization / CompilationPolicy that feeds it (csrc/host/factorization.cpp).

Pins: (1) the PREFIX / SUFFIX / INFIX / FIRST_BYTE_MASK constants of the reference's 12 snapshot class files;
(2) accelerated == plain on every golden row, the inline KATs, generated haystacks, and the C1-C5 bench workloads.
CPU only."""
import json
import os

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200.blob import parse_blob
from tests import workloads
from tests.kats import FIND
from tests.oracle_lib import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def snapshots():
    with open(os.path.join(GOLDEN, "snapshots.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def rows():
    with open(os.path.join(GOLDEN, "matches.json")) as f:
        return json.load(f)


SNAPSHOT_NAMES = ["DigitPlus", "HolmesNearWatson", "RepeatingUnionOfShortStrings", "Sherlock", "SherlockInitialCharCaseInsensitive",
                  "SherlockStreet", "SingleCharacterUnicode", "Suffix", "TwoNamesCaseInsensitiveFirstChar", "UnicodeUnion",
                  "UnionOfManyNames", "aDotc"]


@pytest.mark.parametrize("name", SNAPSHOT_NAMES)
def test_affix_constants_identical_to_reference_snapshot(snapshots, name):
    """The generated class carries PREFIX / SUFFIX / INFIX only when the policy uses them (DFAClassBuilder.addAffixConstants),
    and FIRST_BYTE_MASK only when the mask loop is the seek that is emitted (shouldIncludeFirstByteMask)."""
    snap = snapshots[name]
    a = parse_blob(nb.compile_to_bytes(snap["regex"], snap["flags"])).accel
    assert a.present
    assert (a.prefix if a.use_prefix else None) == snap["prefix"]
    assert (a.suffix if a.use_suffix else None) == snap["suffix"]
    assert (a.infix if a.use_infixes else None) == snap["infix"]
    mask_emitted = a.has_first_byte_mask and not (a.use_prefix or a.use_suffix or a.use_infixes or a.can_seek_for_predicate)
    if snap["first_byte_mask"] is not None:
        assert mask_emitted and a.first_byte_mask == snap["first_byte_mask"]
    else:
        assert not mask_emitted or not a.first_byte_mask  # (an accepting root emits no mask loop either)


def test_policy_of_the_baseline_regexes():
    """SURVEY.md Appendix A.3, hand-traced from the reference's sources."""
    acc = {k: parse_blob(nb.compile_to_bytes(workloads.REGEX[k], 0)).accel for k in workloads.REGEX}
    c1, c2, c3, c4, c5 = (acc[k] for k in ("c1", "c2", "c3", "c4", "c5"))
    assert c1.use_prefix and c1.prefix == "http://" and c1.use_max_start and not c1.use_suffix and not c1.use_infixes
    assert not c2.use_prefix and not c2.use_suffix and c2.use_infixes and c2.infix == "-" and c2.use_max_start
    assert not (c3.use_prefix or c3.use_suffix or c3.use_infixes or c3.can_seek_for_predicate)
    assert c3.has_first_byte_mask and not c3.byte_check_first_char and not c3.use_max_start
    assert c4.use_prefix and c4.prefix == "a" and c4.use_suffix and c4.suffix == "c" and c4.use_max_start
    assert not (c5.use_prefix or c5.use_suffix or c5.use_infixes or c5.can_seek_for_predicate or c5.has_first_byte_mask)


def test_accelerated_equals_plain_on_every_golden_row(rows):
    for r in rows:
        for flags in ([r["flags"]] if r["flags"] is not None else [0, nb.LEFTMOST_LONGEST, nb.CASE_INSENSITIVE, nb.DOTALL, r["java_random_flags"]]):
            ora = Oracle(nb.compile_to_bytes(r["pattern"], flags))
            for frm in range(0, len(r["haystack"]) + 1):
                assert ora.index_forwards(r["haystack"], frm, True) == ora.index_forwards(r["haystack"], frm, False), (r, flags, frm)


def test_accelerated_equals_plain_on_inline_kats():
    for regex, flags, hay, frm, exp in FIND:
        ora = Oracle(nb.compile_to_bytes(regex, flags))
        assert ora.index_forwards(hay, frm, True) == ora.index_forwards(hay, frm, False)
        data, offsets, cw = nb.pack_haystacks([hay])
        got = ora.match_batch(2, data, offsets, cw, np.array([frm], dtype=np.int32), accelerated=True)
        assert (bool(got[0][0]), int(got[1][0]), int(got[2][0])) == exp


ACCEL_REGEXES = [
    # one per seek form: prefix, prefix + suffix, suffix, infix, predicate (range / single / case pair), first-byte mask, none
    "http://.+", "a[ab]{7}c", "a.c", "Sherlock", "[Ss]herlock", "anywhere|somewhere", r"\d{3}-\d{2}-\d{4}",
    "Holmes.{1,10}Watson|Watson.{1,10}Holmes", "[0-9]+", "[a-f]+x", "q[a-z ]*7", "[Ss]+t", "(ab|a|bcdef|g)+",
    "Sherlock|Holmes|Watson|Irene|Adler|John|Baker", "[A-Za-z0-9._%+-]+@[A-Za-z0-9.-]+", "ε|λ", "[؀-ۿ]+", "a*", "(a|b)*c", "x{0,2}yz",
    "abc|abd", "(abc){1,2}", "[ab]c[de]", "the [Cc]rown", "a{2}b", "z+", "[^a]+a",
]


@pytest.mark.parametrize("regex", ACCEL_REGEXES)
@pytest.mark.parametrize("flags", [0, nb.LEFTMOST_LONGEST])
def test_accelerated_equals_plain_on_generated_haystacks(regex, flags):
    """Haystacks over the pattern's own chars plus noise, so that every seek form finds, skips and gives up."""
    ora = Oracle(nb.compile_to_bytes(regex, flags))
    rng = np.random.default_rng(abs(hash((regex, flags))) % (2 ** 32))
    lit = [c for c in regex if c.isalnum() or c in " @-.:/"]
    alphabet = sorted(set(lit + list("ab c-7.xyz@0qS")))
    if any(ord(c) > 255 for c in regex):
        alphabet += list("ελ؀ۿ٣")
    strings = []
    for _ in range(400):
        n = int(rng.integers(0, 48))
        strings.append("".join(alphabet[j] for j in rng.integers(0, len(alphabet), n)))
    data, offsets, cw = nb.pack_haystacks(strings)
    for frm in (None, rng.integers(0, 5, size=len(strings)).astype(np.int32)):
        a = ora.match_batch(2, data, offsets, cw, frm, accelerated=True)
        b = ora.match_batch(2, data, offsets, cw, frm, accelerated=False)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("key,gen,cw", [("c2", workloads.c2_lines, 1), ("c3", workloads.c3_lines, 1), ("c4", workloads.c4_lines, 1),
                                        ("c5", workloads.c5_lines, 2), ("c2", workloads.c2_lines_utf16, 2)])
def test_accelerated_equals_plain_on_bench_workloads(key, gen, cw):
    ora = Oracle(nb.compile_to_bytes(workloads.REGEX[key], 0))
    data, offsets = gen(100_000)
    a = ora.match_batch(2, data, offsets, cw, threads=4, accelerated=True)
    b = ora.match_batch(2, data, offsets, cw, threads=4, accelerated=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[0].sum() > 0


def test_c1_strings():
    ora = Oracle(nb.compile_to_bytes(workloads.REGEX["c1"], 0))
    data, offsets, cw = nb.pack_haystacks(workloads.c1_strings())
    a = ora.match_batch(2, data, offsets, cw, accelerated=True)
    b = ora.match_batch(2, data, offsets, cw, accelerated=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert "indexOf(PREFIX)" in ora.accel_summary()
