"""Pins the CPU oracle (oracle/needle_oracle.c) - and with it the host compiler - to the reference's own
vectors.  CPU only.

1. matches.txt: all 200 rows (DFACompilerTest.fileBasedTests :701-773), each with the flags the Java
   test uses (explicit column, or the java.util.Random(1024) draw reproduced in the fixture), find() must
   give exactly (start, end); rows that are not LEFTMOST_LONGEST must also agree with a backtracking
   engine (Python `re` stands in for java.util.regex, :728-742).
2. The inline assertions of DFACompilerTest.java (tests/kats.py).
3. The search-method properties of SearchMethodTestUtil.java:48-120 over random prefixes/suffixes.
4. The tables decoded from the 12 snapshot class files, walked by the oracle with NO compiler involved,
   against Python `re` on generated haystacks.
"""
import json
import os
import random
import re

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200.blob import TABLE_NAMES
from tests import kats
from tests.oracle_lib import Oracle

_CACHE = {}


def oracle_for(regex, flags=0):
    key = (regex, flags)
    if key not in _CACHE:
        _CACHE[key] = Oracle(nb.compile_to_bytes(regex, flags))
    return _CACHE[key]


@pytest.fixture(scope="module")
def rows(golden_dir):
    with open(os.path.join(golden_dir, "matches.json"), encoding="utf-8") as f:
        return json.load(f)


def row_flags(r):
    return r["flags"] if r["flags"] is not None else r["java_random_flags"]


def test_matches_txt_all_rows(rows):
    assert len(rows) == 200
    bad = []
    for r in rows:
        got = oracle_for(r["pattern"], row_flags(r)).find(r["haystack"])
        exp = (r["matched"], r["start"], r["end"])
        if got != exp:
            bad.append((r["line"], r["pattern"], r["haystack"], hex(row_flags(r)), got, exp))
    assert not bad, bad


def test_matches_txt_rows_without_flags_hold_under_every_flag_combination(rows):
    # the Java test draws random flags for these rows, so their expectations must not depend on flags
    bits = [nb.DOTALL, nb.CASE_INSENSITIVE, nb.UNICODE_CASE, nb.UNICODE_CHARACTER_CLASS, nb.LEFTMOST_LONGEST]
    for r in rows:
        if r["flags"] is not None:
            continue
        for k in range(32):
            fl = sum(b for i, b in enumerate(bits) if k >> i & 1)
            assert oracle_for(r["pattern"], fl).find(r["haystack"]) == (r["matched"], r["start"], r["end"]), (r, hex(fl))


def test_matches_txt_find_properties(rows):
    # the `find(pattern, spec.target)` property check the Java test runs after each successful row (:744-750)
    for r in rows:
        if r["matched"]:
            check_find_properties(oracle_for(r["pattern"], row_flags(r)), r["haystack"])


# Java-isms Python `re` does not share: nested class, \S vs U+2001, `.` vs \r
PY_RE_SKIP_LINES = {162, 163, 164, 217, 229}


def test_matches_txt_agrees_with_backtracking_engine(rows):
    checked = 0
    for r in rows:
        fl = row_flags(r)
        if r["flags"] is None:
            fl &= nb.DOTALL | nb.CASE_INSENSITIVE  # flag-independent rows: keep the bits `re` can express
        if fl & (nb.LEFTMOST_LONGEST | nb.UNICODE_CASE | nb.UNICODE_CHARACTER_CLASS) or r["line"] in PY_RE_SKIP_LINES:
            continue
        pyflags = (re.DOTALL if fl & nb.DOTALL else 0) | (re.IGNORECASE if fl & nb.CASE_INSENSITIVE else 0) | re.ASCII
        pat = re.sub(r"\(\?<(\w+)>", r"(?P<\1>", r["pattern"])
        m = re.compile(pat, pyflags).search(r["haystack"])
        got = oracle_for(r["pattern"], fl).find(r["haystack"])
        assert got == ((True, m.start(), m.end()) if m else (False, -1, -1)), (r, got)
        checked += 1
    assert checked > 100


def check_find_properties(o, s, start=0):
    """SearchMethodTestUtil.find(Pattern, String, int, int) :48-96"""
    assert o.contained_in(s), s
    found, st, en = o.find(s, start)
    assert found, s                                     # property 1
    assert o.matches(s[st:en]), (s, st, en)             # property 2
    prefix = s[start:st]
    assert (o.matches("") and prefix == "") or not o.matches(prefix), (s, prefix)  # property 3
    for k in range(start, st):                          # property 5: no earlier start matches
        assert not o.matches(s[k:en]), (s, k, en)
    if o.matches(""):                                   # property 6
        assert st == start


def check_match(o, s):
    """SearchMethodTestUtil.match :110-120"""
    assert o.matches(s) and o.contained_in(s), s
    assert o.find(s) == (True, 0, len(s)), (s, o.find(s))
    check_find_properties(o, s)


@pytest.mark.parametrize("regex,flags,match,fail,contained", kats.MATCH_FAIL, ids=[k[0][:30] for k in kats.MATCH_FAIL])
def test_inline_match_fail(regex, flags, match, fail, contained):
    o = oracle_for(regex, flags)
    for s in match:
        check_match(o, s)
    for s in fail:
        assert not o.contained_in(s) and not o.matches(s), s
    for s in contained:
        assert not o.matches(s) and o.contained_in(s), s
        check_find_properties(o, s)


@pytest.mark.parametrize("regex,flags,hay,frm,exp", kats.FIND)
def test_inline_find(regex, flags, hay, frm, exp):
    assert oracle_for(regex, flags).find(hay, frm) == exp


@pytest.mark.parametrize("regex,flags,hay,exp", kats.FIND_ALL)
def test_inline_find_all(regex, flags, hay, exp):
    assert oracle_for(regex, flags).find_iter(hay) == exp


@pytest.mark.parametrize("regex,hay", kats.JDK_DIFFERENTIAL)
def test_jdk_differential(regex, hay):
    # compareResultsToStandardLibrary (DFACompilerTest.java:671-699)
    assert oracle_for(regex, 0).find_iter(hay) == [m.span() for m in re.finditer(regex, hay)]


def test_sherlock_line0(golden_dir):
    # checkMatchesInFileAgainstStandardLibrary (:622-632, 662-669): haystack = line 0 of sherlockholmes.txt
    with open(os.path.join(golden_dir, "sherlock_line0.txt"), encoding="utf-8") as f:
        line0 = f.read()
    for regex in ("Sherlock|Street", "[Ss]herlock"):
        assert oracle_for(regex, 0).find_iter(line0) == [m.span() for m in re.finditer(regex, line0)]
        assert len(oracle_for(regex, 0).find_iter(line0)) == 1


def rand_string(rng, alphabet, lo, hi):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


A_TO_Z = [chr(c) for c in range(65, 91)]          # SearchMethodTestUtil.A_THROUGH_Z
SMALL_BMP = [chr(c) for c in range(0xC5, 0xCA)]   # SearchMethodTestUtil.SMALL_BMP


@pytest.mark.parametrize("regex,flags,match,fail,contained", kats.MATCH_FAIL[:26], ids=[k[0][:30] for k in kats.MATCH_FAIL[:26]])
def test_find_with_random_prefix_suffix(regex, flags, match, fail, contained):
    # QuickTheory.qt().forAll(ALPHABET, ALPHABET).check((prefix, suffix) -> find(pattern, needle, prefix, suffix))
    o = oracle_for(regex, flags)
    rng = random.Random(hash(regex) & 0xFFFF)
    for needle in match:
        if needle == "":
            continue
        for _ in range(40):
            alpha = A_TO_Z if rng.random() < 0.5 else SMALL_BMP
            prefix, suffix = rand_string(rng, alpha, 0, 10), rand_string(rng, alpha, 0, 10)
            for s in (needle, prefix + needle, prefix + needle + suffix, needle + suffix):
                check_find_properties(o, s)


# ---- snapshot tables straight into the oracle (no regex compiler on this path)
def oracle_from_snapshot(snap, min_length, max_length, reverse_mode, reverse_char, max_char):
    cm = np.zeros(65537, dtype=np.uint16)
    for cls, lo, hi in snap["byte_class_runs"]:
        cm[lo:hi + 1] = cls
    tables, accepting = [], []
    for tn in TABLE_NAMES:
        t = snap["tables"][tn]
        tables.append(np.array(t["entries"], dtype=np.int16))
        acc = np.zeros(t["n_states"], dtype=np.uint8)
        acc[t["accepting"]] = 1
        accepting.append(acc)
    stride = snap["tables"]["Matches"]["stride"]
    return Oracle.from_tables(cm[:65536], stride, min_length, max_length, reverse_mode, reverse_char, tables, accepting, max_char)


SNAP_ALPHABETS = {
    "DigitPlus": "0129ab {", "aDotc": "abc.\nx", "SingleCharacterUnicode": "εελa ", "UnicodeUnion": "ελaκ ",
    "RepeatingUnionOfShortStrings": "abcdefgx", "Sherlock": "Sherlock s", "SherlockInitialCharCaseInsensitive": "Ssherlock ",
    "SherlockStreet": "SherlockStreet ", "Suffix": "anywhersom ", "TwoNamesCaseInsensitiveFirstChar": "SsherlockHholmes ",
    "UnionOfManyNames": "SherlockHolmesWatsonIreneAdlerJohnBaker ", "HolmesNearWatson": "HolmesWatson \n",
}


@pytest.mark.parametrize("name", sorted(SNAP_ALPHABETS))
def test_snapshot_tables_walked_by_oracle_agree_with_backtracking_engine(golden_dir, name):
    with open(os.path.join(golden_dir, "snapshots.json"), encoding="utf-8") as f:
        snap = json.load(f)[name]
    regex = snap["regex"]
    # lengths / reverse mode are not stored in the class file as data; derive them from the regex text the
    # way the reference does (Node.minLength/maxLength) via our compiler's header ONLY (tables come from the snapshot)
    from needle_b200.blob import parse_blob
    hdr = parse_blob(nb.compile_to_bytes(regex, 0))
    o = oracle_from_snapshot(snap, hdr.min_length, hdr.max_length, hdr.reverse_mode, hdr.reverse_char,
                             [t.max_char for t in hdr.tables])
    pat = re.compile(regex)
    rng = random.Random(1234)
    alphabet = SNAP_ALPHABETS[name]
    words = re.findall(r"[A-Za-z]+", regex) or [regex]
    for it in range(400):
        parts = []
        for _ in range(rng.randint(0, 6)):
            parts.append(rng.choice(words) if rng.random() < 0.4 else rand_string(rng, alphabet, 0, 6))
        s = "".join(parts)
        m = pat.search(s)
        exp = (True, m.start(), m.end()) if m else (False, -1, -1)
        assert o.find(s) == exp, (regex, s, o.find(s), exp)
        assert o.contained_in(s) == bool(m), (regex, s)
        assert o.matches(s) == bool(pat.fullmatch(s)), (regex, s)
