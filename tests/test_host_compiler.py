"""Host pipeline (regex -> NFA -> 4 DFAs -> tables) against the reference's own fixtures.  CPU only.

* the 12 snapshot class files (SnapshotTests.java:30-57, decoded by tools/make_golden.py): BYTE_CLASSES,
  the four STATES_* tables, their byte/short width, the accepting states - must be IDENTICAL
* NFAToDFACompilerTest.java:11-35: state counts before minimisation
* DFATest.java:186-352: byte class assertions
* error surface of RegexParser / DFACompiler
"""
import ctypes
import json
import os

import numpy as np
import pytest

import needle_b200 as nb
from needle_b200 import _lib
from needle_b200.blob import TABLE_NAMES, parse_blob


@pytest.fixture(scope="module")
def snapshots(golden_dir):
    with open(os.path.join(golden_dir, "snapshots.json"), encoding="utf-8") as f:
        return json.load(f)


def debug_dfa(regex, mode, flags=0, want_classes=False):
    L = _lib.lib()
    L.ndl_debug_dfa.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
                                ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p]
    u = regex.encode("utf-16-le")
    buf = ctypes.create_string_buffer(u, len(u))
    raw, mini = ctypes.c_int32(), ctypes.c_int32()
    classes = np.zeros(65536, dtype=np.uint8) if want_classes else None
    rc = L.ndl_debug_dfa(ctypes.cast(buf, ctypes.c_void_p), len(u) // 2, flags, mode, ctypes.byref(raw), ctypes.byref(mini),
                         classes.ctypes.data if want_classes else None)
    assert rc == 0, _lib.last_error()
    return raw.value, mini.value, classes


SNAPSHOT_NAMES = ["DigitPlus", "HolmesNearWatson", "RepeatingUnionOfShortStrings", "Sherlock",
                  "SherlockInitialCharCaseInsensitive", "SherlockStreet", "SingleCharacterUnicode", "Suffix",
                  "TwoNamesCaseInsensitiveFirstChar", "UnicodeUnion", "UnionOfManyNames", "aDotc"]


@pytest.mark.parametrize("name", SNAPSHOT_NAMES)
def test_tables_identical_to_reference_snapshot(snapshots, name):
    snap = snapshots[name]
    b = parse_blob(nb.compile_to_bytes(snap["regex"], snap["flags"]))
    # BYTE_CLASSES: everything 0 except the fillBytes runs; index 65536 (catch-all) is not addressable by a char
    expected = np.zeros(65537, dtype=np.int64)
    for cls, lo, hi in snap["byte_class_runs"]:
        expected[lo:hi + 1] = cls
    assert np.array_equal(expected[:65536], b.class_map.astype(np.int64))
    assert expected[65535] == 0
    for k, tn in enumerate(TABLE_NAMES):
        ref, mine = snap["tables"][tn], b.tables[k]
        assert ref["stride"] == b.stride
        assert ref["n_states"] == mine.n_states
        assert ref["width"] == mine.width
        assert ref["entries"] == mine.entries.reshape(-1).tolist()
        assert ref["accepting"] == np.nonzero(mine.accepting)[0].tolist()
    # the generated class has an indexBackwards method iff the pattern can have more than one length
    assert snap["has_index_backwards"] == (b.reverse_mode != 2)
    if b.reverse_mode == 1:  # single-char reverse scan: the char and Integer.MAX_VALUE appear in indexBackwards
        assert b.reverse_char in snap["index_backwards_int_constants"]
        assert 2147483647 in snap["index_backwards_int_constants"]
    if b.reverse_mode == 0 and b.tables[3].max_char < 0xFFFF:
        assert b.tables[3].max_char in snap["index_backwards_int_constants"]
    # matches() compares each char with DFA.maxChar() of the matching DFA
    assert b.tables[0].max_char in snap["matches_int_constants"]


def test_state_counts_before_minimisation():
    # NFAToDFACompilerTest.java:11-35
    assert debug_dfa("(AB){1,2}", 0)[0] == 7
    assert debug_dfa("(AB){1,2}", 1)[0] == 3
    assert debug_dfa("(AB){1,2}", 2)[0] == 6


def test_byte_vs_short_tables():
    # DFACompilerTest.java:575-589: same regex, byte-sized matching DFA, short-sized search DFA
    b = parse_blob(nb.compile_to_bytes(".{0,43}A", 0))
    assert b.tables[0].n_states <= 127 and b.tables[0].width == 1
    assert b.tables[2].n_states > 127 and b.tables[2].width == 2


def test_byte_classes_literal():
    # DFATest.java:186-198
    bc = debug_dfa("abc", 0, want_classes=True)[2]
    assert not bc[:ord("a")].any()
    assert (bc[ord("a")], bc[ord("b")], bc[ord("c")]) == (1, 2, 3)
    assert not bc[ord("d"):65536].any()


def test_byte_classes_two_ranges_then_literal():
    # DFATest.java:201-217
    bc = debug_dfa("[A-Za-z]+ab", 0, want_classes=True)[2]
    assert not bc[:ord("A")].any()
    assert bc[ord("A")] == 1 and bc[ord("Z")] == 1 and bc[ord("a")] == 2 and bc[ord("b")] == 3
    assert (bc[ord("c"):ord("z") + 1] == 1).all()
    assert not bc[ord("z") + 1:65535].any()


def test_byte_classes_with_dot():
    # DFATest.java:220-236 (DOTALL)
    bc = debug_dfa("[A-Za-z]+.b", 0, flags=nb.DOTALL, want_classes=True)[2]
    assert (bc[:ord("A")] == 1).all()
    assert bc[ord("A")] == 2 and bc[ord("Z")] == 2 and bc[ord("a")] == 2 and bc[ord("b")] == 3
    assert (bc[ord("c"):ord("z") + 1] == 2).all()
    assert (bc[ord("z") + 1:65535] == 1).all()


def test_byte_classes_url():
    # DFATest.java:239-250
    bc = debug_dfa("http://.+", 0, flags=nb.DOTALL, want_classes=True)[2]
    assert (bc[:ord("/")] == 1).all() and bc[ord("/")] == 2
    assert (bc[ord("0"):ord(":")] == 1).all() and bc[ord(":")] == 3


def test_byte_classes_more():
    # DFATest.java:296-305, 331-347
    bc = debug_dfa("h:.+", 0, want_classes=True)[2]
    assert (bc[0], bc[ord(":")], bc[ord(";")], bc[ord("h")], bc[ord("i")]) == (1, 2, 1, 3, 1)
    bc = debug_dfa("Hol.{0,2}Wat|Wat.{0,2}Hol", 0, want_classes=True)[2]
    got = [bc[ord(c)] for c in "\0HIWXablmopt"]
    assert got == [1, 2, 1, 3, 1, 4, 1, 5, 1, 6, 1, 7]


def test_lengths_and_reverse_mode():
    info = nb.compile_to_bytes  # noqa
    b = parse_blob(nb.compile_to_bytes(r"\d{3}-\d{2}-\d{4}", 0))
    assert (b.min_length, b.max_length, b.reverse_mode) == (11, 11, 2)
    b = parse_blob(nb.compile_to_bytes("http://.+", 0))
    assert (b.min_length, b.max_length, b.reverse_mode) == (8, -1, 0)
    b = parse_blob(nb.compile_to_bytes("[A-Za-z0-9._%+-]+@[A-Za-z0-9.-]+", 0))
    assert (b.min_length, b.max_length) == (3, -1)
    b = parse_blob(nb.compile_to_bytes("a[ab]{7}c", 0))
    assert (b.min_length, b.max_length, b.reverse_mode) == (9, 9, 2)
    assert b.tables[2].n_states >= 256 and b.tables[2].width == 2  # the "256-state DFA" of BASELINE config 4


# RegexParserMalformedRegexTest.java / RegexParser.java:113-116, 381-392, 427-429, 516-518 / readme.md:80-88
MALFORMED = ["(", ")", "a)", "(a", "[", "[a", "[a-", "a{", "a{1", "a{1,", "a{2,1}", "{1}", "*", "+", "?", "a|", "|", "^a", "a$",
             r"\b", r"\B", r"\A", r"\z", r"\Z", r"\G", r"\p{L}", r"\1", r"\k", "a*?", "a+?", "a??", "a{1}?", "a*+", "a++",
             "a?+", "a{1,2}+", r"\xa0", r"\x0", "\\", "[b-a]", r"\c"]


@pytest.mark.parametrize("regex", MALFORMED)
def test_malformed_regexes_are_rejected(regex):
    with pytest.raises(nb.PatternClassCompilationException):
        nb.compile_to_bytes(regex, 0)


def test_syntax_error_is_the_cause():
    # DFACompiler.compileToBytes wraps the parser's PatternSyntaxException (DFACompiler.java:71-73)
    with pytest.raises(nb.PatternClassCompilationException) as ei:
        nb.compile_to_bytes("a**?", 0)
    assert isinstance(ei.value.__cause__, nb.PatternSyntaxException)


def test_unknown_flags_rejected():
    with pytest.raises(ValueError):  # IllegalArgumentException, CompilerOptions.java:10-12
        nb.compile_to_bytes("a", 0x4)


WELL_FORMED = ["", "()", "()|abc", "a{0,0}", "[]]", "[-]", "[a-]", "[[a-c]]", r"\x41", r"\0101", r"\$\{[^}]*}", "(?:abc)+",
               "(?<name>abc)+", r"\h\H\v\V\s\S\w\W\d\D", r"[\[\]\\]", "a{3}", "(a|b)*c{2,3}"]


@pytest.mark.parametrize("regex", WELL_FORMED)
def test_well_formed_regexes_compile(regex):
    for flags in (0, nb.DOTALL, nb.CASE_INSENSITIVE, nb.LEFTMOST_LONGEST):
        assert len(nb.compile_to_bytes(regex, flags)) > 44


def test_blob_roundtrip_and_corruption():
    blob = nb.compile_to_bytes("Sherlock|Street", 0)
    info = _lib.BlobInfo()
    assert _lib.lib().ndl_blob_info_get(blob, len(blob), ctypes.byref(info)) == 0
    assert info.stride == 16 and list(info.n_states) == [13, 13, 13, 13] and info.reverse_mode == 1
    for bad in (blob[:-1], blob[:40], b"XXXX" + blob[4:], blob[:100] + bytes([blob[100] ^ 1]) + blob[101:]):
        assert _lib.lib().ndl_blob_info_get(bad, len(bad), ctypes.byref(info)) == _lib.NDL_EBLOB
