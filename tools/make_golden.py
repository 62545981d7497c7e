#!/usr/bin/env python3
"""Build tests/golden/*.json from the reference's own test fixtures.

Run in the build container (needs /root/reference); the outputs are committed so that nothing on the
GPU box reads /root/reference.

  matches.json    the 200 rows of needle-compiler/src/test/resources/matches.txt, parsed with the rules of
                  RegexTestSpecParser.java:31-51, 94-141, plus - for rows without a flags column - the exact
                  flags DFACompilerTest.generateFlags draws (`ALL_FLAGS & new Random(1024).nextInt(ALL_FLAGS)`,
                  DFACompilerTest.java:26, 775-782; java.util.Random is restated below).
  snapshots.json  the static state of the 12 generated classes in resources/snapshots/*.class
                  (SnapshotTests.java:30-57 lists regex -> class name; all compiled with flags 0):
                  BYTE_CLASSES runs, STATES_* tables, accepting states, PREFIX/SUFFIX/INFIX, FIRST_BYTE_MASK,
                  whether indexBackwards exists, and the int constants of matches() (they contain maxChar).
  sherlock_line0.txt  line 0 of resources/sherlockholmes.txt, the haystack of the JDK differential tests
                  (DFACompilerTest.java:622-632, 671-699).
"""
import glob
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from classfile import ClassFile, int_constants, run_clinit  # noqa: E402

REF = "/root/reference/needle-compiler/src/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

ALL_FLAGS = 0x20 | 0x02 | 0x40 | 0x800000 | 0x100


class JavaRandom:
    """java.util.Random (LCG, 48-bit state)."""

    def __init__(self, seed):
        self.seed = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def next(self, bits):
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.seed >> (48 - bits)
        return v - (1 << 32) if v >= (1 << 31) else v

    def next_int(self, bound):
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):  # Java: `u - r + m < 0` after int overflow
                return r
            u = self.next(31)


def jtrim(s):
    """java.lang.String.trim(): strips chars <= U+0020 only (str.strip() would also eat U+2001 etc.)."""
    a, b = 0, len(s)
    while a < b and s[a] <= " ":
        a += 1
    while b > a and s[b - 1] <= " ":
        b -= 1
    return s[a:b]


class SpecParser:
    """RegexTestSpecParser.chomp / readSpec."""

    def __init__(self, s):
        self.s, self.idx = s, 0

    def chomp(self):
        s, start, seen, in_quote = self.s, self.idx, False, False
        while self.idx < len(s):
            c = s[self.idx]
            if c == " " and not in_quote:
                if seen:
                    return jtrim(s[start:self.idx])
            elif c == "'":
                if in_quote:
                    self.idx += 1
                    sub = jtrim(s[start:self.idx])
                    return sub[1:-1]
                in_quote = True
            else:
                seen = True
            self.idx += 1
        if not seen:
            raise ValueError("Tried to chomp but didn't see anything")
        return jtrim(s[start:self.idx])

    def optional(self):
        return self.chomp() if len(self.s) > self.idx else None


def parse_matches():
    rows = []
    rnd = JavaRandom(1024)
    with open(f"{REF}/resources/matches.txt", encoding="utf-8") as f:
        lines = f.read().split("\n")
    for lineno, raw in enumerate(lines, 1):
        if not raw.strip() or raw.startswith("#"):
            continue
        p = SpecParser(jtrim(raw))
        pattern = p.chomp()
        target = p.chomp().replace("\\n", "\n").replace("\\r", "\r")
        ok = p.chomp() == "y"
        start = end = -1
        if ok:
            start, end = int(p.chomp()), int(p.chomp())
        fl = p.optional()
        flags = int(fl, 16) if fl is not None else None
        row = {"line": lineno, "pattern": pattern, "haystack": target, "matched": ok, "start": start, "end": end,
               "flags": flags}
        if flags is None:
            row["java_random_flags"] = ALL_FLAGS & rnd.next_int(ALL_FLAGS)
        rows.append(row)
    return rows


SNAPSHOT_REGEXES = {  # SnapshotTests.java:30-57
    "Sherlock": "Sherlock",
    "SherlockStreet": "Sherlock|Street",
    "SherlockInitialCharCaseInsensitive": "[Ss]herlock",
    "UnionOfManyNames": "Sherlock|Holmes|Watson|Irene|Adler|John|Baker",
    "Suffix": "anywhere|somewhere",
    "HolmesNearWatson": "Holmes.{1,10}Watson|Watson.{1,10}Holmes",
    "TwoNamesCaseInsensitiveFirstChar": "([Ss]herlock)|([Hh]olmes)",
    "aDotc": "a.c",
    "DigitPlus": "[0-9]+",
    "SingleCharacterUnicode": "ε",
    "UnicodeUnion": "ε|λ",
    "RepeatingUnionOfShortStrings": "(ab|a|bcdef|g)+",
}

SPECS = {"Matches": "STATES_MATCHES", "ContainedIn": "STATES_CONTAINEDIN", "Forwards": "STATES_FORWARDS",
         "Backwards": "STATES_BACKWARDS"}


def decode_snapshots():
    out = {}
    for path in sorted(glob.glob(f"{REF}/resources/snapshots/*.class")):
        name = os.path.basename(path)[:-6]
        cf = ClassFile(open(path, "rb").read())
        st = run_clinit(cf)
        methods = {(m["name"], m["desc"]): m for m in cf.methods}
        bc = st["BYTE_CLASSES"]
        entry = {
            "regex": SNAPSHOT_REGEXES[name],
            "flags": 0,
            "byte_class_runs": [list(r) for r in bc["runs"]],  # (class, first, last) fillBytes calls
            "byte_classes_len": len(bc["data"]),
            "tables": {},
            "prefix": st.get("PREFIX"),
            "suffix": st.get("SUFFIX"),
            "infix": st.get("INFIX"),
            "first_byte_mask": [i for i, v in enumerate(st["FIRST_BYTE_MASK"]["data"]) if v] if "FIRST_BYTE_MASK" in st else None,
            "has_index_backwards": ("indexBackwards", "(II)I") in methods,
            "matches_int_constants": int_constants(cf, methods[("matches", "()Z")]),
        }
        if entry["has_index_backwards"]:
            entry["index_backwards_int_constants"] = int_constants(cf, methods[("indexBackwards", "(II)I")])
        for spec, field in SPECS.items():
            arr = st[field]
            stride = arr["stride"]
            acc_name = "ACCEPTED_ARRAY_" + spec
            if acc_name in st:
                accepting = [i for i, v in enumerate(st[acc_name]["data"]) if v]
            else:
                accepting = [int_constants(cf, methods[("wasAccepted" + spec, "(I)Z")])[0]]
            entry["tables"][spec] = {
                "width": 2 if arr["type"] == 9 else 1,  # T_SHORT = 9, T_BYTE = 8
                "stride": stride,
                "n_states": len(arr["data"]) // stride,
                "entries": arr["data"],
                "accepting": accepting,
                "strings": arr["strings"],
            }
        out[name] = entry
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    rows = parse_matches()
    assert len(rows) == 200, len(rows)
    with open(os.path.join(OUT, "matches.json"), "w", encoding="utf-8") as f:
        json.dump(rows, f, ensure_ascii=True, indent=0)
    snaps = decode_snapshots()
    assert len(snaps) == 12
    with open(os.path.join(OUT, "snapshots.json"), "w", encoding="utf-8") as f:
        json.dump(snaps, f, ensure_ascii=True)
    with open(f"{REF}/resources/sherlockholmes.txt", encoding="utf-8") as f:
        line0 = f.readline().rstrip("\n")
    with open(os.path.join(OUT, "sherlock_line0.txt"), "w", encoding="utf-8") as f:
        f.write(line0)
    print(f"{len(rows)} match rows, {len(snaps)} snapshots, line0 = {len(line0)} chars")


if __name__ == "__main__":
    main()
