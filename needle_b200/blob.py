"""Reader for the pattern table blob (layout: needle_b200/csrc/host/blob.cpp).

Host-side convenience only (tests, tooling, `Pattern.info`); the product path hands the raw bytes to
`ndl_pattern_create` and never interprets them in Python.
"""
import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np

MAGIC = 0x424C444E
TABLE_NAMES = ("Matches", "ContainedIn", "Forwards", "Backwards")


@dataclass
class Table:
    n_states: int
    width: int
    max_char: int
    accepting: np.ndarray  # uint8[n_states]
    entries: np.ndarray    # int16[n_states, stride]


@dataclass
class Blob:
    version: int
    flags: int
    min_length: int
    max_length: int
    stride: int
    byte_class_count: int
    reverse_mode: int
    reverse_char: int
    class_map: np.ndarray  # uint16[65536]
    tables: List[Table] = field(default_factory=list)


def _fnv1a(b: bytes) -> int:
    h = 2166136261
    for x in b:
        h = ((h ^ x) * 16777619) & 0xFFFFFFFF
    return h


def parse_blob(b: bytes) -> Blob:
    magic, version, flags, mn, mx, stride, bcc, rmode, rchar, n_runs = struct.unpack_from("<Iiiiiiiiii", b, 0)[:10]
    if magic != MAGIC:
        raise ValueError("bad magic")
    if struct.unpack_from("<I", b, len(b) - 4)[0] != _fnv1a(b[:-4]):
        raise ValueError("checksum mismatch")
    pos = 40
    class_map = np.zeros(65536, dtype=np.uint16)
    c = 0
    for _ in range(n_runs):
        last, cls = struct.unpack_from("<HH", b, pos)
        pos += 4
        class_map[c:last + 1] = cls
        c = last + 1
    out = Blob(version, flags, mn, mx, stride, bcc, rmode, rchar, class_map)
    for _ in range(4):
        n_states, width, max_char = struct.unpack_from("<iii", b, pos)
        pos += 12
        acc = np.frombuffer(b, dtype=np.uint8, count=n_states, offset=pos).copy()
        pos = (pos + n_states + 3) & ~3
        ent = np.frombuffer(b, dtype="<i2", count=n_states * stride, offset=pos).reshape(n_states, stride).copy()
        pos = (pos + 2 * n_states * stride + 3) & ~3
        out.tables.append(Table(n_states, width, max_char, acc, ent))
    return out
