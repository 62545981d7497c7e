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
class Accel:
    """The accelerator record of a version-2 blob (csrc/host/pattern.h Accel; DFAClassBuilder.java:365-429)."""
    present: bool = False
    use_prefix: bool = False
    use_suffix: bool = False
    use_infixes: bool = False
    use_max_start: bool = False
    can_seek_for_predicate: bool = False
    has_first_byte_mask: bool = False
    byte_check_first_char: bool = False
    post_prefix_accepting: bool = False
    follow_accepting: bool = False
    inner_must_call_was_accepted: bool = False
    post_prefix_state: int = 0
    follow_state: int = 0
    pred_kind: int = 0
    pred_a: int = 0
    pred_b: int = 0
    prefix: str = ""
    suffix: str = ""
    infix: str = ""
    first_byte_mask: List[int] = field(default_factory=list)  # the indices (0..128) that are set


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
    accel: Accel = field(default_factory=Accel)


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
    if version >= 2:
        magic2, f, pps, fs, pk, pa, pb = struct.unpack_from("<IIiiiii", b, pos)
        if magic2 != 0x4C434341:
            raise ValueError("accelerator record missing")
        pos += 28
        strs = []
        for _ in range(3):
            (n,) = struct.unpack_from("<I", b, pos)
            pos += 4
            strs.append(b[pos:pos + 2 * n].decode("utf-16-le", "surrogatepass"))
            pos = (pos + 2 * n + 3) & ~3
        mask = [i for i in range(129) if b[pos + i]]
        out.accel = Accel(bool(f & 1024), bool(f & 1), bool(f & 2), bool(f & 4), bool(f & 8), bool(f & 16), bool(f & 32), bool(f & 64),
                          bool(f & 128), bool(f & 256), bool(f & 512), pps, fs, pk, pa, pb, strs[0], strs[1], strs[2], mask)
    return out
