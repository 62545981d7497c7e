"""Host-side mirror of needle's public surface over the C ABI.

Same names, argument meaning and error behaviour as the reference so that its tests read the same here:

    DFACompiler.compile(regex, className[, flags])   needle-compiler/.../DFACompiler.java:16-37
    Pattern.matcher(s) -> Matcher                     needle-types/.../Pattern.java:33
    Matcher.matches / containedIn / find / find(start, end) / start / end     Matcher.java:14-25
    Precompile.precompile(regex, className, dir[, flags])                    precompile/Precompile.java:30-53

plus the one call a GPU needs and the reference lacks: `Pattern.match_batch` over many haystacks.
Every match call goes to the CUDA kernels through `ndl_match_batch`; there is no CPU path here.
"""
import ctypes
import os
from typing import Iterable, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib

# com.justinblank.strings.Pattern flag constants (Pattern.java:9-31)
DOTALL = 0x20
CASE_INSENSITIVE = 0x02
UNICODE_CASE = 0x40
UNICODE_CHARACTER_CLASS = 0x100
LEFTMOST_LONGEST = 0x800000
ALL_FLAGS = DOTALL | CASE_INSENSITIVE | UNICODE_CASE | LEFTMOST_LONGEST | UNICODE_CHARACTER_CLASS

INT_MAX = 0x7FFFFFFF


NO_START = 2 ** 63 - 1  # "no accepting index yet" of a reverse scan (the reference uses Integer.MAX_VALUE)


class PatternException(RuntimeError):
    """com.justinblank.strings.PatternException"""


class PatternSyntaxException(PatternException):
    """com.justinblank.strings.PatternSyntaxException (RegexParser.java:91-97)"""


class PatternClassCompilationException(PatternException):
    """com.justinblank.strings.PatternClassCompilationException (DFACompiler.java:34-36, 71-73).
    `__cause__` carries the PatternSyntaxException / IllegalStateException analogue, as in the reference."""


class NeedleCudaError(RuntimeError):
    """A CUDA failure, or no usable device: the match path has no fallback."""


def _raise(code: int, what: str):
    msg = f"{what}: {_lib.last_error()}"
    if code == _lib.NDL_ESYNTAX:
        # DFACompiler.compileToBytes wraps the parser's exception (DFACompiler.java:71-73)
        raise PatternClassCompilationException(msg) from PatternSyntaxException(_lib.last_error())
    if code == _lib.NDL_ETOOLARGE:
        raise PatternClassCompilationException(msg) from OverflowError(_lib.last_error())  # IllegalStateException
    if code == _lib.NDL_EFLAGS:
        raise ValueError(msg)  # IllegalArgumentException (CompilerOptions.java:10-12)
    if code == _lib.NDL_ECOMPILE:
        raise PatternClassCompilationException(msg)
    if code in (_lib.NDL_ECUDA, _lib.NDL_ENCCL):
        raise NeedleCudaError(msg)
    if code == _lib.NDL_EBLOB:
        raise ValueError(msg)
    raise RuntimeError(f"{msg} (code {code})")


def compile_to_bytes(regex: str, flags: int = 0) -> bytes:
    """regex -> table blob (host only, no GPU needed).  DFACompiler.compileToBytes analogue."""
    if regex is None:
        raise TypeError("regex string cannot be null")
    L = _lib.lib()
    u = regex.encode("utf-16-le", "surrogatepass")
    buf = ctypes.create_string_buffer(u, len(u)) if u else None
    blob = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    rc = L.ndl_compile(ctypes.cast(buf, ctypes.c_void_p) if buf is not None else None, len(u) // 2, flags,
                       ctypes.byref(blob), ctypes.byref(n))
    if rc != _lib.NDL_OK:
        _raise(rc, f"Failed to create pattern for regex '{regex}'")
    try:
        return ctypes.string_at(blob, n.value)
    finally:
        L.ndl_blob_free(blob)


def encode_haystack(s: str) -> Tuple[bytes, int]:
    """A java.lang.String as packed chars: Latin-1 bytes when every char fits (byte b == char b),
    else UTF-16LE code units.  Returns (bytes, char_width)."""
    try:
        return s.encode("latin-1"), 1
    except UnicodeEncodeError:
        return s.encode("utf-16-le", "surrogatepass"), 2


def pack_haystacks(strings: Sequence[Union[str, bytes]], char_width: Optional[int] = None):
    """Pack strings into (data uint8 array, offsets uint64[n+1] in chars, char_width)."""
    if char_width is None:
        char_width = 1
        for s in strings:
            if isinstance(s, str) and any(ord(ch) > 0xFF for ch in s):
                char_width = 2
                break
    parts = []
    offsets = np.zeros(len(strings) + 1, dtype=np.uint64)
    total = 0
    for i, s in enumerate(strings):
        if isinstance(s, str):
            b = s.encode("latin-1") if char_width == 1 else s.encode("utf-16-le", "surrogatepass")
        else:
            b = bytes(s)
        parts.append(b)
        total += len(b) // char_width
        offsets[i + 1] = total
    data = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if total else np.zeros(0, dtype=np.uint8)
    return data, offsets, char_width


class Pattern:
    """A compiled regex resident on one GPU (`ndl_pattern`).  Stateless and shareable, like the
    reference's generated Pattern class."""

    def __init__(self, blob: bytes, device: int = 0, regex: Optional[str] = None, class_name: Optional[str] = None):
        self.blob = bytes(blob)
        self.regex = regex
        self.class_name = class_name
        self._h = ctypes.c_void_p()
        rc = _lib.lib().ndl_pattern_create(self.blob, len(self.blob), device, ctypes.byref(self._h))
        if rc != _lib.NDL_OK:
            self._h = None
            _raise(rc, "ndl_pattern_create")
        self.device = device

    @classmethod
    def from_file(cls, path: str, device: int = 0) -> "Pattern":
        """Load a blob written by Precompile.precompile."""
        with open(path, "rb") as f:
            return cls(f.read(), device, class_name=os.path.splitext(os.path.basename(path))[0])

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().ndl_pattern_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self) -> "_lib.BlobInfo":
        out = _lib.BlobInfo()
        rc = _lib.lib().ndl_blob_info_get(self.blob, len(self.blob), ctypes.byref(out))
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_blob_info_get")
        return out

    # -- the reference surface
    def matcher(self, s: str) -> "Matcher":
        return Matcher(self, s)

    # -- the batch call
    def match_batch(self, mode: int, data: np.ndarray, offsets: np.ndarray, char_width: int = 1,
                    from_: Optional[np.ndarray] = None):
        """Run one batch from HOST buffers through `ndl_match_batch` (copies inside the call).
        Returns (matched uint8[n], start int32[n], end int32[n]); start/end are None unless mode is find."""
        n = len(offsets) - 1
        data = np.ascontiguousarray(data).view(np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        matched = np.zeros(n, dtype=np.uint8)
        start = np.full(n, -1, dtype=np.int32) if mode == _lib.MODE_FIND else None
        end = np.full(n, -1, dtype=np.int32) if mode == _lib.MODE_FIND else None
        if from_ is not None:
            from_ = np.ascontiguousarray(from_, dtype=np.int32)
        if n == 0:
            return matched, start, end
        rc = _lib.lib().ndl_match_batch(
            self._h, mode, data.ctypes.data if data.size else None, offsets.ctypes.data, n, char_width,
            from_.ctypes.data if from_ is not None else None, matched.ctypes.data,
            start.ctypes.data if start is not None else None, end.ctypes.data if end is not None else None,
            _lib.MEM_HOST, None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_match_batch")
        return matched, start, end

    def match_batch_ptrs(self, mode: int, data_ptr: int, offsets_ptr: int, n: int, char_width: int, matched_ptr: int,
                         start_ptr: int = 0, end_ptr: int = 0, from_ptr: int = 0, mem_kind: int = _lib.MEM_DEVICE,
                         stream: int = 0):
        """Raw-pointer form (device tensors from torch, or pinned host buffers): one stream-ordered pass."""
        rc = _lib.lib().ndl_match_batch(self._h, mode, data_ptr or None, offsets_ptr or None, n, char_width,
                                        from_ptr or None, matched_ptr or None, start_ptr or None, end_ptr or None,
                                        mem_kind, stream or None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_match_batch")

    def match_lines(self, mode: int, data: np.ndarray, n: int, line_chars: int, char_width: int = 1):
        """match_batch over n fixed-length records of `line_chars` chars from a HOST buffer (ndl_match_lines): no offsets array."""
        data = np.ascontiguousarray(data).view(np.uint8)
        matched = np.zeros(n, dtype=np.uint8)
        start = np.full(n, -1, dtype=np.int32) if mode == _lib.MODE_FIND else None
        end = np.full(n, -1, dtype=np.int32) if mode == _lib.MODE_FIND else None
        if n:
            rc = _lib.lib().ndl_match_lines(self._h, mode, data.ctypes.data if data.size else None, n, line_chars, char_width,
                                            matched.ctypes.data, start.ctypes.data if start is not None else None,
                                            end.ctypes.data if end is not None else None, _lib.MEM_HOST, None)
            if rc != _lib.NDL_OK:
                _raise(rc, "ndl_match_lines")
        return matched, start, end

    def match_lines_ptrs(self, mode: int, data_ptr: int, n: int, line_chars: int, char_width: int, matched_ptr: int, start_ptr: int = 0,
                         end_ptr: int = 0, mem_kind: int = _lib.MEM_DEVICE, stream: int = 0):
        """Raw-pointer form of match_lines (device tensors, or pinned host buffers)."""
        rc = _lib.lib().ndl_match_lines(self._h, mode, data_ptr or None, n, line_chars, char_width, matched_ptr or None,
                                        start_ptr or None, end_ptr or None, mem_kind, stream or None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_match_lines")

    def find_all_batch(self, data: np.ndarray, offsets: np.ndarray, char_width: int = 1):
        """All non-overlapping matches of every haystack (`while (m.find())`, DFACompilerTest.java:678-699) from HOST
        buffers, as CSR: returns (counts uint32[n], match_offsets uint64[n+1], starts int32[total], ends int32[total]).
        Two passes through `ndl_find_all_batch`: count, prefix sum on the host, fill."""
        n = len(offsets) - 1
        data = np.ascontiguousarray(data).view(np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint32)
        moff = np.zeros(n + 1, dtype=np.uint64)
        if n == 0:
            return counts, moff, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
        L = _lib.lib()
        rc = L.ndl_find_all_batch(self._h, data.ctypes.data if data.size else None, offsets.ctypes.data, n, char_width,
                                  counts.ctypes.data, None, None, None, _lib.MEM_HOST, None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_find_all_batch")
        moff[1:] = np.cumsum(counts, dtype=np.uint64)
        total = int(moff[-1])
        starts = np.full(max(total, 1), -1, dtype=np.int32)
        ends = np.full(max(total, 1), -1, dtype=np.int32)
        rc = L.ndl_find_all_batch(self._h, data.ctypes.data if data.size else None, offsets.ctypes.data, n, char_width,
                                  counts.ctypes.data, moff.ctypes.data, starts.ctypes.data, ends.ctypes.data, _lib.MEM_HOST, None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_find_all_batch")
        return counts, moff, starts[:total], ends[:total]

    def find_long(self, data: np.ndarray, from_: int = 0, char_width: int = 1):
        """find() over ONE haystack of any length (64-bit offsets; BASELINE config 4) from a HOST array.
        Returns (matched, start, end)."""
        data = np.ascontiguousarray(data).view(np.uint8)
        return self.find_long_ptrs(data.ctypes.data if data.size else 0, data.size // char_width, char_width, from_, _lib.MEM_HOST)

    def find_long_ptrs(self, data_ptr: int, n_chars: int, char_width: int = 1, from_: int = 0, mem_kind: int = _lib.MEM_DEVICE,
                       stream: int = 0):
        """Raw-pointer form of find_long (device memory by default).  The three results come back on the host."""
        m = ctypes.c_uint8()
        st, en = ctypes.c_int64(), ctypes.c_int64()
        if mem_kind == _lib.MEM_DEVICE:
            mem_kind = _lib.MEM_DEVICE_DATA  # device haystack, the three scalars through host pointers
        rc = _lib.lib().ndl_find_long(self._h, data_ptr or None, n_chars, char_width, from_, ctypes.byref(m), ctypes.byref(st),
                                      ctypes.byref(en), mem_kind, stream or None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_find_long")
        return bool(m.value), st.value, en.value

    def find_long_from(self, data_ptr: int, n_chars: int, entry_state: int = 0, char_width: int = 1, from_: int = 0,
                       mem_kind: int = _lib.MEM_HOST, stream: int = 0):
        """Forward scan of one rank's chunk of a haystack split across GPUs (ndl_find_long_from), started in
        `entry_state` of the FORWARDS automaton.  Returns (end, exit_state): `end` = chunk-local index after the last
        accepting step or -1, `exit_state` = state after the last char read (forwards_state_count = dead)."""
        ex = ctypes.c_int32()
        if mem_kind == _lib.MEM_DEVICE:
            mem_kind = _lib.MEM_DEVICE_DATA  # the haystack stays on the device, the three scalars come back through host pointers
        m, en = ctypes.c_uint8(), ctypes.c_int64()
        rc = _lib.lib().ndl_find_long_from(self._h, data_ptr or None, n_chars, char_width, from_, entry_state, -1,
                                           ctypes.byref(m), None, ctypes.byref(en), ctypes.byref(ex), mem_kind, stream or None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_find_long_from")
        return en.value, ex.value

    def find_long_back(self, data_ptr: int, n_chars: int, index: int, entry_state: int = 0, last_init: int = NO_START,
                       char_width: int = 1, lower: int = 0, mem_kind: int = _lib.MEM_HOST, stream: int = 0):
        """Reverse scan of one rank's chunk (ndl_find_long_back) over [lower, index] downwards from `entry_state` of
        the BACKWARDS automaton.  Returns (start, exit_state): smallest accepting chunk-local index or last_init."""
        st, ex = ctypes.c_int64(), ctypes.c_int32()
        rc = _lib.lib().ndl_find_long_back(self._h, data_ptr or None, n_chars, char_width, index, lower, entry_state, last_init,
                                           ctypes.byref(st), ctypes.byref(ex), mem_kind, stream or None)
        if rc != _lib.NDL_OK:
            _raise(rc, "ndl_find_long_back")
        return st.value, ex.value

    def walk_host(self, data: np.ndarray, entry_state: int = 0, char_width: int = 1) -> int:
        """The FORWARDS automaton walked on the host over a (short) array from `entry_state`; returns the exit state.  The
        entry-state guess of a sharded find (ndl_forwards_walk_host): no device work."""
        data = np.ascontiguousarray(data).view(np.uint8)
        r = _lib.lib().ndl_forwards_walk_host(self._h, data.ctypes.data if data.size else None, data.size // char_width, char_width,
                                              entry_state)
        if r < 0:
            raise ValueError("ndl_forwards_walk_host: bad argument")
        return r

    @property
    def forwards_state_count(self) -> int:
        return _lib.lib().ndl_forwards_state_count(self._h)

    @property
    def backwards_state_count(self) -> int:
        return _lib.lib().ndl_backwards_state_count(self._h)

    @property
    def backwards_root_accepting(self) -> bool:
        return bool(_lib.lib().ndl_backwards_root_accepting(self._h))

    @property
    def reverse_mode(self) -> int:
        return _lib.lib().ndl_reverse_mode(self._h)

    @property
    def min_length(self) -> int:
        return _lib.lib().ndl_min_length(self._h)

    def find_all(self, strings: Sequence[Union[str, bytes]]):
        """Convenience: first find() per string.  Returns list of (matched, start, end)."""
        data, offsets, cw = pack_haystacks(strings)
        m, s, e = self.match_batch(_lib.MODE_FIND, data, offsets, cw)
        return [(bool(a), int(b), int(c)) for a, b, c in zip(m, s, e)]


class Matcher:
    """com.justinblank.strings.Matcher over one string.  Not thread safe (mutable nextStart/start/end,
    DFAClassBuilder.java:688-694), like the reference."""

    def __init__(self, pattern: Pattern, s: str):
        if s is None:
            raise TypeError("string cannot be null")
        self._p = pattern
        self.string = s
        raw, self._cw = encode_haystack(s)
        self._data = np.frombuffer(raw, dtype=np.uint8).copy() if raw else np.zeros(0, dtype=np.uint8)
        self.length = len(raw) // self._cw
        self._offsets = np.array([0, self.length], dtype=np.uint64)
        self._next_start = 0
        self._start = -1
        self._end = -1

    def _run(self, mode, from_=None):
        f = None if from_ is None else np.array([from_], dtype=np.int32)
        return self._p.match_batch(mode, self._data, self._offsets, self._cw, f)

    def matches(self) -> bool:
        return bool(self._run(_lib.MODE_MATCHES)[0][0])

    def containedIn(self) -> bool:
        return bool(self._run(_lib.MODE_CONTAINEDIN)[0][0])

    contained_in = containedIn

    def find(self, start: Optional[int] = None, end: Optional[int] = None) -> bool:
        """find() resumes at nextStart; find(start, end) honours `start` and - like the reference, whose
        indexForwards overwrites its second argument (DFAClassBuilder.java:349) - ignores `end`."""
        if start is None:
            start = self._next_start
        if self._next_start == -1:
            return False
        m, s, e = self._run(_lib.MODE_FIND, start)
        self._end = self._next_start = int(e[0])
        if m[0]:
            self._start = int(s[0])
            return True
        return False

    def start(self) -> int:
        return self._start

    def end(self) -> int:
        return self._end


class DFACompiler:
    """com.justinblank.strings.DFACompiler"""

    @staticmethod
    def compile(regex: str, class_name: str, flags: int = 0, device: int = 0) -> Pattern:
        if class_name is None:
            raise PatternClassCompilationException("name cannot be null")
        return Pattern(compile_to_bytes(regex, flags), device, regex=regex, class_name=class_name)

    @staticmethod
    def compileToBytes(regex: str, class_name: str, flags: int = 0) -> bytes:
        if class_name is None:
            raise PatternClassCompilationException("name cannot be null")
        return compile_to_bytes(regex, flags)

    compile_to_bytes = compileToBytes


class Precompile:
    """com.justinblank.strings.precompile.Precompile: writes `<dir>/<className>.ndlb` (the table blob)
    where the reference writes `<dir>/<className>.class`."""

    @staticmethod
    def precompile(regex: str, class_name: str, directory: str, flags: int = 0) -> str:
        blob = compile_to_bytes(regex, flags)
        path = os.path.join(directory, class_name + ".ndlb")
        with open(path, "wb") as f:
            f.write(blob)
        return path


def iter_find(pattern: Pattern, s: str) -> Iterable[Tuple[int, int]]:
    """while (m.find()) yield (m.start(), m.end()) - the loop of DFACompilerTest.java:678-699."""
    m = pattern.matcher(s)
    nxt = 0
    while m.find():
        yield (m.start(), m.end())
        if m.end() <= nxt:  # nextStart did not advance (empty match): the reference would repeat it forever; stop instead
            break
        nxt = m.end()
