"""ctypes wrapper of oracle/libneedle_oracle.so - the CPU restatement of the reference's generated loops.

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "libneedle_oracle.so")
INT32_MAX = 0x7FFFFFFF
INT64_MAX = 0x7FFFFFFFFFFFFFFF

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(ORACLE_PATH)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.ndlo_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(vp)]
        L.ndlo_from_tables.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp,
                                       ctypes.POINTER(vp)]
        L.ndlo_free.argtypes = [vp]
        L.ndlo_free.restype = None
        L.ndlo_matches.argtypes = [vp, vp, i64, ctypes.c_int]
        L.ndlo_contained_in.argtypes = [vp, vp, i64, ctypes.c_int]
        L.ndlo_index_forwards.argtypes = [vp, vp, i64, ctypes.c_int, i64]
        L.ndlo_index_forwards.restype = i64
        L.ndlo_index_backwards.argtypes = [vp, vp, ctypes.c_int, i64, i64, i64]
        L.ndlo_index_backwards.restype = i64
        L.ndlo_find.argtypes = [vp, vp, i64, ctypes.c_int, i64, i64, ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.ndlo_scan_from.argtypes = [vp, vp, i64, ctypes.c_int, i64, ctypes.c_int32, i64, ctypes.POINTER(ctypes.c_int32)]
        L.ndlo_scan_from.restype = i64
        L.ndlo_forwards_state_count.argtypes = [vp]
        L.ndlo_scan_back_from.argtypes = [vp, vp, ctypes.c_int, i64, i64, ctypes.c_int32, i64, ctypes.POINTER(ctypes.c_int32)]
        L.ndlo_scan_back_from.restype = i64
        L.ndlo_backwards_state_count.argtypes = [vp]
        L.ndlo_backwards_root_accepting.argtypes = [vp]
        L.ndlo_match_batch.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_uint64, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int]
        L.ndlo_match_batch_accel.argtypes = L.ndlo_match_batch.argtypes
        L.ndlo_index_forwards_accel.argtypes = [vp, vp, i64, ctypes.c_int, i64]
        L.ndlo_index_forwards_accel.restype = i64
        L.ndlo_accel_summary.argtypes = [vp]
        L.ndlo_accel_summary.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _encode(s):
    """str -> (buffer, n_chars, char_width) exactly like needle_b200.encode_haystack."""
    if isinstance(s, (bytes, bytearray)):
        raw, cw = bytes(s), 1
    else:
        try:
            raw, cw = s.encode("latin-1"), 1
        except UnicodeEncodeError:
            raw, cw = s.encode("utf-16-le", "surrogatepass"), 2
    buf = ctypes.create_string_buffer(raw, max(len(raw), 2))
    return buf, len(raw) // cw, cw


class Oracle:
    def __init__(self, blob=None, handle=None):
        self._h = ctypes.c_void_p()
        if blob is not None:
            if lib().ndlo_load(blob, len(blob), ctypes.byref(self._h)) != 0:
                raise ValueError("oracle could not parse the pattern blob")
        else:
            self._h = handle

    @classmethod
    def from_tables(cls, class_map, stride, min_length, max_length, reverse_mode, reverse_char, tables, accepting, max_char):
        """tables: 4 int16 arrays (n_states*stride), accepting: 4 uint8 arrays; order Matches, ContainedIn, Forwards, Backwards."""
        cm = np.ascontiguousarray(class_map, dtype=np.uint16)
        tabs = [np.ascontiguousarray(t, dtype=np.int16) for t in tables]
        accs = [np.ascontiguousarray(a, dtype=np.uint8) for a in accepting]
        n_states = (ctypes.c_int32 * 4)(*[len(a) for a in accs])
        mc = (ctypes.c_int32 * 4)(*max_char)
        tp = (ctypes.c_void_p * 4)(*[t.ctypes.data for t in tabs])
        ap = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in accs])
        h = ctypes.c_void_p()
        rc = lib().ndlo_from_tables(cm.ctypes.data, stride, min_length, max_length, reverse_mode, reverse_char,
                                    ctypes.cast(n_states, ctypes.c_void_p), ctypes.cast(mc, ctypes.c_void_p),
                                    ctypes.cast(tp, ctypes.c_void_p), ctypes.cast(ap, ctypes.c_void_p), ctypes.byref(h))
        if rc != 0:
            raise ValueError("ndlo_from_tables failed")
        return cls(handle=h)

    def __del__(self):
        try:
            if self._h:
                lib().ndlo_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- single string, the Matcher surface
    def matches(self, s) -> bool:
        buf, n, cw = _encode(s)
        return bool(lib().ndlo_matches(self._h, buf, n, cw))

    def contained_in(self, s) -> bool:
        buf, n, cw = _encode(s)
        return bool(lib().ndlo_contained_in(self._h, buf, n, cw))

    def find(self, s, from_=0):
        """First find(from, len) on a fresh Matcher: (matched, start, end)."""
        buf, n, cw = _encode(s)
        st, en = ctypes.c_int64(), ctypes.c_int64()
        m = lib().ndlo_find(self._h, buf, n, cw, from_, INT32_MAX, ctypes.byref(st), ctypes.byref(en))
        return bool(m), st.value, en.value

    def find_iter(self, s):
        """All successive find() results on one Matcher (nextStart = end).  A match that does not move nextStart forward
        (an empty match) would repeat forever in the reference: it is reported once and ends the list."""
        out, frm = [], 0
        while True:
            m, st, en = self.find(s, frm)
            if not m:
                break
            out.append((st, en))
            if en <= frm:
                break
            frm = en
        return out

    def find_all_batch(self, data, offsets, char_width=1, threads=1):
        """find_iter over a batch, as CSR (counts, match_offsets, starts, ends): rounds of ndlo_match_batch with `from` offsets."""
        n = len(offsets) - 1
        frm = np.zeros(n, dtype=np.int32)
        active = np.ones(n, dtype=bool)
        rounds = []
        while active.any():
            m, s, e = self.match_batch(2, data, offsets, char_width, frm, threads)
            hit = active & (m == 1)
            rounds.append((np.nonzero(hit)[0], s[hit].copy(), e[hit].copy()))
            adv = hit & (e > frm)
            frm = np.where(adv, e, frm).astype(np.int32)
            active = adv
        counts = np.zeros(n, dtype=np.uint32)
        for idx, _, _ in rounds:
            counts[idx] += 1
        moff = np.zeros(n + 1, dtype=np.uint64)
        moff[1:] = np.cumsum(counts, dtype=np.uint64)
        starts = np.zeros(int(moff[-1]), dtype=np.int32)
        ends = np.zeros(int(moff[-1]), dtype=np.int32)
        fill = moff[:-1].astype(np.int64).copy()
        for idx, s, e in rounds:
            starts[fill[idx]] = s
            ends[fill[idx]] = e
            fill[idx] += 1
        return counts, moff, starts, ends

    def scan_from(self, data, entry_state=0, last_init=-1, from_=0, cw=1):
        """(last, exit_state) of the forward scan of a chunk from an arbitrary state (multi-rank protocol check)."""
        data = np.ascontiguousarray(data).view(np.uint8)
        ex = ctypes.c_int32()
        last = lib().ndlo_scan_from(self._h, data.ctypes.data if data.size else None, data.size // cw, cw, from_, entry_state,
                                    last_init, ctypes.byref(ex))
        return last, ex.value

    def forwards_state_count(self):
        return lib().ndlo_forwards_state_count(self._h)

    def scan_back_from(self, data, index, entry_state=0, last_init=2 ** 63 - 1, lower=0, cw=1):
        """(start, exit_state) of the reverse scan of chunk[lower, index] from an arbitrary BACKWARDS state."""
        data = np.ascontiguousarray(data).view(np.uint8)
        ex = ctypes.c_int32()
        st = lib().ndlo_scan_back_from(self._h, data.ctypes.data if data.size else None, cw, index, lower, entry_state, last_init,
                                       ctypes.byref(ex))
        return st, ex.value

    def backwards_state_count(self):
        return lib().ndlo_backwards_state_count(self._h)

    def backwards_root_accepting(self):
        return bool(lib().ndlo_backwards_root_accepting(self._h))

    def accel_summary(self) -> str:
        """Which accelerators the accelerated indexForwards of this pattern uses (CompilationPolicy restated)."""
        return lib().ndlo_accel_summary(self._h).decode()

    def index_forwards(self, s, from_=0, accelerated=False):
        buf, n, cw = _encode(s)
        fn = lib().ndlo_index_forwards_accel if accelerated else lib().ndlo_index_forwards
        return fn(self._h, buf, n, cw, from_)

    # -- batches
    def match_batch(self, mode, data, offsets, char_width=1, from_=None, threads=1, accelerated=False):
        """ndlo_match_batch.  accelerated: find() runs indexForwards with the accelerators the reference's CompilationPolicy picks
        (indexOf prefix / suffix / infix seek, predicate seek, first-byte mask) - same results, the speed the JVM version has."""
        n = len(offsets) - 1
        data = np.ascontiguousarray(data).view(np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        matched = np.zeros(n, dtype=np.uint8)
        start = np.full(n, -1, dtype=np.int32) if mode == 2 else None
        end = np.full(n, -1, dtype=np.int32) if mode == 2 else None
        if from_ is not None:
            from_ = np.ascontiguousarray(from_, dtype=np.int32)
        if n:
            fn = lib().ndlo_match_batch_accel if accelerated else lib().ndlo_match_batch
            rc = fn(self._h, mode, data.ctypes.data if data.size else None, offsets.ctypes.data, n, char_width,
                    from_.ctypes.data if from_ is not None else None, matched.ctypes.data,
                    start.ctypes.data if start is not None else None,
                    end.ctypes.data if end is not None else None, threads)
            if rc != 0:
                raise RuntimeError("ndlo_match_batch failed")
        return matched, start, end
