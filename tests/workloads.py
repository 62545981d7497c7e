"""Synthetic inputs of the BASELINE.json configs (definitions: SURVEY.md section 8(d)), at any size.
Shared by the GPU parity tests and bench.py so that both see the same distributions."""
import numpy as np

REGEX = {
    "c1": r"http://.+",
    "c2": r"\d{3}-\d{2}-\d{4}",
    "c3": r"[A-Za-z0-9._%+-]+@[A-Za-z0-9.-]+",
    "c4": r"a[ab]{7}c",
    "c5": "[؀-ۿ]+",
}


def _alphabet(chars: str) -> np.ndarray:
    return np.frombuffer(chars.encode("latin-1"), dtype=np.uint8)


def c2_lines(n: int, seed: int = 0x5EED0002, line_len: int = 64):
    """n lines x 64 ASCII bytes over [0-9a-z -]; 25 % carry one planted ddd-dd-dddd.  Returns (data, offsets)."""
    rng = np.random.default_rng(seed)
    alpha = _alphabet("0123456789abcdefghijklmnopqrstuvwxyz -")
    data = alpha[rng.integers(0, len(alpha), size=n * line_len, dtype=np.uint8)]
    planted = np.nonzero(rng.random(n) < 0.25)[0]
    pos = rng.integers(0, line_len - 11 + 1, size=len(planted))
    digits = rng.integers(0, 10, size=(len(planted), 11), dtype=np.uint8) + ord("0")
    digits[:, 3] = ord("-")
    digits[:, 6] = ord("-")
    idx = (planted * line_len + pos)[:, None] + np.arange(11)[None, :]
    data[idx] = digits
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_len)
    return data, offsets


def c3_lines(n: int, seed: int = 0x5EED0003):
    """n ragged lines, length U[8,120], over [a-z0-9 .,;_-]; 30 % carry one planted local@domain."""
    rng = np.random.default_rng(seed)
    alpha = _alphabet("abcdefghijklmnopqrstuvwxyz0123456789 .,;_-")
    lens = rng.integers(8, 121, size=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    data = alpha[rng.integers(0, len(alpha), size=total, dtype=np.uint8)]
    local_alpha = _alphabet("abcdefghijklmnopqrstuvwxyz0123456789._%+-")
    dom_alpha = _alphabet("abcdefghijklmnopqrstuvwxyz0123456789.-")
    planted = np.nonzero(rng.random(n) < 0.30)[0]
    for i in planted:
        ll, dl = int(rng.integers(1, 13)), int(rng.integers(1, 17))
        addr = np.concatenate([local_alpha[rng.integers(0, len(local_alpha), ll)], [ord("@")],
                               dom_alpha[rng.integers(0, len(dom_alpha), dl)]]).astype(np.uint8)
        L = int(lens[i])
        if len(addr) > L:
            addr = addr[:L]
        p = int(rng.integers(0, L - len(addr) + 1))
        o = int(offsets[i]) + p
        data[o:o + len(addr)] = addr
    return data, offsets


def c4_lines(n: int, seed: int = 0x5EED0004, line_len: int = 64):
    """n lines x 64 bytes over {a,b} with a sprinkling of 'c' (batched variant of the 256-state DFA run)."""
    rng = np.random.default_rng(seed)
    data = (rng.integers(0, 2, size=n * line_len, dtype=np.uint8) + ord("a")).astype(np.uint8)
    cs = rng.integers(0, n * line_len, size=max(1, n // 2))
    data[cs] = ord("c")
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_len)
    return data, offsets


def c5_lines(n: int, seed: int = 0x5EED0005, line_chars: int = 32):
    """n lines x 32 UTF-16LE chars over U+0020-03FF; 20 % carry a planted run of 1-8 chars from U+0600-06FF."""
    rng = np.random.default_rng(seed)
    chars = rng.integers(0x20, 0x400, size=n * line_chars).astype(np.uint16)
    planted = np.nonzero(rng.random(n) < 0.20)[0]
    for i in planted:
        k = int(rng.integers(1, 9))
        p = int(rng.integers(0, line_chars - k + 1))
        chars[i * line_chars + p:i * line_chars + p + k] = rng.integers(0x600, 0x700, size=k)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(line_chars)
    return chars.view(np.uint8), offsets


def c2_lines_utf16(n: int, seed: int = 0x5EED0002, line_chars: int = 32):
    """The C2 lines as UTF-16LE (what the JNI shim gets from a java.lang.String): n lines x 32 chars = 64 bytes."""
    data, offsets = c2_lines(n, seed, line_chars)
    return data.astype(np.uint16).view(np.uint8), offsets


def c1_strings(n: int = 1000, seed: int = 0x5EED0001):
    rng = np.random.default_rng(seed)
    a = "abcdefghijklmnopqrstuvwxyz0123456789./"
    out = []
    for i in range(n):
        if rng.random() < 0.5:
            out.append("http://" + "".join(a[j] for j in rng.integers(0, len(a), int(rng.integers(5, 41)))))
        else:
            s = "".join((a + " ")[j] for j in rng.integers(0, len(a) + 1, int(rng.integers(5, 48))))
            out.append(s.replace("http://", "http:/-"))
    return out
