"""The C-ABI library loads and exports every symbol include/needle_b200.h declares.  CPU only: no
compute entry point is called without a GPU (it must refuse, not fall back)."""
import ctypes
import os
import re

import pytest

import needle_b200 as nb
from needle_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "needle_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ndl_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = declared_symbols()
    assert set(names) == set(_lib.SYMBOLS), (names, _lib.SYMBOLS)
    for n in names:
        assert hasattr(L, n), n


def test_version_and_flags():
    assert b"needle_b200" in _lib.lib().ndl_version()
    # identical ints to com.justinblank.strings.Pattern / java.util.regex.Pattern
    assert (nb.CASE_INSENSITIVE, nb.DOTALL, nb.UNICODE_CASE, nb.UNICODE_CHARACTER_CLASS, nb.LEFTMOST_LONGEST) == (2, 32, 64, 256, 0x800000)


def test_error_codes():
    L = _lib.lib()
    blob = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    assert L.ndl_compile_utf8(b"(", 0, ctypes.byref(blob), ctypes.byref(n)) == _lib.NDL_ESYNTAX
    assert b"Regex=" in L.ndl_last_error()
    assert L.ndl_compile_utf8(b"a", 4, ctypes.byref(blob), ctypes.byref(n)) == _lib.NDL_EFLAGS
    assert L.ndl_compile_utf8(b"a", 0, None, None) == _lib.NDL_EINVAL
    assert L.ndl_compile_utf8(b"abc", 0, ctypes.byref(blob), ctypes.byref(n)) == _lib.NDL_OK
    L.ndl_blob_free(blob)


def test_utf8_and_utf16_entry_points_agree():
    L = _lib.lib()
    blob = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    regex = "[ά-ώ]+ε|λ"
    assert L.ndl_compile_utf8(regex.encode("utf-8"), 0, ctypes.byref(blob), ctypes.byref(n)) == 0
    a = ctypes.string_at(blob, n.value)
    L.ndl_blob_free(blob)
    assert a == nb.compile_to_bytes(regex, 0)


@pytest.mark.skipif(_lib.lib().ndl_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_a_gpu():
    blob = nb.compile_to_bytes("abc", 0)
    with pytest.raises(nb.NeedleCudaError):
        nb.Pattern(blob, device=0)


def test_precompile_writes_the_blob(tmp_path):
    # Precompile.precompile analogue (precompile/Precompile.java:30-53)
    path = nb.Precompile.precompile("http://.+", "OversimplifiedURLMatcher", str(tmp_path))
    assert os.path.basename(path) == "OversimplifiedURLMatcher.ndlb"
    with open(path, "rb") as f:
        assert f.read() == nb.compile_to_bytes("http://.+", 0)
