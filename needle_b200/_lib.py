"""ctypes binding of libneedle_b200.so (the C ABI declared in include/needle_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C needle_b200/csrc`.  If it is missing
the import fails loudly - there is no Python or CPU fallback for the match path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NEEDLE_B200_LIB", os.path.join(_HERE, "libneedle_b200.so"))

NDL_OK = 0
NDL_ESYNTAX, NDL_ECOMPILE, NDL_ETOOLARGE, NDL_EFLAGS = -1, -2, -3, -4
NDL_ECUDA, NDL_ENCCL, NDL_EINVAL, NDL_EBLOB, NDL_ENOMEM = -5, -6, -7, -8, -9
MODE_MATCHES, MODE_CONTAINEDIN, MODE_FIND = 0, 1, 2
MEM_HOST, MEM_DEVICE, MEM_DEVICE_DATA = 0, 1, 2


class BlobInfo(ctypes.Structure):
    _fields_ = [
        ("version", ctypes.c_int32), ("flags", ctypes.c_int32), ("min_length", ctypes.c_int32),
        ("max_length", ctypes.c_int32), ("stride", ctypes.c_int32), ("byte_class_count", ctypes.c_int32),
        ("reverse_mode", ctypes.c_int32), ("reverse_char", ctypes.c_int32),
        ("n_states", ctypes.c_int32 * 4), ("entry_width", ctypes.c_int32 * 4),
        ("max_char", ctypes.c_int32 * 4), ("n_accepting", ctypes.c_int32 * 4),
    ]


# every symbol include/needle_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "ndl_compile", "ndl_compile_utf8", "ndl_blob_free", "ndl_blob_info_get", "ndl_pattern_create",
    "ndl_pattern_destroy", "ndl_match_batch", "ndl_find_long", "ndl_last_error", "ndl_version",
    "ndl_device_count", "ndl_kernel_launches", "ndl_pattern_device", "ndl_find_long_from", "ndl_forwards_state_count",
    "ndl_find_long_back", "ndl_backwards_state_count", "ndl_backwards_root_accepting", "ndl_reverse_mode", "ndl_min_length",
    "ndl_find_all_batch", "ndl_match_lines", "ndl_pattern_device_count", "ndl_host_alloc", "ndl_host_free", "ndl_forwards_walk_host",
)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C needle_b200/csrc`).  needle_b200 has no fallback without its native library.")
    L = ctypes.CDLL(LIB_PATH)
    u8p, i32p, u64p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint64)
    L.ndl_compile.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(u8p), ctypes.POINTER(ctypes.c_size_t)]
    L.ndl_compile.restype = ctypes.c_int
    L.ndl_compile_utf8.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(u8p), ctypes.POINTER(ctypes.c_size_t)]
    L.ndl_compile_utf8.restype = ctypes.c_int
    L.ndl_blob_free.argtypes = [u8p]
    L.ndl_blob_free.restype = None
    L.ndl_blob_info_get.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(BlobInfo)]
    L.ndl_blob_info_get.restype = ctypes.c_int
    L.ndl_pattern_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.ndl_pattern_create.restype = ctypes.c_int
    L.ndl_pattern_destroy.argtypes = [ctypes.c_void_p]
    L.ndl_pattern_destroy.restype = None
    L.ndl_match_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_void_p]
    L.ndl_match_batch.restype = ctypes.c_int
    L.ndl_match_lines.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ndl_match_lines.restype = ctypes.c_int
    L.ndl_find_all_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ndl_find_all_batch.restype = ctypes.c_int
    L.ndl_find_long.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ndl_find_long.restype = ctypes.c_int
    L.ndl_find_long_from.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.c_void_p]
    L.ndl_find_long_from.restype = ctypes.c_int
    L.ndl_find_long_back.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ndl_find_long_back.restype = ctypes.c_int
    L.ndl_forwards_walk_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int32]
    L.ndl_forwards_walk_host.restype = ctypes.c_int32
    for f in ("ndl_forwards_state_count", "ndl_backwards_state_count", "ndl_backwards_root_accepting", "ndl_reverse_mode", "ndl_min_length"):
        getattr(L, f).argtypes = [ctypes.c_void_p]
        getattr(L, f).restype = ctypes.c_int
    L.ndl_last_error.restype = ctypes.c_char_p
    L.ndl_version.restype = ctypes.c_char_p
    L.ndl_device_count.restype = ctypes.c_int
    L.ndl_kernel_launches.restype = ctypes.c_uint64
    L.ndl_pattern_device.argtypes = [ctypes.c_void_p]
    L.ndl_pattern_device.restype = ctypes.c_int
    L.ndl_pattern_device_count.argtypes = [ctypes.c_void_p]
    L.ndl_pattern_device_count.restype = ctypes.c_int
    L.ndl_host_alloc.argtypes = [ctypes.c_size_t]
    L.ndl_host_alloc.restype = ctypes.c_void_p
    L.ndl_host_free.argtypes = [ctypes.c_void_p]
    L.ndl_host_free.restype = None
    _ = (i32p, u64p)
    _lib = L
    return L


def last_error() -> str:
    return lib().ndl_last_error().decode("utf-8", "replace")
