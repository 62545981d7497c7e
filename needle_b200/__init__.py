"""needle_b200 - B200-native DFA regex matching behind needle's Pattern/Matcher surface.

Host side (this package + the C++ in csrc/host): regex -> NFA -> four DFAs -> table blob, restated from
hyperpape/needle.  Device side (csrc/kernels): hand-written sm_100a CUDA kernels that run the DFA
state-transition loop over batches of haystacks.  The boundary is the C ABI in include/needle_b200.h.
"""
from .pattern import (ALL_FLAGS, CASE_INSENSITIVE, DOTALL, LEFTMOST_LONGEST, UNICODE_CASE, UNICODE_CHARACTER_CLASS,
                      DFACompiler, Matcher, NeedleCudaError, Pattern, PatternClassCompilationException, PatternException,
                      PatternSyntaxException, Precompile, compile_to_bytes, encode_haystack, iter_find, pack_haystacks)
from ._lib import MODE_CONTAINEDIN, MODE_FIND, MODE_MATCHES, MEM_DEVICE, MEM_DEVICE_DATA, MEM_HOST

__all__ = [
    "ALL_FLAGS", "CASE_INSENSITIVE", "DOTALL", "LEFTMOST_LONGEST", "UNICODE_CASE", "UNICODE_CHARACTER_CLASS",
    "DFACompiler", "Matcher", "NeedleCudaError", "Pattern", "PatternClassCompilationException", "PatternException",
    "PatternSyntaxException", "Precompile", "compile_to_bytes", "encode_haystack", "iter_find", "pack_haystacks",
    "MODE_CONTAINEDIN", "MODE_FIND", "MODE_MATCHES", "MEM_DEVICE", "MEM_DEVICE_DATA", "MEM_HOST",
]
