#!/usr/bin/env python3
"""Throughput of patterns whose tables are too large for the replicated shared-memory images (fixed 64-byte lines)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402
import ctypes  # noqa: E402

n = 5_000_000
dev = torch.device("cuda", 0)
text = np.frombuffer(("Holmes and then Watson said to Mr. Sherlock Holmes, my dear Watson " * 8).encode(), dtype=np.uint8)
rng = np.random.default_rng(1)
data = text[(rng.integers(0, len(text) - 64, size=n)[:, None] + np.arange(64)[None, :])].reshape(-1).copy()
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(64)
data_d = torch.from_numpy(data).to(dev)
off_d = torch.from_numpy(off.view(np.int64)).to(dev)
m = torch.zeros(n, dtype=torch.uint8, device=dev)
s = torch.zeros(n, dtype=torch.int32, device=dev)
e = torch.zeros(n, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream()
name = nb._lib.lib().ndl_debug_kernel_name
name.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
name.restype = ctypes.c_char_p
for regex in ("Holmes.{1,10}Watson|Watson.{1,10}Holmes", "Holmes.{0,25}Watson|Watson.{0,25}Holmes", "[Ss]herlock", "Sherlock|Street", "(Holmes|Watson|Sherlock|Street|dear|said)+",
              "Sherlock|Holmes|Watson|Irene|Adler|John|Baker"):
    pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
    for mode in (2, 1):
        def step():
            pat.match_batch_ptrs(mode, data_d.data_ptr(), off_d.data_ptr(), n, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{regex[:44]:44s} mode {mode}: {n * 64 / ms / 1e6:8.1f} GB/s  states {pat.forwards_state_count}  {name(pat._h, mode, 1).decode()}  matches {int(m.sum())}", flush=True)
