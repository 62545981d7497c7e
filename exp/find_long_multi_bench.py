#!/usr/bin/env python3
"""find() over one HOST-resident haystack: one GPU against every GPU of the box (ndl_pattern_create(device = -1), the split /
guess / resolve protocol inside the library).  python exp/find_long_multi_bench.py [GiB].  Not part of the product or the tests."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from needle_b200 import _lib  # noqa: E402
from tests import workloads  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30))
host = torch.randint(ord("a"), ord("b") + 1, (n,), dtype=torch.uint8).pin_memory()
host[n - 9:] = torch.tensor(list(b"abababbac"), dtype=torch.uint8)
pageable = host.numpy().copy() if gib <= 4 else None
blob = nb.compile_to_bytes(workloads.REGEX["c4"], 0)
print("GPUs:", _lib.lib().ndl_device_count())
for label, device in (("one GPU", 0), ("all GPUs", -1)):
    pat = nb.Pattern(blob, device=device)
    for name, ptr in (("pinned", host.data_ptr()), ("pageable", pageable.ctypes.data if pageable is not None else None)):
        if ptr is None:
            continue
        for _ in range(2):
            r = pat.find_long_ptrs(ptr, n, 1, 0, nb.MEM_HOST)
        assert r == (True, n - 9, n), r
        t0 = time.perf_counter()
        for _ in range(3):
            pat.find_long_ptrs(ptr, n, 1, 0, nb.MEM_HOST)
        dt = (time.perf_counter() - t0) / 3
        print(f"{label:9s} {name:9s}: {dt * 1e3:8.1f} ms  {n / dt / 1e9:7.1f} GB/s end to end", flush=True)
