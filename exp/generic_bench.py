#!/usr/bin/env python3
"""Throughput of the per-thread generic walkers: find_all_batch and a pattern with no shared-memory image."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

n = 2_000_000
data, off = workloads.c3_lines(n)
for regex in (workloads.REGEX["c3"], r"[0-9]+", workloads.REGEX["c2"]):
    pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
    pat.find_all_batch(data, off, 1)
    t0 = time.perf_counter()
    counts, moff, st, en = pat.find_all_batch(data, off, 1)
    dt = time.perf_counter() - t0
    print(f"find_all_batch {regex[:20]:20s}: {int(off[-1]) / dt / 1e9:7.2f} GB/s end to end (two passes, host buffers), {int(counts.sum())} matches", flush=True)
# device-pointer timing of the count pass alone
dev = torch.device("cuda", 0)
d = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
o = torch.from_numpy(off.view(np.int64)).to(dev)
c = torch.zeros(n, dtype=torch.int32, device=dev)
L = nb._lib.lib()
for regex in (workloads.REGEX["c3"], r"[0-9]+"):
    pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
    for _ in range(2):
        L.ndl_find_all_batch(pat._h, d.data_ptr(), o.data_ptr(), n, 1, c.data_ptr(), None, None, None, 1, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.ndl_find_all_batch(pat._h, d.data_ptr(), o.data_ptr(), n, 1, c.data_ptr(), None, None, None, 1, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"count pass {regex[:20]:20s}: {int(off[-1]) / ms / 1e6:8.1f} GB/s ({ms:.3f} ms)", flush=True)
