#!/usr/bin/env python3
"""find() over one device-resident UTF-16 haystack (ndl_find_long, char_width 2).  Not part of the product or the tests."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from needle_b200 import _lib  # noqa: E402
from tests import workloads  # noqa: E402

n = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 2 << 30  # chars
g = torch.Generator(device="cuda")
g.manual_seed(7)
for name, regex, lo, hi, plant in (("c4 a[ab]{7}c", workloads.REGEX["c4"], ord("a"), ord("b") + 1, list(b"abababbac")),
                                   ("c5 BMP class", workloads.REGEX["c5"], ord("a"), ord("z") + 1, [0x0627, 0x0644])):
    data = torch.randint(lo, hi, (n,), dtype=torch.int16, device="cuda", generator=g)
    data[n - len(plant):] = torch.tensor(plant, dtype=torch.int16, device="cuda")
    pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
    for _ in range(2):
        r = pat.find_long_ptrs(data.data_ptr(), n, 2, 0, nb.MEM_DEVICE)
    assert r == (True, n - len(plant), n), r
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pat.find_long_ptrs(data.data_ptr(), n, 2, 0, nb.MEM_DEVICE)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n} chars = {2 * n / 1e9:.1f} GB, passes {_lib.lib().ndl_debug_long_passes()}, {dt * 1e3:.2f} ms = {2 * n / dt / 1e9:.0f} GB/s", flush=True)
    del data
