#!/usr/bin/env python3
"""One record length, a few launches (for ncu): python exp/reclen_one.py L [regex key].  Not part of the product."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

L = int(sys.argv[1])
key = sys.argv[2] if len(sys.argv) > 2 else "c2"
total = 640_000_000
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"0123456789abcdefghijklmnopqrstuvwxyz -", dtype=np.uint8)
data = torch.from_numpy(alpha[rng.integers(0, len(alpha), size=total, dtype=np.uint8)]).cuda()
pat = nb.Pattern(nb.compile_to_bytes(workloads.REGEX[key], 0), device=0)
n = total // L
m = torch.zeros(n, dtype=torch.uint8, device="cuda")
s = torch.zeros(n, dtype=torch.int32, device="cuda")
e = torch.zeros(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream()
for _ in range(4):
    pat.match_lines_ptrs(2, data.data_ptr(), n, L, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pat.match_lines_ptrs(2, data.data_ptr(), n, L, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{key} {L}-byte records: {n * L / ms / 1e6:.1f} GB/s ({ms:.3f} ms)")
