#!/usr/bin/env python3
"""One ragged shape, a few launches (for ncu): python exp/ragged_one.py lo hi [regex key].  Not part of the product."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
key = sys.argv[3] if len(sys.argv) > 3 else "c2"
total = 400_000_000
rng = np.random.default_rng(4)
src, _ = workloads.c3_lines(total // 60 + 1000)
data = torch.from_numpy(np.ascontiguousarray(src[:total])).cuda()
lens = rng.integers(lo, hi + 1, size=total // max(8, lo))
cs = np.cumsum(lens)
n = int(np.searchsorted(cs, total - 1))
off = np.zeros(n + 1, dtype=np.uint64)
off[1:] = cs[:n]
off_d = torch.from_numpy(off.view(np.int64)).cuda()
m = torch.zeros(n, dtype=torch.uint8, device="cuda")
s = torch.zeros(n, dtype=torch.int32, device="cuda")
e = torch.zeros(n, dtype=torch.int32, device="cuda")
pat = nb.Pattern(nb.compile_to_bytes(workloads.REGEX[key], 0), device=0)
stream = torch.cuda.current_stream()
for _ in range(4):
    pat.match_batch_ptrs(2, data.data_ptr(), off_d.data_ptr(), n, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
torch.cuda.synchronize()
print("done", n, int(off[-1]))
