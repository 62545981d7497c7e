#!/usr/bin/env python3
"""find() over ragged batches of several length distributions (device-resident, CUDA events): python exp/ragged_shapes_bench.py.
Not part of the product or the tests."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

dev = torch.device("cuda", 0)
total = 400_000_000
rng = np.random.default_rng(4)
src, _ = workloads.c3_lines(total // 60 + 1000)
src = src[:total]
data = torch.from_numpy(np.ascontiguousarray(src)).to(dev)
stream = torch.cuda.current_stream()
shapes = {
    "U[8,120] (C3)": lambda k: rng.integers(8, 121, size=k),
    "U[40,200]": lambda k: rng.integers(40, 201, size=k),
    "log-like N(150,60) in [20,400]": lambda k: np.clip(rng.normal(150, 60, size=k), 20, 400).astype(np.int64),
    "U[100,300]": lambda k: rng.integers(100, 301, size=k),
    "U[200,1000]": lambda k: rng.integers(200, 1001, size=k),
    "U[200,3000]": lambda k: rng.integers(200, 3001, size=k),
}
for name, gen in shapes.items():
    lens = gen(total // 8)
    cs = np.cumsum(lens)
    n = int(np.searchsorted(cs, total - 1))
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = cs[:n]
    off_d = torch.from_numpy(off.view(np.int64)).to(dev)
    m = torch.zeros(n, dtype=torch.uint8, device=dev)
    s = torch.zeros(n, dtype=torch.int32, device=dev)
    e = torch.zeros(n, dtype=torch.int32, device=dev)
    nbytes = int(off[-1])
    for key in ("c2", "c3"):
        pat = nb.Pattern(nb.compile_to_bytes(workloads.REGEX[key], 0), device=0)

        def step():
            pat.match_batch_ptrs(2, data.data_ptr(), off_d.data_ptr(), n, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:34s} {key}: {nbytes / ms / 1e6:8.1f} GB/s ({ms:.3f} ms, {n} lines)", flush=True)
