#!/usr/bin/env python3
"""find() over fixed-length records of several byte lengths (device-resident, CUDA events): python exp/reclen_bench.py.
Not part of the product or the tests."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

dev = torch.device("cuda", 0)
total = 640_000_000
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"0123456789abcdefghijklmnopqrstuvwxyz -", dtype=np.uint8)
data = torch.from_numpy(alpha[rng.integers(0, len(alpha), size=total, dtype=np.uint8)]).to(dev)
stream = torch.cuda.current_stream()
for key in ("c2", "c3"):
    pat = nb.Pattern(nb.compile_to_bytes(workloads.REGEX[key], 0), device=0)
    for L in (24, 40, 48, 64, 80, 96, 100, 112, 128, 160, 200, 256, 320, 400, 512, 768, 1024, 2048, 4096, 16384):
        n = total // L
        m = torch.zeros(n, dtype=torch.uint8, device=dev)
        s = torch.zeros(n, dtype=torch.int32, device=dev)
        e = torch.zeros(n, dtype=torch.int32, device=dev)

        def step():
            pat.match_lines_ptrs(2, data.data_ptr(), n, L, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{key} {L:4d}-byte records: {n * L / ms / 1e6:8.1f} GB/s ({ms:.3f} ms)", flush=True)
