#!/usr/bin/env python3
"""A/B of table layouts on fixed 64-byte lines (NDL_Q_FORCE is read at pattern creation): exp/layout_ab.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

n = 10_000_000
dev = torch.device("cuda", 0)
data, off = workloads.c4_lines(n)
data_d = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
off_d = torch.from_numpy(off.view(np.int64)).to(dev)
m = torch.zeros(n, dtype=torch.uint8, device=dev)
s = torch.zeros(n, dtype=torch.int32, device=dev)
e = torch.zeros(n, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream()
for regex in ("a[ab]{7}c", "a[ab]{6}c", "a[ab]{5}c", "a[ab]{4}c", "a[ab]{3}c"):
    for force in ("", "4,32,4", "2,32,4", "4,16,4", "4,8,4", "4,4,2", "4,2,2", "4,1,2", "2,16,4", "2,8,4"):
        if force:
            os.environ["NDL_Q_FORCE"] = force
        else:
            os.environ.pop("NDL_Q_FORCE", None)
        pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
        name = nb._lib.lib().ndl_debug_kernel_name
        name.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int, __import__("ctypes").c_int]
        name.restype = __import__("ctypes").c_char_p
        kn = name(pat._h, 2, 1).decode()
        if force and "linesq" not in kn:
            continue
        if force and f"{force.split(',')[0]} chars" not in kn:
            continue

        def step():
            pat.match_batch_ptrs(2, data_d.data_ptr(), off_d.data_ptr(), n, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"{regex:12s} force={force or 'auto':8s} {640.0 / ms:8.1f} GB/s  {kn}  matches {int(m.sum())}", flush=True)
