#!/usr/bin/env python3
"""Timing of ndl_find_all_batch (all matches per haystack, two-pass CSR) on device-resident batches:
python exp/find_all_bench.py [n_lines].  Not part of the product or the tests."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from needle_b200 import _lib  # noqa: E402
from tests import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
L = _lib.lib()
stream = torch.cuda.current_stream()
for name, key, gen in (("c2 SSN, 64-byte lines", "c2", workloads.c2_lines), ("c3 e-mail, ragged lines", "c3", workloads.c3_lines),
                       ("c4 256-state, 64-byte lines", "c4", workloads.c4_lines), ("digits+ on c2 lines", None, workloads.c2_lines)):
    nn = n if key != "c3" else min(n, 4_000_000)
    data, off = gen(nn)
    regex = workloads.REGEX[key] if key else "[0-9]+"
    pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
    data_d = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
    off_d = torch.from_numpy(off.view(np.int64)).to(dev)
    counts = torch.zeros(nn, dtype=torch.int32, device=dev)
    nbytes = int(off[-1])

    def call(moff, st, en):
        rc = L.ndl_find_all_batch(pat._h, data_d.data_ptr(), off_d.data_ptr(), nn, 1, counts.data_ptr(), moff, st, en, nb.MEM_DEVICE,
                                  ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0, _lib.last_error()
    call(None, None, None)
    moff = torch.zeros(nn + 1, dtype=torch.int64, device=dev)
    moff[1:] = torch.cumsum(counts.to(torch.int64), 0)
    total = int(moff[-1].item())
    st = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
    en = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
    for label, args in (("count pass", (None, None, None)), ("fill pass", (moff.data_ptr(), st.data_ptr(), en.data_ptr()))):
        for _ in range(3):
            call(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call(*args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name:30s} {label:10s}: {nbytes / ms / 1e6:8.1f} GB/s ({ms:.3f} ms, {total} matches)", flush=True)
