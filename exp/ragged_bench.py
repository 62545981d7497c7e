#!/usr/bin/env python3
"""A/B timing of ragged-line batches (device pointers, CUDA events): python exp/ragged_bench.py [n_lines].
Select the library with NEEDLE_B200_LIB.  Not part of the product or the tests."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
dev = torch.device("cuda", 0)
data, off = workloads.c3_lines(n)
data_d = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
off_d = torch.from_numpy(off.view(np.int64)).to(dev)
m = torch.zeros(n, dtype=torch.uint8, device=dev)
s = torch.zeros(n, dtype=torch.int32, device=dev)
e = torch.zeros(n, dtype=torch.int32, device=dev)
nbytes = int(off[-1])
stream = torch.cuda.current_stream()
for name, regex in (("c3 e-mail", workloads.REGEX["c3"]), ("c2 ssn", workloads.REGEX["c2"]), ("digits+", r"[0-9]+"), ("c4", workloads.REGEX["c4"]),
                    ("sherlock", "[Ss]herlock")):
    pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
    for mode in (2, 1):
        def step():
            pat.match_batch_ptrs(mode, data_d.data_ptr(), off_d.data_ptr(), n, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
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
        print(f"{name:10s} mode {mode}: {nbytes / ms / 1e6:8.1f} GB/s  ({ms:.3f} ms, matches {int(m.sum())})", flush=True)

# long lines: fixed 512 / 4096 bytes and ragged 200..3000
for label, lens in (("fixed 512 B", np.full(1_000_000, 512)), ("fixed 4 KB", np.full(125_000, 4096)),
                    ("ragged 200..3000 B", np.random.default_rng(3).integers(200, 3000, size=300_000))):
    offl = np.zeros(len(lens) + 1, dtype=np.uint64)
    offl[1:] = np.cumsum(lens)
    tot = int(offl[-1])
    src, _ = workloads.c3_lines(tot // 60 + 1000)
    dl = torch.from_numpy(np.ascontiguousarray(src[:tot])).to(dev)
    ol = torch.from_numpy(offl.view(np.int64)).to(dev)
    nl = len(lens)
    for name, regex in (("c3 e-mail", workloads.REGEX["c3"]), ("c2 ssn", workloads.REGEX["c2"])):
        pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)

        def step():
            pat.match_batch_ptrs(2, dl.data_ptr(), ol.data_ptr(), nl, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
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
        print(f"{label:20s} {name:10s} find: {tot / ms / 1e6:8.1f} GB/s  ({ms:.3f} ms)", flush=True)
