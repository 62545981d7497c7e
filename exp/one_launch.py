#!/usr/bin/env python3
"""A few device-resident launches of find() over one workload, for ncu: python exp/one_launch.py c3 [lines] [launches] [mode].
Select the library with NEEDLE_B200_LIB.  Not part of the product or the tests."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402

GEN = {"c2": (workloads.c2_lines, 1), "c3": (workloads.c3_lines, 1), "c4": (workloads.c4_lines, 1), "c5": (workloads.c5_lines, 2)}
key = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 2
regex = os.environ.get("NDL_REGEX", workloads.REGEX[key])
gen, cw = GEN[key]
data, off = gen(n)
dev = torch.device("cuda", 0)
data_d = torch.from_numpy(np.ascontiguousarray(data).view(np.uint8)).to(dev)
off_d = torch.from_numpy(off.view(np.int64)).to(dev)
m = torch.zeros(n, dtype=torch.uint8, device=dev)
s = torch.zeros(n, dtype=torch.int32, device=dev)
e = torch.zeros(n, dtype=torch.int32, device=dev)
pat = nb.Pattern(nb.compile_to_bytes(regex, 0), device=0)
stream = torch.cuda.current_stream()
nbytes = int(off[-1] - off[0]) * cw
for _ in range(launches):
    pat.match_batch_ptrs(mode, data_d.data_ptr(), off_d.data_ptr(), n, cw, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(launches):
    pat.match_batch_ptrs(mode, data_d.data_ptr(), off_d.data_ptr(), n, cw, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=stream.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / launches
print(f"{key} mode {mode}: {nbytes / ms / 1e6:.1f} GB/s ({ms:.3f} ms per launch, {int(m.sum())} matches)")
