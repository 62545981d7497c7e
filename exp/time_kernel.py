"""Scratch: time the device-resident batch call only (no parity guard). Usage: NEEDLE_B200_LIB=... python exp/time_kernel.py [workload]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import needle_b200 as nb
from tests import workloads
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
gen = {"c2": workloads.c2_lines, "c3": workloads.c3_lines, "c4b": workloads.c4_lines}[w]
key = {"c2": "c2", "c3": "c3", "c4b": "c4"}[w]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
data, off = gen(n)
pat = nb.Pattern(nb.compile_to_bytes(workloads.REGEX[key], 0))
d = torch.from_numpy(data).cuda(); o = torch.from_numpy(off.view(np.int64)).cuda()
m = torch.zeros(n, dtype=torch.uint8, device="cuda"); s = torch.zeros(n, dtype=torch.int32, device="cuda"); e = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
f = lambda: pat.match_batch_ptrs(2, d.data_ptr(), o.data_ptr(), n, 1, m.data_ptr(), s.data_ptr(), e.data_ptr(), stream=st)
for _ in range(5): f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); 
for _ in range(100): f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 100
print(f"{w} lib={os.environ.get('NEEDLE_B200_LIB','default')} {ms*1000:.1f} us/launch  {len(data)/ms/1e6:.0f} GB/s  matches={int(m.sum())}")
