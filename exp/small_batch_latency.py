#!/usr/bin/env python3
"""Latency of small host batches through ndl_match_batch (BASELINE configs[0] shape: 1k short strings).  Not part of the product."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needle_b200 as nb  # noqa: E402
from tests import workloads  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402

blob = nb.compile_to_bytes("http://.+", 0)
pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
strings = workloads.c1_strings(1000)
for n in (1, 32, 1000, 10_000, 100_000):
    ss = (strings * (n // len(strings) + 1))[:n]
    data, off, cw = nb.pack_haystacks(ss)
    for mode in (0, 2):
        for _ in range(20):
            pat.match_batch(mode, data, off, cw)
        t0 = time.perf_counter()
        reps = 200 if n <= 10_000 else 30
        for _ in range(reps):
            pat.match_batch(mode, data, off, cw)
        dt = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(max(1, reps // 10)):
            ora.match_batch(mode, data, off, cw)
        dc = (time.perf_counter() - t0) / max(1, reps // 10)
        print(f"n={n:7d} mode {mode}: GPU call {dt * 1e6:9.1f} us ({n / dt / 1e6:8.2f} M strings/s)   CPU oracle, 1 thread {dc * 1e6:9.1f} us", flush=True)
