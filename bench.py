#!/usr/bin/env python3
"""bench.py - the reference's headline metric on B200: input GB/s scanned (+ matches/s) by the DFA match
hot path, next to the CPU restatement of the reference's generated loops on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4b|c3]

A "step" is one pass of find() over one batch of synthetic haystacks.  Default workload = BASELINE.json
configs[1]: `\\d{3}-\\d{2}-\\d{4}` over 10 M synthetic 64-byte ASCII lines per GPU (640 MB, larger than
the 126 MB L2, so every step streams from HBM).  N > 1 is launched by torchrun, one rank per GPU; the
regex is compiled on rank 0 and its table blob is NCCL-broadcast; haystack batches are sharded by rank
(weak scaling: each GPU scans its own 10 M lines) with no data-path collective.

`value`  whole-job GB/s with inputs resident in HBM (device pointers through ndl_match_batch).
`e2e`    same metric through ndl_match_batch with HOST (pinned) buffers: H2D of data+offsets and D2H of
         the results inside the timed region.
`--impl reference`  the reference arm: needle's own code is JVM bytecode generated at run time and no JVM
         exists on this image, so it times the C restatement of the generated loops (oracle/, kind "port")
         on all host cores, on the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import workloads  # noqa: E402

METRIC = "input_gb_per_s_scanned"
UNIT = "GB/s"

WORKLOADS = {
    # name: (regex key, description, generator, default lines per GPU, char width)
    "c2": ("c2", "BASELINE configs[1]: '\\d{3}-\\d{2}-\\d{4}' find() over 10M synthetic 64-byte ASCII lines per GPU", workloads.c2_lines, 10_000_000, 1),
    "c4b": ("c4", "BASELINE configs[3] batched variant: 256-state DFA 'a[ab]{7}c' find() over 64-byte lines of {a,b}", workloads.c4_lines, 10_000_000, 1),
    "c3": ("c3", "BASELINE configs[2]: email-like regex find() over mixed-length lines (8..120 B)", workloads.c3_lines, 10_000_000, 1),
    "c5": ("c5", "BASELINE configs[4]: BMP char-class regex find() over UTF-16LE lines of 32 chars (64 B)", workloads.c5_lines, 10_000_000, 2),
    "c2w": ("c2", "BASELINE configs[1] regex over UTF-16LE lines of 32 chars (64 B), i.e. java.lang.String payloads", workloads.c2_lines_utf16, 10_000_000, 2),
    # the configuration of BASELINE.json's target sentence: 8 GiB of batched haystacks, 256-state DFA
    "c4b8g": ("c4", "BASELINE north-star target: 256-state DFA 'a[ab]{7}c' find() over 8 GiB of batched 64-byte lines of {a,b} (2^27 lines)", workloads.c4_lines, 1 << 27, 1),
}
# workloads above this many lines are generated as one host block of this size, repeated on the device
BLOCK_LINES = 1 << 24


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """The reference arm: the C restatement of the reference's generated loops on the host cores."""
    if rank != 0:
        return
    import needle_b200 as nb
    from tests.oracle_lib import Oracle

    key, desc, gen, default_lines, cw = WORKLOADS[args.workload]
    regex = workloads.REGEX[key]
    ora = Oracle(nb.compile_to_bytes(regex, 0))
    threads = host_threads()
    n_sample = min(args.lines or default_lines, 2_000_000)
    data, offsets = gen(n_sample)
    in_bytes = int(offsets[-1] - offsets[0]) * cw
    for _ in range(args.warmup):
        ora.match_batch(2, data, offsets, cw, threads=threads)
    t0 = time.perf_counter()
    matches = 0
    for _ in range(args.steps):
        m, _, _ = ora.match_batch(2, data, offsets, cw, threads=threads)
        matches = int(m.sum())
    dt = time.perf_counter() - t0
    gbs = in_bytes * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "regex": regex, "mode": "find", "lines_per_step": n_sample, "bytes_per_step": in_bytes,
                   "note": "no JVM on this image: C restatement of needle's generated Matcher loops (oracle/), all host threads"},
        "matches_per_s": matches * args.steps / dt,
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_sample} lines ({in_bytes / 1e6:.0f} MB) of the workload per step"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: keep this rank (and the pinned buffers it allocates next) on the CPUs NVML reports as local to its
    GPU, so that eight host->device streams do not all pull from one socket's memory.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def run_ours(args, rank, local_rank, world):
    import torch

    import needle_b200 as nb
    from needle_b200 import _lib

    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    key, desc, gen, default_lines, cw = WORKLOADS[args.workload]
    regex = workloads.REGEX[key]
    n = args.lines or default_lines

    # regex -> table blob on rank 0; ONE NCCL broadcast of the blob; every rank uploads its own device image
    blob = nb.compile_to_bytes(regex, 0) if rank == 0 else None
    if dist:
        from needle_b200.sharding import broadcast_blob
        blob = broadcast_blob(blob, src=0, device=dev)
    pat = nb.Pattern(blob, device=local_rank)

    # this rank's shard of the job: its own n lines (weak scaling), seeded by rank
    reps = 1
    if n > BLOCK_LINES:  # fixed-length workloads only: one host block, repeated on the device
        assert n % BLOCK_LINES == 0
        reps, n_host = n // BLOCK_LINES, BLOCK_LINES
    else:
        n_host = n
    data_h, off_h = gen(n_host, seed=0x5EED0000 + 16 * rank + int(key[1]))
    data_h = np.ascontiguousarray(data_h).view(np.uint8)
    pin = torch.cuda.is_available()
    data_p = torch.from_numpy(data_h).pin_memory() if pin else torch.from_numpy(data_h)
    off_p = torch.from_numpy(off_h.view(np.int64)).pin_memory()
    data_d, off_d = data_p.to(dev), off_p.to(dev)
    if reps > 1:
        line_chars = int(off_h[1] - off_h[0])
        data_d = data_d.repeat(reps)
        off_d = torch.arange(n + 1, dtype=torch.int64, device=dev) * line_chars
    in_bytes = int(off_h[-1] - off_h[0]) * cw * reps
    matched_d = torch.zeros(n, dtype=torch.uint8, device=dev)
    start_d = torch.zeros(n, dtype=torch.int32, device=dev)
    end_d = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        pat.match_batch_ptrs(nb.MODE_FIND, data_d.data_ptr(), off_d.data_ptr(), n, cw, matched_d.data_ptr(), start_d.data_ptr(),
                             end_d.data_ptr(), stream=stream.cuda_stream)

    def sync_all():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()

    # parity guard (untimed): a sample of the batch against the oracle; a mismatch voids the run
    from tests.oracle_lib import Oracle
    ns = min(n_host, 200_000)
    em, es, ee = Oracle(blob).match_batch(2, data_h, off_h[:ns + 1], cw, threads=host_threads())
    if not (np.array_equal(matched_d[:ns].cpu().numpy(), em) and np.array_equal(start_d[:ns].cpu().numpy(), es)
            and np.array_equal(end_d[:ns].cpu().numpy(), ee)):
        raise SystemExit("bench: GPU results differ from the oracle - refusing to report a number")
    if reps > 1:  # size-independent property at full size: every repetition of the block gives the block's results
        m0, s0, e0 = matched_d[:n_host], start_d[:n_host], end_d[:n_host]
        for r in range(1, reps):
            sl = slice(r * n_host, (r + 1) * n_host)
            if not (torch.equal(matched_d[sl], m0) and torch.equal(start_d[sl], s0) and torch.equal(end_d[sl], e0)):
                raise SystemExit("bench: repeated blocks give different results - refusing to report a number")

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    launches0 = _lib.lib().ndl_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    sampler.start()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    sync_all()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.lib().ndl_kernel_launches() - launches0
    n_matches = int(matched_d.sum().item())

    # ---- the same batch through ndl_match_lines (fixed-length records: no offsets array to read), reported beside `value`
    stride_ms = stride_e2e_ms = None
    fixed_len = bool(np.all(np.diff(off_h[:1001]) == off_h[1] - off_h[0])) and args.workload != "c3"
    if fixed_len:
        line_chars = int(off_h[1] - off_h[0])
        m2, s2, e2 = torch.zeros_like(matched_d), torch.zeros_like(start_d), torch.zeros_like(end_d)

        def step_stride():
            pat.match_lines_ptrs(nb.MODE_FIND, data_d.data_ptr(), n, line_chars, cw, m2.data_ptr(), s2.data_ptr(), e2.data_ptr(),
                                 stream=stream.cuda_stream)
        for _ in range(args.warmup):
            step_stride()
        sync_all()
        if not (torch.equal(m2, matched_d) and torch.equal(s2, start_d) and torch.equal(e2, end_d)):
            raise SystemExit("bench: ndl_match_lines and ndl_match_batch disagree - refusing to report a number")
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(args.steps):
            step_stride()
        a1.record(stream)
        sync_all()
        stride_ms = a0.elapsed_time(a1)
        del m2, s2, e2

    # ---- timed region 2: end to end from pinned host buffers through the same C-ABI call
    matched_h = torch.zeros(n_host, dtype=torch.uint8).pin_memory()
    start_h = torch.zeros(n_host, dtype=torch.int32).pin_memory()
    end_h = torch.zeros(n_host, dtype=torch.int32).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    e2e_bytes = in_bytes // reps  # the host path is timed on the host-resident block

    def step_host():
        pat.match_batch_ptrs(nb.MODE_FIND, data_p.data_ptr(), off_p.data_ptr(), n_host, cw, matched_h.data_ptr(), start_h.data_ptr(),
                             end_h.data_ptr(), mem_kind=nb.MEM_HOST, stream=stream.cuda_stream)

    step_host()
    sync_all()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        step_host()
    e1.record(stream)
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    assert np.array_equal(matched_h[:ns].numpy(), em)
    if fixed_len:
        def step_host_stride():
            pat.match_lines_ptrs(nb.MODE_FIND, data_p.data_ptr(), n_host, line_chars, cw, matched_h.data_ptr(), start_h.data_ptr(),
                                 end_h.data_ptr(), mem_kind=nb.MEM_HOST, stream=stream.cuda_stream)
        step_host_stride()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host_stride()
        sync_all()
        stride_e2e_ms = (time.perf_counter() - t0) * 1e3
        assert np.array_equal(matched_h[:ns].numpy(), em)

    # max over ranks, sum of bytes over ranks
    t = torch.tensor([ms, e2e_ms, stride_ms or 0.0, stride_e2e_ms or 0.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(in_bytes), float(n_matches), float(launches), float(e2e_bytes)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, e2e_ms, stride_ms, stride_e2e_ms = t.tolist()
    job_bytes, job_matches, job_launches, job_e2e_bytes = tot.tolist()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        value = job_bytes * args.steps / (ms * 1e-3) / 1e9
        e2e_value = job_e2e_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9
        kernel_ms = ms / args.steps  # one kernel launch per step
        achieved = in_bytes / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": desc, "regex": regex, "mode": "find", "lines_per_gpu": n, "bytes_per_gpu_per_step": in_bytes,
                       "l2": f"inputs ({in_bytes / 1e6:.0f} MB per step) are larger than the 126 MB L2; no flush needed",
                       "sharding": "contiguous line ranges per rank, table blob NCCL-broadcast once, no data-path collective",
                       **({"host_binding": f"each rank bound to the {numa_cpus} CPUs local to its GPU (NVML affinity)"} if numa_cpus else {})},
            "matches_per_s": job_matches * args.steps / (ms * 1e-3),
            # equally spaced offsets are checked on the host (every one of them) and computed on the device instead of copied
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_bytes + (0 if fixed_len else off_h.nbytes)),
                    "d2h_bytes_per_step": int(9 * n_host),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    **({"note": f"host path timed on the host-resident block of {n_host} lines"} if reps > 1 else {})},
            "gpu_launches": int(job_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload) if n == default_lines else None,
                         "kernel": kernel_name(pat, cw), "kernel_ms": kernel_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": in_bytes,
                         "achieved_with_metadata": (in_bytes + 17 * n) / (kernel_ms * 1e-3) / 1e9,
                         "frac_of_nominal_8tbs": achieved / 8000.0,
                         "note": "algorithmic bytes = haystack bytes only (SURVEY.md 8d); achieved_with_metadata adds the 8 B/line of offsets the "
                                 "launch must read and the 9 B/line of results it must write (the API's own traffic, also HBM-bound)"},
        }
        if fixed_len and stride_ms:
            sv = job_bytes * args.steps / (stride_ms * 1e-3) / 1e9
            line["stride_api"] = {
                "call": "ndl_match_lines (fixed-length records, no offsets array)", "value": sv, "unit": UNIT,
                "ms_per_step": stride_ms / args.steps, "roofline_frac": sv / world / peak,
                "achieved_with_metadata": (in_bytes + 9 * n) / (stride_ms / args.steps * 1e-3) / 1e9,
                "e2e": {"value": job_e2e_bytes * e2e_steps / (stride_e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                        "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": int(9 * n_host)}}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(blob, data_h, off_h, cw)
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def run_long(args, rank, local_rank, world):
    """BASELINE configs[3]: 256-state DFA over ONE 8 GiB haystack, find() start/end offsets.  The buffer is {a,b}
    noise with the only match in its last 9 bytes, so the whole buffer must be scanned.  N = 1: ndl_find_long.
    N > 1: the haystack is split into N contiguous chunks, one per GPU (strong scaling: the same 8 GiB in total),
    and needle_b200.sharding.find_long_sharded runs the entry-state guess / all-gather / verify protocol over
    ndl_find_long_from (one 40-byte NCCL all-gather per step on the data path)."""
    import torch

    import needle_b200 as nb
    from needle_b200 import _lib
    from needle_b200 import sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    regex = workloads.REGEX["c4"]
    n_total = args.lines or (8 << 30)
    base, top = n_total * rank // world, n_total * (rank + 1) // world
    n = top - base
    blob = nb.compile_to_bytes(regex, 0) if rank == 0 else None
    if dist:
        blob = sharding.broadcast_blob(blob, src=0, device=dev)
    pat = nb.Pattern(blob, device=local_rank)
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0004 + rank)
    data = torch.randint(ord("a"), ord("b") + 1, (n,), dtype=torch.uint8, device="cuda", generator=g)
    if rank == world - 1:
        data[n - 9:] = torch.tensor(list(b"abababbac"), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    want = (True, n_total - 9, n_total)

    if world == 1:
        def step():
            return pat.find_long_ptrs(data.data_ptr(), n, 1, 0, nb.MEM_DEVICE, stream.cuda_stream)
    else:
        halo = sharding.exchange_halo(data[n - sharding.HALO:], device=dev)  # once: the haystack does not change
        allgather = sharding.tensor_allgather(dev)
        fd, bd = pat.forwards_state_count, pat.backwards_state_count

        def step():
            return sharding.find_long_sharded(
                lambda entry: pat.find_long_from(data.data_ptr(), n, entry, mem_kind=nb.MEM_DEVICE, stream=stream.cuda_stream),
                lambda index, entry, li: pat.find_long_back(data.data_ptr(), n, index, entry, li, mem_kind=nb.MEM_DEVICE, stream=stream.cuda_stream),
                lambda: pat.find_long_from(halo.ctypes.data, halo.size, 0, mem_kind=nb.MEM_HOST, stream=stream.cuda_stream)[1],
                base, n, rank, world, allgather, fd, bd, pat.reverse_mode, pat.min_length, pat.backwards_root_accepting)

    def sync_all():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    res = None
    for _ in range(max(1, args.warmup)):
        res = step()
    if res != want:
        raise SystemExit(f"bench: find over the long haystack returned {res}, expected {want}")
    sampler = ClockSampler(local_rank)
    launches0 = _lib.lib().ndl_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    sampler.start()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    ms = max(ev0.elapsed_time(ev1), wall_ms if world > 1 else 0.0)  # the multi-rank step has host-side hand-overs
    launches = _lib.lib().ndl_kernel_launches() - launches0
    # end to end from pinned host memory on a (at most) 1 GiB buffer per rank (same content law)
    ne = min(n, 1 << 30)
    e2e_steps, e2e_s = 3, None
    host = data[n - ne:].cpu().pin_memory()
    if world == 1:
        def step_host():
            return pat.find_long_ptrs(host.data_ptr(), ne, 1, 0, nb.MEM_HOST, stream.cuda_stream)
    else:
        def step_host():  # the same protocol over host-resident chunks of ne bytes per rank
            return sharding.find_long_sharded(
                lambda entry: pat.find_long_from(host.data_ptr(), ne, entry, mem_kind=nb.MEM_HOST, stream=stream.cuda_stream),
                lambda index, entry, li: pat.find_long_back(host.data_ptr(), ne, index, entry, li, mem_kind=nb.MEM_HOST, stream=stream.cuda_stream),
                lambda: pat.find_long_from(halo.ctypes.data, halo.size, 0, mem_kind=nb.MEM_HOST, stream=stream.cuda_stream)[1],
                rank * ne, ne, rank, world, allgather, fd, bd, pat.reverse_mode, pat.min_length, pat.backwards_root_accepting)
    step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = step_host()
    sync_all()
    e2e_s = time.perf_counter() - t0
    assert r == (True, world * ne - 9, world * ne), r
    if dist:
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        value = n_total * args.steps / (ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: 256-state DFA 'a[ab]{7}c', ONE haystack, find() start/end (int64), match in the last 9 bytes",
                       "regex": regex, "mode": "find", "haystack_bytes": n_total, "l2": "haystack larger than L2",
                       "sharding": "one contiguous chunk per rank; entry-state guess from a 16-byte halo, one all-gather of (entry, end, exit) per step"},
            "matches_per_s": args.steps / (ms * 1e-3),
            "gpu_launches": int(tot.item()), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": value / world, "peak": peak, "unit": "GB/s", "frac": value / world / peak, "traffic": None,
                         "kernel": "long8_kernel (+ head/tail/seam helper kernels inside the timed call)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n},
        }
        if e2e_s is not None:
            line["e2e"] = {"value": world * ne * e2e_steps / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": world * ne,
                           "d2h_bytes_per_step": 17 * world, "steps": e2e_steps,
                           "note": f"host path measured on a {ne >> 20} MiB haystack per rank"}
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def kernel_name(pat, cw):
    import ctypes

    from needle_b200 import _lib
    L = _lib.lib()
    L.ndl_debug_kernel_name.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.ndl_debug_kernel_name.restype = ctypes.c_char_p
    return L.ndl_debug_kernel_name(pat._h, 2, cw).decode()


def measured_traffic(workload):
    """DRAM bytes per launch of the bench kernel, from the committed ncu capture (None if not captured)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)[workload]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def cpu_baseline(blob, data, offsets, cw):
    """The oracle port timed on this box's host cores on a bounded sample of the same workload."""
    from tests.oracle_lib import Oracle
    ora = Oracle(blob)
    threads = host_threads()
    n = min(len(offsets) - 1, 1_000_000)
    t0 = time.perf_counter()
    ora.match_batch(2, data, offsets[:n + 1], cw, threads=threads)
    calib = time.perf_counter() - t0
    reps = int(max(1, min(200, 10.0 / max(calib, 1e-3))))  # ~10 s of CPU work
    t0 = time.perf_counter()
    for _ in range(reps):
        ora.match_batch(2, data, offsets[:n + 1], cw, threads=threads)
    dt = time.perf_counter() - t0
    nbytes = int(offsets[n] - offsets[0]) * cw
    # one thread, on a tenth of the sample: what a single Matcher loop of the reference does
    n1 = max(1, n // 10)
    t0 = time.perf_counter()
    ora.match_batch(2, data, offsets[:n1 + 1], cw, threads=1)
    single = int(offsets[n1] - offsets[0]) * cw / (time.perf_counter() - t0) / 1e9
    return {"value": nbytes * reps / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": single,
            "sample": f"first {n} lines ({nbytes / 1e6:.0f} MB) x {reps} passes, {threads} threads; C restatement of the generated loops, not the JVM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4long"])
    ap.add_argument("--lines", type=int, default=0, help="lines per GPU (default: the workload's full size)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "c4long" and args.impl != "reference":
        run_long(args, rank, local_rank, world)
    elif args.impl == "reference":
        if args.workload == "c4long":
            args.workload = "c4b"
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
