#!/usr/bin/env python3
"""bench.py - the reference's headline metric on B200: input GB/s scanned (+ matches/s) by the DFA match
hot path, next to the CPU restatement of the reference's generated loops on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-extras]

Default workload = the configuration BASELINE.json's target sentence is quoted on: the 256-state DFA `a[ab]{7}c`,
find() over 8 GiB of batched haystacks (2^27 lines of 64 bytes) per GPU.  A "step" is LAUNCHES_PER_STEP = 8 passes of
the hot path over that resident batch (one launch streams 8 GiB, 65 x the 126 MB L2, so nothing is served from cache;
eight launches make a step about 14 ms of device work and the default 20 steps a sustained region of about 0.3 s).
N > 1 is launched by torchrun, one rank per GPU; the regex is compiled on rank 0 and its table blob is NCCL-broadcast;
batches are sharded by rank (weak scaling: each GPU scans its own 8 GiB) with no data-path collective.

`value`  whole-job GB/s with inputs resident in HBM (device pointers through ndl_match_batch), CUDA events.
`e2e`    same metric through ndl_match_batch with HOST (pinned) buffers: H2D of the haystacks (+ offsets when they are
         not equally spaced) and D2H of the results inside the timed region; `link_peak` is a raw concurrent
         cudaMemcpyAsync H2D + D2H of the same bytes at the same N, `frac` = e2e / link_peak.
`extra`  (N = 1) the other BASELINE configs - c2, c3, c5, c4long - measured the same way, shorter; (N > 1) the single
         8 GiB haystack of configs[3] split across the ranks (strong scaling).
`--impl reference`  the reference arm: needle's own matcher is JVM bytecode generated at run time and no JVM exists on
         this image, so it times the C restatement of the generated loops (oracle/) on all host cores, on the same
         workload - both the plain form and the form with the reference's indexOf / first-byte accelerators
         (kind "port", variant "port+accelerators"; `value` is the accelerated one, what the JVM would run).
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
LAUNCHES_PER_STEP = 8

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
DEFAULT_WORKLOAD = "c4b8g"
# workloads above this many lines are generated as one host block of this size, repeated on the device
BLOCK_LINES = 1 << 24
# sizes of the N = 1 `extra` sub-records (ragged / planted generators loop in Python: keep them short)
EXTRA_LINES = {"c2": 10_000_000, "c3": 4_000_000, "c5": 10_000_000}  # c5 at 4 M lines is a 0.05 ms launch: launch overhead shows


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


def workload_config(name, n, in_bytes, launches_per_step):
    """The `config` object: identical in both arms (the driver compares them)."""
    key, desc, _, _, cw = WORKLOADS[name]
    return {"workload": desc, "regex": workloads.REGEX[key], "mode": "find", "char_width": cw, "lines_per_gpu": n,
            "bytes_per_gpu_per_launch": in_bytes, "launches_per_step": launches_per_step,
            "l2": f"every launch streams {in_bytes / 1e6:.0f} MB, larger than the 126 MB L2; no flush needed",
            "sharding": "contiguous line ranges per rank, table blob NCCL-broadcast once, no data-path collective"}


def host_block(name, n, rank):
    """This rank's lines on the host: (data uint8, offsets uint64, n_host, reps).  Workloads above BLOCK_LINES are one
    host block repeated `reps` times on the device (fixed-length workloads only)."""
    key, _, gen, _, _ = WORKLOADS[name]
    reps, n_host = 1, n
    if n > BLOCK_LINES:
        assert n % BLOCK_LINES == 0
        reps, n_host = n // BLOCK_LINES, BLOCK_LINES
    data_h, off_h = gen(n_host, seed=0x5EED0000 + 16 * rank + int(key[1]))
    return np.ascontiguousarray(data_h).view(np.uint8), off_h, n_host, reps


# ---------------------------------------------------------------------------------------------------------------
# the reference arm
# ---------------------------------------------------------------------------------------------------------------
def time_oracle(ora, data, offsets, cw, threads, passes, accelerated):
    t0 = time.perf_counter()
    m = None
    for _ in range(passes):
        m, _, _ = ora.match_batch(2, data, offsets, cw, threads=threads, accelerated=accelerated)
    return time.perf_counter() - t0, int(m.sum())


def run_reference(args, rank, world):
    """The reference arm: the C restatement of the reference's generated loops on the host cores - with the accelerators
    the reference's CompilationPolicy picks for this regex (the number reported) and without them."""
    if rank != 0:
        return
    import needle_b200 as nb
    from tests.oracle_lib import Oracle

    name = "c4b" if args.workload == "c4long" else args.workload
    key, desc, gen, default_lines, cw = WORKLOADS[name]
    regex = workloads.REGEX[key]
    n = args.lines or default_lines
    ora = Oracle(nb.compile_to_bytes(regex, 0))
    threads = host_threads()
    data, offsets, n_host, reps = host_block(name, n, 0)
    # a step scans a bounded sample of the workload: the host block (at most 2^24 lines), once
    in_bytes_launch = int(offsets[-1] - offsets[0]) * cw * reps
    sample_bytes = int(offsets[-1] - offsets[0]) * cw
    for _ in range(args.warmup):
        ora.match_batch(2, data, offsets, cw, threads=threads, accelerated=True)
    dt, matches = time_oracle(ora, data, offsets, cw, threads, args.steps, True)
    gbs = sample_bytes * args.steps / dt / 1e9
    plain_steps = max(1, min(args.steps, 5))
    dtp, _ = time_oracle(ora, data, offsets, cw, threads, plain_steps, False)
    plain_gbs = sample_bytes * plain_steps / dtp / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(name, n, in_bytes_launch, LAUNCHES_PER_STEP if name == DEFAULT_WORKLOAD else 1),
        "matches_per_s": matches * args.steps / dt,
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "variant": "port+accelerators", "plain_value": plain_gbs,
                         "accelerators": ora.accel_summary(),
                         "sample": f"{n_host} lines ({sample_bytes / 1e6:.0f} MB) of the workload per step, all {threads} host threads; no JVM on "
                                   "this image: C restatement of needle's generated Matcher loops (oracle/), find() with the "
                                   "indexOf / first-byte accelerators CompilationPolicy picks; plain_value = the same loops without them"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the timed loop is oracle/libneedle_oracle.so (ndlo_match_batch); libneedle_b200.so is mapped only to compile the regex to tables",
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


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Bench:
    """One rank's state: device, stream, process group."""

    def __init__(self, rank, local_rank, world):
        import torch
        self.torch = torch
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.current_stream()

    def sync_all(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def pattern(self, regex):
        """regex -> table blob on rank 0; ONE NCCL broadcast of the blob; every rank uploads its own device image."""
        import needle_b200 as nb
        blob = nb.compile_to_bytes(regex, 0) if self.rank == 0 else None
        if self.dist:
            from needle_b200.sharding import broadcast_blob
            blob = broadcast_blob(blob, src=0, device=self.dev)
        return nb.Pattern(blob, device=self.local_rank), blob


def kernel_name(pat, cw):
    import ctypes

    from needle_b200 import _lib
    L = _lib.lib()
    L.ndl_debug_kernel_name.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.ndl_debug_kernel_name.restype = ctypes.c_char_p
    return L.ndl_debug_kernel_name(pat._h, 2, cw).decode()


def measured_traffic(workload, n):
    """DRAM bytes per launch of the bench kernel from the committed ncu capture of the same kernel at the same size:
    (bytes, source) - a static figure, not measured in this run (ncu cannot run inside the timed bench)."""
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                t = json.load(f)[workload]
            if int(t.get("lines", n)) != n:
                continue
            return t["dram_bytes_read"] + t["dram_bytes_write"], f"static: profiles/{fn} (ncu capture of this kernel at this size: {t.get('source', 'same workload and size')})"
        except Exception:
            continue
    return None, None


def measure_batch(b, name, n, steps, warmup, launches_per_step, with_e2e=True, with_stride=True):
    """Device-resident and end-to-end timing of find() over one batched workload on this rank; returns the per-rank
    measurements (reduced over ranks by the caller)."""
    torch = b.torch
    import needle_b200 as nb
    from needle_b200 import _lib
    from tests.oracle_lib import Oracle

    key, desc, gen, _, cw = WORKLOADS[name]
    pat, blob = b.pattern(workloads.REGEX[key])
    data_h, off_h, n_host, reps = host_block(name, n, b.rank)
    data_p = torch.from_numpy(data_h).pin_memory()
    off_p = torch.from_numpy(off_h.view(np.int64)).pin_memory()
    data_d, off_d = data_p.to(b.dev), off_p.to(b.dev)
    if reps > 1:
        line_chars = int(off_h[1] - off_h[0])
        data_d = data_d.repeat(reps)
        off_d = torch.arange(n + 1, dtype=torch.int64, device=b.dev) * line_chars
    in_bytes = int(off_h[-1] - off_h[0]) * cw * reps
    matched_d = torch.zeros(n, dtype=torch.uint8, device=b.dev)
    start_d = torch.zeros(n, dtype=torch.int32, device=b.dev)
    end_d = torch.zeros(n, dtype=torch.int32, device=b.dev)
    stream = b.stream

    def step_device():
        for _ in range(launches_per_step):
            pat.match_batch_ptrs(nb.MODE_FIND, data_d.data_ptr(), off_d.data_ptr(), n, cw, matched_d.data_ptr(), start_d.data_ptr(),
                                 end_d.data_ptr(), stream=stream.cuda_stream)

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()

    # parity guard (untimed): EVERY line of the batch against the oracle (the host block on all host threads; for a
    # repeated block every repetition must reproduce the block's results).  A mismatch voids the run.
    em, es, ee = Oracle(blob).match_batch(2, data_h, off_h, cw, threads=host_threads())
    em_d, es_d, ee_d = (torch.from_numpy(x).to(b.dev) for x in (em, es, ee))
    for r in range(reps):
        sl = slice(r * n_host, (r + 1) * n_host)
        if not (torch.equal(matched_d[sl], em_d) and torch.equal(start_d[sl], es_d) and torch.equal(end_d[sl], ee_d)):
            raise SystemExit(f"bench[{name}]: GPU results differ from the oracle - refusing to report a number")
    del em_d, es_d, ee_d

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(b.local_rank)
    launches0 = _lib.lib().ndl_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.sync_all()
    sampler.start()
    ev0.record(stream)
    for _ in range(steps):
        step_device()
    ev1.record(stream)
    b.sync_all()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.lib().ndl_kernel_launches() - launches0
    n_matches = int(matched_d.sum().item())
    out = {"name": name, "desc": desc, "n": n, "n_host": n_host, "reps": reps, "cw": cw, "in_bytes": in_bytes, "ms": ms, "launches": launches,
           "n_matches": n_matches, "clocks": clocks, "kernel": kernel_name(pat, cw), "blob": blob, "data_h": data_h, "off_h": off_h,
           "steps": steps, "launches_per_step": launches_per_step}

    # ---- the same batch through ndl_match_lines (fixed-length records: no offsets array to read), reported beside `value`
    fixed_len = bool(np.all(np.diff(off_h[:1001]) == off_h[1] - off_h[0])) and name != "c3"
    out["fixed_len"] = fixed_len
    if fixed_len and with_stride:
        line_chars = int(off_h[1] - off_h[0])
        m2, s2, e2 = torch.zeros_like(matched_d), torch.zeros_like(start_d), torch.zeros_like(end_d)

        def step_stride():
            for _ in range(launches_per_step):
                pat.match_lines_ptrs(nb.MODE_FIND, data_d.data_ptr(), n, line_chars, cw, m2.data_ptr(), s2.data_ptr(), e2.data_ptr(),
                                     stream=stream.cuda_stream)
        for _ in range(min(warmup, 2)):
            step_stride()
        b.sync_all()
        if not (torch.equal(m2, matched_d) and torch.equal(s2, start_d) and torch.equal(e2, end_d)):
            raise SystemExit(f"bench[{name}]: ndl_match_lines and ndl_match_batch disagree - refusing to report a number")
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(steps):
            step_stride()
        a1.record(stream)
        b.sync_all()
        out["stride_ms"] = a0.elapsed_time(a1)
        del m2, s2, e2

    if with_e2e:
        # ---- timed region 2: end to end from pinned host buffers through the same C-ABI call (the host-resident block)
        matched_h = torch.zeros(n_host, dtype=torch.uint8).pin_memory()
        start_h = torch.zeros(n_host, dtype=torch.int32).pin_memory()
        end_h = torch.zeros(n_host, dtype=torch.int32).pin_memory()
        e2e_steps = max(1, min(steps, 5))
        e2e_bytes = in_bytes // reps

        def step_host():
            pat.match_batch_ptrs(nb.MODE_FIND, data_p.data_ptr(), off_p.data_ptr(), n_host, cw, matched_h.data_ptr(), start_h.data_ptr(),
                                 end_h.data_ptr(), mem_kind=nb.MEM_HOST, stream=stream.cuda_stream)

        step_host()
        b.sync_all()
        e2e_ms = None
        for _ in range(2):  # two timed repetitions, the faster one counts: freshly pinned host pages settle during the first
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(e2e_steps):
                step_host()
            e1.record(stream)
            b.sync_all()
            rep_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
            e2e_ms = rep_ms if e2e_ms is None else min(e2e_ms, rep_ms)
        out["e2e_ms"] = e2e_ms
        out["e2e_steps"], out["e2e_bytes"] = e2e_steps, e2e_bytes
        if not (np.array_equal(matched_h.numpy(), em) and np.array_equal(start_h.numpy(), es) and np.array_equal(end_h.numpy(), ee)):
            raise SystemExit(f"bench[{name}]: host-buffer path differs from the oracle - refusing to report a number")
        out["h2d_bytes"] = int(e2e_bytes + (0 if fixed_len else off_h.nbytes))
        out["d2h_bytes"] = int(9 * n_host)
        # ---- the link itself: the same bytes as raw concurrent cudaMemcpyAsync H2D + D2H, all ranks at once
        s2 = torch.cuda.Stream()
        res_d = torch.zeros(9 * n_host, dtype=torch.uint8, device=b.dev)
        res_h = torch.zeros(9 * n_host, dtype=torch.uint8).pin_memory()
        dst = data_d[:data_p.numel()]

        def raw_copy():
            dst.copy_(data_p, non_blocking=True)
            with torch.cuda.stream(s2):
                res_h.copy_(res_d, non_blocking=True)
        raw_copy()
        b.sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            raw_copy()
        b.sync_all()
        out["link_ms"] = (time.perf_counter() - t0) * 1e3
        del res_d, res_h, matched_h, start_h, end_h
    del data_d, off_d, matched_d, start_d, end_d, data_p, off_p, pat
    torch.cuda.empty_cache()
    return out


def summarise_batch(b, m, peak, peak_src, headline):
    """Reduce one workload's measurements over the ranks and build its record (rank 0 prints it)."""
    world = b.world
    steps, lps = m["steps"], m["launches_per_step"]
    ms, e2e_ms, stride_ms, link_ms = b.reduce([m["ms"], m.get("e2e_ms", 0.0), m.get("stride_ms", 0.0), m.get("link_ms", 0.0)], "MAX")
    job_bytes, job_matches, job_launches, job_e2e_bytes = b.reduce(
        [float(m["in_bytes"]), float(m["n_matches"]), float(m["launches"]), float(m.get("e2e_bytes", 0))], "SUM")
    value = job_bytes * steps * lps / (ms * 1e-3) / 1e9
    kernel_ms = ms / (steps * lps)  # one kernel launch per pass
    achieved = m["in_bytes"] / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(m["name"], m["n"])
    rec = {
        "value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "matches_per_s": job_matches * steps * lps / (ms * 1e-3),
        "gpu_launches": int(job_launches), "clocks": m["clocks"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": m["kernel"], "kernel_ms": kernel_ms, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": m["in_bytes"],
                     "achieved_with_metadata": (m["in_bytes"] + 17 * m["n"]) / (kernel_ms * 1e-3) / 1e9,
                     "frac_with_metadata": (m["in_bytes"] + 17 * m["n"]) / (kernel_ms * 1e-3) / 1e9 / peak,
                     "traffic_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                     "frac_of_nominal_8tbs": achieved / 8000.0,
                     "note": "algorithmic bytes = haystack bytes only (SURVEY.md 8d); achieved_with_metadata adds the 8 B/line of offsets the "
                             "launch must read and the 9 B/line of results it must write (the API's own traffic, also HBM-bound); traffic_gbs = the ncu-measured "
                             "DRAM bytes per launch (static, see traffic_source) over this run's launch time"},
    }
    if "e2e_ms" in m:
        e2e_value = job_e2e_bytes * m["e2e_steps"] / (e2e_ms * 1e-3) / 1e9
        link = job_e2e_bytes * m["e2e_steps"] / (link_ms * 1e-3) / 1e9
        rec["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
                      "steps": m["e2e_steps"], "ms_per_step": e2e_ms / m["e2e_steps"],
                      "repetitions": "the timed steps run twice after one warm-up call; the faster repetition is reported",
                      "link_peak": link, "frac": e2e_value / link,
                      "link_peak_note": "raw concurrent cudaMemcpyAsync of the same H2D + D2H bytes from / to pinned memory, all ranks at once",
                      **({"note": f"host path timed on the host-resident block of {m['n_host']} lines; one step = one call"} if m["reps"] > 1 else {})}
    if m.get("stride_ms"):
        sv = job_bytes * steps * lps / (stride_ms * 1e-3) / 1e9
        rec["stride_api"] = {"call": "ndl_match_lines (fixed-length records, no offsets array)", "value": sv, "unit": UNIT,
                             "roofline_frac": sv / world / peak}
    if not headline:
        rec["config"] = {"workload": m["desc"], "lines_per_gpu": m["n"], "bytes_per_gpu_per_launch": m["in_bytes"], "launches_per_step": lps}
    return rec


def run_ours(args, rank, local_rank, world):
    b = Bench(rank, local_rank, world)
    peak, peak_src = measured_peak_gbs()
    name = args.workload
    n = args.lines or WORKLOADS[name][3]
    lps = args.launches_per_step or (LAUNCHES_PER_STEP if name == DEFAULT_WORKLOAD and not args.lines else 1)
    m = measure_batch(b, name, n, args.steps, args.warmup, lps)
    rec = summarise_batch(b, m, peak, peak_src, headline=True)
    line = {"metric": METRIC, "value": rec.pop("value"), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(name, n, m["in_bytes"], lps)}
    rec.pop("unit"), rec.pop("steps")
    line.update(rec)
    if b.numa_cpus:
        line["host_binding"] = f"each rank bound to the {b.numa_cpus} CPUs local to its GPU (NVML affinity)"
    extra = {}
    if not args.no_extras and name == DEFAULT_WORKLOAD and not args.lines:
        blob, data_h, off_h, cw = m["blob"], m["data_h"], m["off_h"], m["cw"]
        m = None
        if world == 1:
            # the other BASELINE configs, shorter: device-resident + e2e, same parity guard
            for ex, ex_lines in EXTRA_LINES.items():
                ex_steps = max(3, min(args.steps, 10))
                em = measure_batch(b, ex, ex_lines, ex_steps, min(args.warmup, 3), 8, with_stride=False)
                extra[ex] = summarise_batch(b, em, peak, peak_src, headline=False)
                del em
            extra["c4long"] = measure_long(b, 8 << 30, max(3, min(args.steps, 10)), min(args.warmup, 3))
            line["cpu_baseline"] = cpu_baseline(blob, data_h, off_h, cw)
        else:
            # strong scaling of the single 8 GiB haystack of BASELINE configs[3], split across the ranks
            extra["c4long"] = measure_long(b, 8 << 30, max(3, min(args.steps, 10)), min(args.warmup, 3))
    elif world == 1:
        line["cpu_baseline"] = cpu_baseline(m["blob"], m["data_h"], m["off_h"], m["cw"])
    if extra:
        line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if b.dist:
        b.dist.destroy_process_group()


def measure_long(b, n_total, steps, warmup):
    """BASELINE configs[3]: 256-state DFA over ONE haystack of n_total bytes, find() start/end offsets.  The buffer is {a,b}
    noise with the only match in its last 9 bytes, so the whole buffer must be scanned.  N = 1: ndl_find_long.
    N > 1: the haystack is split into N contiguous chunks, one per GPU (strong scaling: the same bytes in total),
    and needle_b200.sharding.find_long_sharded runs the entry-state guess / all-gather / verify protocol over
    ndl_find_long_from (one 40-byte NCCL all-gather per step on the data path).  Returns the record (all ranks)."""
    torch = b.torch
    import needle_b200 as nb
    from needle_b200 import _lib
    from needle_b200 import sharding

    rank, world, dev, dist = b.rank, b.world, b.dev, b.dist
    regex = workloads.REGEX["c4"]
    base, top = n_total * rank // world, n_total * (rank + 1) // world
    n = top - base
    pat, _ = b.pattern(regex)
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0004 + rank)
    data = torch.randint(ord("a"), ord("b") + 1, (n,), dtype=torch.uint8, device="cuda", generator=g)
    if rank == world - 1:
        data[n - 9:] = torch.tensor(list(b"abababbac"), dtype=torch.uint8, device="cuda")
    stream = b.stream
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
                lambda: pat.walk_host(halo),
                base, n, rank, world, allgather, fd, bd, pat.reverse_mode, pat.min_length, pat.backwards_root_accepting)

    res = None
    for _ in range(max(1, warmup)):
        res = step()
    if res != want:
        raise SystemExit(f"bench: find over the long haystack returned {res}, expected {want}")
    sampler = ClockSampler(b.local_rank)
    launches0 = _lib.lib().ndl_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.sync_all()
    sampler.start()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    b.sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    ms = max(ev0.elapsed_time(ev1), wall_ms if world > 1 else 0.0)  # the multi-rank step has host-side hand-overs
    launches = _lib.lib().ndl_kernel_launches() - launches0
    # end to end from pinned host memory on a (at most) 1 GiB buffer per rank (same content law)
    ne = min(n, 1 << 30)
    e2e_steps = 3
    host = data[n - ne:].cpu().pin_memory()
    if world == 1:
        def step_host():
            return pat.find_long_ptrs(host.data_ptr(), ne, 1, 0, nb.MEM_HOST, stream.cuda_stream)
    else:
        def step_host():  # the same protocol over host-resident chunks of ne bytes per rank
            return sharding.find_long_sharded(
                lambda entry: pat.find_long_from(host.data_ptr(), ne, entry, mem_kind=nb.MEM_HOST, stream=stream.cuda_stream),
                lambda index, entry, li: pat.find_long_back(host.data_ptr(), ne, index, entry, li, mem_kind=nb.MEM_HOST, stream=stream.cuda_stream),
                lambda: pat.walk_host(halo),
                rank * ne, ne, rank, world, allgather, fd, bd, pat.reverse_mode, pat.min_length, pat.backwards_root_accepting)
    step_host()
    b.sync_all()
    t0 = time.perf_counter()
    r = None
    for _ in range(e2e_steps):
        r = step_host()
    b.sync_all()
    e2e_s = time.perf_counter() - t0
    assert r == (True, world * ne - 9, world * ne), r
    ms, e2e_ms = b.reduce([ms, e2e_s * 1e3], "MAX")
    (job_launches,) = b.reduce([float(launches)], "SUM")
    peak, peak_src = measured_peak_gbs()
    value = n_total * steps / (ms * 1e-3) / 1e9
    del data, host
    torch.cuda.empty_cache()
    return {
        "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "scaling": "strong" if world > 1 else "n/a",
        "config": {"workload": "BASELINE configs[3]: 256-state DFA 'a[ab]{7}c', ONE haystack, find() start/end (int64), match in the last 9 bytes",
                   "regex": regex, "mode": "find", "haystack_bytes": n_total, "l2": "haystack larger than L2",
                   "sharding": "one contiguous chunk per rank; entry-state guess from a 16-byte halo, one all-gather of (entry, end, exit) per step"},
        "matches_per_s": steps / (ms * 1e-3),
        "gpu_launches": int(job_launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": value / world, "peak": peak, "unit": "GB/s", "frac": value / world / peak, "traffic": None,
                     "kernel": "long8_kernel (+ head/tail/seam helper kernels inside the timed call)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": n},
        "e2e": {"value": world * ne * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": world * ne,
                "d2h_bytes_per_step": 17 * world, "steps": e2e_steps, "note": f"host path measured on a {ne >> 20} MiB haystack per rank"},
    }


def run_long(args, rank, local_rank, world):
    b = Bench(rank, local_rank, world)
    rec = measure_long(b, args.lines or (8 << 30), args.steps, args.warmup)
    if rank == 0:
        line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic"}
        for k in ("config", "matches_per_s", "gpu_launches", "clocks", "roofline", "e2e"):
            line[k] = rec[k]
        print(json.dumps(line), flush=True)
    if b.dist:
        b.dist.destroy_process_group()


def cpu_baseline(blob, data, offsets, cw):
    """The oracle port timed on this box's host cores on a bounded sample of the same workload (about 10 s of CPU work in
    total): with the accelerators the reference's CompilationPolicy picks (the headline `value` - what the JVM would run)
    and without them (`plain_value`)."""
    from tests.oracle_lib import Oracle
    ora = Oracle(blob)
    threads = host_threads()
    n = min(len(offsets) - 1, 4_000_000)
    o = offsets[:n + 1]
    nbytes = int(o[n] - o[0]) * cw

    def timed(accelerated, budget_s, nthreads, lines):
        oo = offsets[:lines + 1]
        t0 = time.perf_counter()
        ora.match_batch(2, data, oo, cw, threads=nthreads, accelerated=accelerated)
        calib = time.perf_counter() - t0
        reps = int(max(1, min(200, budget_s / max(calib, 1e-3))))
        dt, _ = time_oracle(ora, data, oo, cw, nthreads, reps, accelerated)
        return int(oo[lines] - oo[0]) * cw * reps / dt / 1e9, reps

    accel, reps_a = timed(True, 4.0, threads, n)
    plain, reps_p = timed(False, 4.0, threads, n)
    single, _ = timed(True, 1.0, 1, max(1, n // 10))
    return {"value": accel, "unit": UNIT, "cores": threads, "kind": "port", "variant": "port+accelerators", "plain_value": plain, "single_thread_value": single,
            "accelerators": ora.accel_summary(),
            "sample": f"first {n} lines ({nbytes / 1e6:.0f} MB) x {reps_a} passes (accelerated) / {reps_p} passes (plain), {threads} threads; "
                      "C restatement of the generated loops, not the JVM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + ["c4long"])
    ap.add_argument("--lines", type=int, default=0, help="lines per GPU (default: the workload's full size)")
    ap.add_argument("--launches-per-step", type=int, default=0, help="passes over the batch per step (default: 8 for the default workload, else 1)")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` sub-records of the default run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "c4long":
        run_long(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
