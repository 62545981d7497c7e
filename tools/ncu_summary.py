#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU): key raw metrics + instruction/stall hot spots.
Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout


def run(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run("raw"))))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"== {rep}: kernel {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}", file=out)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:80s} {vals[i]:>16s} {units[i]}", file=out)

rows = list(csv.reader(io.StringIO(run("source"))))
hdr, data = rows[1], rows[2:]
# (a report with several launches repeats the header block per launch: keep the first launch's rows)
for k, r in enumerate(data):
    if len(r) < len(hdr) or r == hdr:
        data = data[:k]
        break
ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ia]) for r in data)
tots = sum(int(r[ist]) for r in data)
agg = collections.Counter()
for r in data:
    for i in stall_cols:
        agg[hdr[i]] += int(r[i] or 0)
print(f"\ntotal warp instructions {tot}, stall samples {tots}", file=out)
print("stall reasons (all samples): " + ", ".join(f"{k}={v / max(1, sum(agg.values())) * 100:.1f}%" for k, v in agg.most_common(8)), file=out)
mix = collections.Counter()
for r in data:
    op = r[isrc].strip().split()[0] if r[isrc].strip() else "?"
    if op.startswith("@"):
        op = r[isrc].strip().split()[1]
    mix[op.split(".")[0]] += int(r[ia])
print("instruction mix: " + ", ".join(f"{k}={v / tot * 100:.1f}%" for k, v in mix.most_common(12)), file=out)
print("\nregions of equal execution count (>0.5% of instructions):", file=out)
prev, start = None, 0
groups = []
for i, r in enumerate(data):
    c = int(r[ia])
    if c != prev:
        if prev is not None:
            groups.append((start, i - 1, prev))
        prev, start = c, i
groups.append((start, len(data) - 1, prev))
for s, e, c in groups:
    n = e - s + 1
    samp = sum(int(data[k][ist]) for k in range(s, e + 1))
    if c * n > tot * 0.005:
        print(f"  sass[{s:4d}..{e:4d}] n={n:4d} exec={c:9d} inst={c * n / tot * 100:5.1f}% stalls={samp / max(1, tots) * 100:5.1f}%  {data[s][isrc].strip()[:50]}", file=out)
