#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU, under gpurun); reports land in gpurun_out/, summaries are made
# here afterwards with tools/ncu_summary.py.
out=gpurun_out
N="--set full --import-source on --clock-control none"
# 1. the bench kernel through ndl_match_batch at the bench's size class (2^25 lines = 2 GiB; same code path as the 8 GiB batch)
timeout 600 ncu $N -k regex:linesq_kernel -s 2 -c 1 -o $out/r02m_c4b_batch python exp/one_launch.py c4 33554432 3 2 > $out/r02m_1.log 2>&1
# 2. iterated find on the staged tiles, SSN regex over 64-byte lines (count pass, then the fill pass)
timeout 600 ncu $N -k regex:linesq_kernel -s 0 -c 2 -o $out/r02m_find_all python exp/find_all_bench.py 4000000 > $out/r02m_2.log 2>&1
# 3. one UTF-16 haystack, chunk-parallel
timeout 600 ncu $N -k regex:long8_kernel -s 1 -c 1 -o $out/r02m_long8_utf16 python exp/find_long_utf16_bench.py 0.5 > $out/r02m_3.log 2>&1
# 4. 96-byte records (run-time chunk count)
timeout 600 ncu $N -k regex:linesq_kernel -s 70 -c 1 -o $out/r02m_reclen python exp/reclen_bench.py > $out/r02m_4.log 2>&1
ls -la $out/r02m_*.ncu-rep
tail -2 $out/r02m_1.log $out/r02m_2.log $out/r02m_3.log $out/r02m_4.log
