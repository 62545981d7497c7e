#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): GPU tests, both bench arms, launch list, DRAM traffic of the bench kernel.
# Outputs under gpurun_out/$1_*; copy what should be judged into profiles/.
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 1200 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench rc=$?"
cat $out/${tag}_bench_default.json | cut -c1-600
# launch list of the same command (serialised, cold-cache: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras > $out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
# DRAM bytes per launch of the bench kernel (one pass per metric group, all launches of the short run)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:linesq -c 60 --csv \
  --log-file $out/${tag}_traffic.csv python bench.py --steps 1 --warmup 1 --no-extras > $out/${tag}_traffic.log 2>&1; echo "traffic rc=$?"
