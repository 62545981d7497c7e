#!/bin/bash
# A/B: warps per CTA x table layout on the default bench workloads (steps kept short)
for lib in "" _q24 _q28 _q32; do
  for force in "" "4,4,2" "4,1,2" "4,16,4"; do
    for w in c2 c4b; do
      out=$(NEEDLE_B200_LIB=$PWD/needle_b200/libneedle_b200$lib.so NDL_Q_FORCE=$force timeout 120 python bench.py --workload $w --steps 100 --warmup 5 2>/dev/null | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), d['roofline']['kernel'], 'stride', round(d.get('stride_api',{}).get('value',0),1))
except Exception as e: print('ERR', e)")
      echo "lib=${lib:-q20} force=${force:-auto} $w: $out"
    done
  done
done
