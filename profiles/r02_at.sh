#!/bin/bash
# round 2, call AT: work order at the headline size (experiment: RXN_REACT_ORDER_ALL=1 lifts the 64-generation limit for single launches)
mkdir -p gpurun_out
for mode in off on; do
  if [ $mode = on ]; then export RXN_REACT_ORDER_ALL=1; else unset RXN_REACT_ORDER_ALL; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r02_at_bench_300a_order_$mode.json 2> gpurun_out/r02_at_bench_300a_order_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_at_bench_300a_order_$mode.json').read().strip().splitlines()[-1])
print('300A 1e7 order-all $mode: %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms']))
PY
done
