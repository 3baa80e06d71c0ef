#!/bin/bash
# round 2, call AD: ascem (N = 24 shape): cells per CTA x lanes per cell: 16x8, 20x8 (bank-conflict-free row stride), 20x16; parity of each
mkdir -p gpurun_out
for s in "16 8" "20 8" "20 16"; do
  set -- $s
  RXN_LANE_CPB=$1 RXN_LANE_G=$2 timeout 600 python bench.py --workload ascem --cells 500000 --steps 3 --warmup 3 --no-extra > gpurun_out/r02_ad_ascem_$1_$2.json 2> gpurun_out/r02_ad_ascem_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_ad_ascem_$1_$2.json').read().strip().splitlines()[-1])
    print('CPB=$1 G=$2: %.3f M/s e2e %.3f frac %.4f kernel_ms %.1f  %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['config']['kernel']))
except Exception as e:
    print('CPB=$1 G=$2 failed', e); print(open('gpurun_out/r02_ad_ascem_$1_$2.err').read()[-1500:])
PY
  RXN_LANE_CPB=$1 RXN_LANE_G=$2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ascem and test_react" 2>&1 | tail -2
done
