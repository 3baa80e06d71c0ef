#!/bin/bash
# round 2, call AB: ascem (BASELINE config 1 chemistry, N = 24 shape, 16 cells per CTA): 2 / 4 / 8 lanes per cell
mkdir -p gpurun_out
for g in 2 4 8; do
  RXN_LANE_G=$g RXN_LANE_CPB=16 timeout 600 python bench.py --workload ascem --cells 300000 --steps 3 --warmup 3 --no-extra > gpurun_out/r02_ab_ascem_g$g.json 2> gpurun_out/r02_ab_ascem_g$g.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_ab_ascem_g$g.json').read().strip().splitlines()[-1])
    print('G=$g: %.3f M/s e2e %.3f frac %.4f kernel_ms %.1f  %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['config']['kernel']))
except Exception as e:
    print('G=$g failed', e); print(open('gpurun_out/r02_ab_ascem_g$g.err').read()[-1500:])
PY
done
