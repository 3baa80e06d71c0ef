#!/bin/bash
# round 2, call AG: N = 8 resident-lane shape, 192 cells per CTA: 1 / 2 / 4 lanes per cell on the scco2_brine chemistry
mkdir -p gpurun_out
for g in 1 2 4; do
  RXN_LANE_G=$g timeout 600 python bench.py --workload scco2_brine --steps 5 --warmup 3 --no-extra > gpurun_out/r02_ag_scco2_g$g.json 2> gpurun_out/r02_ag_scco2_g$g.err
  RXN_LANE_G=$g timeout 600 python bench.py --mode gi --workload scco2_brine --steps 5 --warmup 3 > gpurun_out/r02_ag_gi_scco2_g$g.json 2> gpurun_out/r02_ag_gi_scco2_g$g.err
  python - <<PY
import json
for f in ('r02_ag_scco2_g$g','r02_ag_gi_scco2_g$g'):
    try:
        d=json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.1f M/s e2e %.1f kernel_ms %.3f %s frac %.3f | %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac'], d['config'].get('kernel')))
    except Exception as e:
        print(f, 'failed', e); print(open('gpurun_out/%s.err' % f).read()[-800:])
PY
  RXN_LANE_G=$g timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scco2" 2>&1 | tail -2
done
