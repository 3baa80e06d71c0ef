#!/bin/bash
# round 2, call U: register kernel for small chemistries: GPU suite, calcite / hpt benches against the resident-lane kernel, ncu
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_u_pytest_gpu.log
cat gpurun_out/r02_u_pytest_gpu.log
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2: %.1f M/s e2e %.1f kernel_ms %.3f %s frac %.3f hbm %.3f bad %d | %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['config']['cells_with_nonreference_flags'], d['config']['kernel'][:40]))
except Exception as e: print('$2 failed', e)
PY
}
for wl in calcite hpt_calcite; do
  for sm in 0 1; do
    RXN_SMALL=$sm timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-extra --cells 4000000 > gpurun_out/r02_u_${wl}_small$sm.json 2> gpurun_out/r02_u_${wl}_small$sm.err; show gpurun_out/r02_u_${wl}_small$sm.json "$wl 4e6 RXN_SMALL=$sm"
  done
done
timeout 300 python bench.py --workload calcite --steps 10 --warmup 3 --no-extra > gpurun_out/r02_u_calcite_1e6.json 2>/dev/null; show gpurun_out/r02_u_calcite_1e6.json "calcite 1e6 (config 2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_small -s 2 -c 1 -o gpurun_out/r02_u_small_calcite \
  python bench.py --workload calcite --steps 1 --warmup 1 --cells 4000000 --no-extra > gpurun_out/r02_u_ncu.log 2>&1
ls -la gpurun_out/r02_u*.ncu-rep
