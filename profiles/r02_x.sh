#!/bin/bash
# round 2, call X: per-cell logK cache: GPU suite, hpt benches with and without the cache
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_x_pytest_gpu.log
cat gpurun_out/r02_x_pytest_gpu.log
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2: %.1f M/s e2e %.1f kernel_ms %.3f %s frac %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac']))
except Exception as e: print('$2 failed', e)
PY
}
for c in 0 1; do
  RXN_LOGK_CACHE=$c timeout 300 python bench.py --mode gi --steps 5 --warmup 3 > gpurun_out/r02_x_gi_hpt_c$c.json 2>/dev/null; show gpurun_out/r02_x_gi_hpt_c$c.json "gi hpt cache=$c"
  RXN_LOGK_CACHE=$c timeout 300 python bench.py --workload hpt_calcite --steps 10 --warmup 3 --no-extra > gpurun_out/r02_x_react_hpt_c$c.json 2>/dev/null; show gpurun_out/r02_x_react_hpt_c$c.json "react hpt cache=$c"
done
