#!/bin/bash
# round 2, call AO: work order of the resident-lane kernel for tail-bound chemistries (slowest cells of the previous call first): neutrality test, ascem with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "work_order or ascem" 2>&1 | tail -4 > gpurun_out/r02_ao_pytest.log; cat gpurun_out/r02_ao_pytest.log
for n in 100000 1000000; do
  for mode in off on; do
    if [ $mode = off ]; then export RXN_NO_REACT_ORDER=1; else unset RXN_NO_REACT_ORDER; fi
    timeout 600 python bench.py --workload ascem --cells $n --steps 4 --warmup 3 --no-extra > gpurun_out/r02_ao_ascem_${n}_$mode.json 2> gpurun_out/r02_ao_ascem_${n}_$mode.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_ao_ascem_${n}_$mode.json').read().strip().splitlines()[-1])
    print('ascem $n order $mode: %.3f M/s e2e %.3f frac %.4f kernel_ms %.1f | %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['config']['kernel'][-70:]))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02_ao_ascem_${n}_$mode.err').read()[-1500:])
PY
  done
done
