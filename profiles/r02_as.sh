#!/bin/bash
# round 2, call AS: work order on small batches (all chemistries, resident-lane and tensor-memory kernel): neutrality tests; throughput with / without at rank-sized batches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "work_order" 2>&1 | tail -4 > gpurun_out/r02_as_pytest.log; cat gpurun_out/r02_as_pytest.log
for wl in hanford300a_eq calcite hanford300a_mr; do
for n in 50000 100000 300000 1000000; do
  for mode in off on; do
    if [ $mode = off ]; then export RXN_NO_REACT_ORDER=1; else unset RXN_NO_REACT_ORDER; fi
    timeout 600 python bench.py --workload $wl --cells $n --steps 10 --warmup 3 --no-extra > gpurun_out/r02_as_${wl}_${n}_$mode.json 2> gpurun_out/r02_as_${wl}_${n}_$mode.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_as_${wl}_${n}_$mode.json').read().strip().splitlines()[-1])
    print('$wl $n order $mode: %.2f M/s e2e %.2f frac %.4f kernel_ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms']))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02_as_${wl}_${n}_$mode.err').read()[-800:])
PY
  done
done
done
