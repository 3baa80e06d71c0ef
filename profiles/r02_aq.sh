#!/bin/bash
# round 2, call AQ: host-buffer RTReact with the chunk kernels on two alternating streams (16 chunks): full-size parity tests, default bench, calcite / multirate e2e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "full_size or baseline_size or 1000000 or 200000 or device_resident or l2g" 2>&1 | tail -4 > gpurun_out/r02_aq_pytest.log; cat gpurun_out/r02_aq_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_aq_bench_default.json 2> gpurun_out/r02_aq_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_aq_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']))
for o in d.get('other_configs', []):
    print(o.get('config'), o.get('name'), '%.2f M/s' % (o.get('value', 0)/1e6), 'e2e %.2f' % (o.get('e2e', {}).get('value', 0)/1e6), o.get('error', ''))
PY
timeout 300 python bench.py --workload hanford300a_mr --steps 5 --warmup 3 --no-extra > gpurun_out/r02_aq_react_mr.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_aq_react_mr.json').read().strip().splitlines()[-1]); print('react mr: %.2f M/s e2e %.2f frac %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac']))"
