#!/bin/bash
# round 2, call S: full GPU suite + smoke at the current source state; headline bench (default line incl. the other configs)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_s_pytest_gpu.log
cat gpurun_out/r02_s_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_s_smoke.log 2>&1; tail -4 gpurun_out/r02_s_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_s_bench_default.json 2> gpurun_out/r02_s_bench_default.err; tail -2 gpurun_out/r02_s_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms']))
for o in d.get('other_configs', []):
    print(o.get('config'), o.get('name'), '%.1f M/s' % (o.get('value', 0)/1e6), 'e2e %.1f' % (o.get('e2e', {}).get('value', 0)/1e6), o.get('roofline', {}).get('bound'), '%.3f' % o.get('roofline', {}).get('frac', 0), o.get('error', ''))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
