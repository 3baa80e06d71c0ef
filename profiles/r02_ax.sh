#!/bin/bash
# round 2, call AX: final policy of the work order (tail-bound chemistries; >= 8 primaries below 16 generations of resident cells): tests, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "work_order" 2>&1 | tail -4 > gpurun_out/r02_ax_pytest.log; cat gpurun_out/r02_ax_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_ax_bench_default.json 2> gpurun_out/r02_ax_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_ax_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f launches %d work_order=%s unordered=%s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['config']['work_order'][:40], d['roofline']['unordered']))
for o in d.get('other_configs', []):
    print(o.get('config'), o.get('name'), '%.2f M/s' % (o.get('value', 0)/1e6), 'e2e %.2f' % (o.get('e2e', {}).get('value', 0)/1e6), o.get('roofline', {}).get('bound'), '%.3f' % o.get('roofline', {}).get('frac', 0), o.get('error', ''))
PY
timeout 300 python bench.py --cells 100000 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_ax_bench_300a_100k.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_ax_bench_300a_100k.json').read().strip().splitlines()[-1])
u=d['roofline']['unordered']
print('300A 1e5: %.2f M/s frac %.4f | unordered %.2f M/s frac %.4f | %s' % (d['value']/1e6, d['roofline']['frac'], u['value']/1e6, u['frac'], d['config']['work_order'][:30]))
PY
