#!/bin/bash
# round 2, call AE: evidence after the microbial / immobile and ascem-shape changes: GPU suite, smoke, default bench line (now with the
# config-1 chemistry in other_configs), ascem at 10^6 cells (unchunked host-buffer path), launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_ae_pytest_gpu.log
cat gpurun_out/r02_ae_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_ae_smoke.log 2>&1; tail -3 gpurun_out/r02_ae_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_ae_bench_default.json 2> gpurun_out/r02_ae_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_ae_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']))
for o in d.get('other_configs', []):
    print(o.get('config'), o.get('name'), '%.2f M/s' % (o.get('value', 0)/1e6), 'e2e %.2f' % (o.get('e2e', {}).get('value', 0)/1e6), o.get('roofline', {}).get('bound'), '%.3f' % o.get('roofline', {}).get('frac', 0), 'cpu %.3f M' % (o.get('cpu_baseline', {}).get('value', 0)/1e6), o.get('error', ''))
PY
timeout 600 python bench.py --workload ascem --steps 3 --warmup 3 --no-extra > gpurun_out/r02_ae_bench_ascem.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_ae_bench_ascem.json').read().strip().splitlines()[-1])
print('ascem 1e6: %.3f M/s e2e %.3f frac %.4f kernel_ms %.1f cpu %.3f M  %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['cpu_baseline']['value']/1e6, d['config']['kernel']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ae_launches.csv python bench.py --steps 2 --warmup 1 --cells 1000000 --no-extra > /dev/null 2>&1
