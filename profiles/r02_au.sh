#!/bin/bash
# round 2, call AU: work order for every batch size, single launches and the chunks of the host-buffer call: neutrality tests, default bench with the unordered steps beside it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "work_order" 2>&1 | tail -4 > gpurun_out/r02_au_pytest.log; cat gpurun_out/r02_au_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_au_bench_default.json 2> gpurun_out/r02_au_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_au_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f launches %d unordered %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], {k: (round(v,4) if isinstance(v,float) else v) for k,v in (d['roofline']['unordered'] or {}).items() if k!='note'}))
for o in d.get('other_configs', []):
    print(o.get('config'), o.get('name'), '%.2f M/s' % (o.get('value', 0)/1e6), 'e2e %.2f' % (o.get('e2e', {}).get('value', 0)/1e6), o.get('roofline', {}).get('bound'), '%.3f' % o.get('roofline', {}).get('frac', 0), o.get('error', ''))
PY
tail -3 gpurun_out/r02_au_bench_default.err
