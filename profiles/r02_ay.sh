#!/bin/bash
# round 2, call AY: final evidence at the final source state (work-order policy, two-stream chunks, streaming kinetic state, microbial / immobile): GPU suite, smoke, default bench line, gi lines, kinetic-state kernel, launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_ay_pytest_gpu.log
cat gpurun_out/r02_ay_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_ay_smoke.log 2>&1; tail -3 gpurun_out/r02_ay_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_ay_bench_default.json 2> gpurun_out/r02_ay_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_ay_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f kernel_ms %.2f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']))
for o in d.get('other_configs', []):
    print(o.get('config'), o.get('name'), '%.2f M/s' % (o.get('value', 0)/1e6), 'e2e %.2f' % (o.get('e2e', {}).get('value', 0)/1e6), o.get('roofline', {}).get('bound'), '%.3f' % o.get('roofline', {}).get('frac', 0), 'cpu %.3f M' % (o.get('cpu_baseline', {}).get('value', 0)/1e6), o.get('error', ''))
PY
for wl in hpt_calcite hanford300a_eq hanford300a_mr scco2_brine; do
  timeout 300 python bench.py --mode gi --workload $wl --steps 5 --warmup 3 > gpurun_out/r02_ay_gi_$wl.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ay_gi_$wl.json').read().strip().splitlines()[-1]); print('gi $wl: %.1f M/s e2e %.1f kernel_ms %.3f %s frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac']))
PY
done
timeout 300 python bench.py --workload hanford300a_mr --steps 5 --warmup 3 --no-extra > gpurun_out/r02_ay_react_mr.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_ay_react_mr.json').read().strip().splitlines()[-1]); print('react mr: %.2f M/s e2e %.2f frac %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac']))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_ay_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/r02_ay_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ay_launches.csv python bench.py --steps 2 --warmup 1 --cells 1000000 --no-extra > /dev/null 2>&1
