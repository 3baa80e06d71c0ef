#!/bin/bash
# round 2, call AF: the aqueous chemistry of the MPHASE CO2 deck (scco2_brine, 8 primaries / 12 complexes / 2 kinetic minerals): GPU parity, RTReact and global-implicit bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "scco2" 2>&1 | tail -4 > gpurun_out/r02_af_pytest_scco2.log; cat gpurun_out/r02_af_pytest_scco2.log
timeout 600 python bench.py --workload scco2_brine --steps 5 --warmup 3 --no-extra > gpurun_out/r02_af_react_scco2.json 2> gpurun_out/r02_af_react_scco2.err
timeout 600 python bench.py --mode gi --workload scco2_brine --steps 5 --warmup 3 > gpurun_out/r02_af_gi_scco2.json 2> gpurun_out/r02_af_gi_scco2.err
python - <<'PY'
import json
for f in ('react','gi'):
    try:
        d=json.loads(open('gpurun_out/r02_af_%s_scco2.json' % f).read().strip().splitlines()[-1])
        print(f, '%.1f M/s e2e %.1f kernel_ms %.3f %s frac %.3f cpu %.2f M its %s | %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac'], d['cpu_baseline']['value']/1e6, d['config'].get('mean_newton_iterations'), d['config'].get('kernel')))
    except Exception as e:
        print(f, 'failed', e); print(open('gpurun_out/r02_af_%s_scco2.err' % f).read()[-1200:])
PY
