#!/bin/bash
# round 2, call M: after the column-wise Jacobian output and the multirate rate table: GPU suite, gi + react benches
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_m_pytest_gpu.log
cat gpurun_out/r02_m_pytest_gpu.log
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2: %.1f M/s e2e %.1f kernel_ms %.3f %s frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac']))
except Exception as e: print('$2 failed', e)
PY
}
for wl in hanford300a_eq hanford300a_mr; do
  timeout 300 python bench.py --mode gi --workload $wl --steps 5 --warmup 3 > gpurun_out/r02_m_gi_$wl.json 2> gpurun_out/r02_m_gi_$wl.err; show gpurun_out/r02_m_gi_$wl.json "gi $wl"
done
timeout 300 python bench.py --workload hanford300a_mr --steps 3 --warmup 3 --cells 1000000 > gpurun_out/r02_m_react_mr.json 2> gpurun_out/r02_m_react_mr.err; show gpurun_out/r02_m_react_mr.json "react mr 1e6"
timeout 300 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/r02_m_react_300a.json 2> gpurun_out/r02_m_react_300a.err; show gpurun_out/r02_m_react_300a.json "react 300A 2e6"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_m_gi_launches.csv python bench.py --mode gi --workload hanford300a_mr --steps 2 --warmup 1 --cells 200000 > /dev/null 2>&1
grep "k_gi_tm" gpurun_out/r02_m_gi_launches.csv | tail -4 | cut -c1-60,300-
