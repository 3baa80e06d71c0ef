#!/bin/bash
# round 2, call AC2: ascem with the G = 8 default: GPU parity, bench at 10^6 cells, ncu capture of k_react_lane<24,16,8>
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "ascem" 2>&1 | tail -5 > gpurun_out/r02_ac2_pytest_ascem.log; cat gpurun_out/r02_ac2_pytest_ascem.log
timeout 900 python bench.py --workload ascem --steps 3 --warmup 3 --no-extra > gpurun_out/r02_ac2_bench_ascem.json 2> gpurun_out/r02_ac2_bench_ascem.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ac2_bench_ascem.json').read().strip().splitlines()[-1])
print('ascem 1e6: %.3f M/s e2e %.3f frac %.4f kernel_ms %.1f  %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['kernel_ms'], d['config']['kernel']))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_lane -c 1 -o gpurun_out/r02_ac2_ascem_lane_g8 -f python bench.py --workload ascem --cells 100000 --steps 1 --warmup 0 --no-extra > gpurun_out/r02_ac2_ncu.log 2>&1; tail -3 gpurun_out/r02_ac2_ncu.log
