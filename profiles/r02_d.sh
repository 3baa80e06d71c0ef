#!/bin/bash
# round 2, call D: cp.async load; G sweep on 2e6 cells, full 1e7 config for the best, multirate + other chemistries
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "react" 2>&1 | tail -3 > gpurun_out/r02_d_pytest_react.log
cat gpurun_out/r02_d_pytest_react.log
show() { python -c "
import json,sys
try:
  d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', '%.2f M/s e2e %.2f kernel_ms %.2f frac %.4f bad %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['cells_with_nonreference_flags']), d['config']['kernel'][:60])
except Exception as e: print('$2 failed', e)
"; }
for g in 4; do
  RXN_TM_G=$g timeout 200 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/r02_d_300a_tm_g$g.json 2> gpurun_out/r02_d_300a_tm_g$g.err; show gpurun_out/r02_d_300a_tm_g$g.json "300A 2e6 TM G=$g"
  RXN_TM_G=$g timeout 200 python bench.py --steps 3 --warmup 3 --workload hanford300a_mr --cells 1000000 > gpurun_out/r02_d_mr_tm_g$g.json 2> gpurun_out/r02_d_mr_tm_g$g.err; show gpurun_out/r02_d_mr_tm_g$g.json "mr 1e6 TM G=$g"
done
RXN_TM=0 timeout 200 python bench.py --steps 3 --warmup 3 --workload hanford300a_mr --cells 1000000 > gpurun_out/r02_d_mr_lane.json 2> gpurun_out/r02_d_mr_lane.err; show gpurun_out/r02_d_mr_lane.json "mr 1e6 lane"
RXN_TM_G=4 timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_d_300a_full_g4.json 2> gpurun_out/r02_d_300a_full_g4.err; show gpurun_out/r02_d_300a_full_g4.json "300A 1e7 TM G=4"

