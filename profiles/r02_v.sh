#!/bin/bash
# round 2, call V: rxn_exp (lean exp for the on-chip kernels): GPU suite, react benches of every workload, gi
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_v_pytest_gpu.log
cat gpurun_out/r02_v_pytest_gpu.log
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2: %.1f M/s e2e %.1f kernel_ms %.3f %s frac %.4f bad %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac'], d['config'].get('cells_with_nonreference_flags')))
except Exception as e: print('$2 failed', e)
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r02_v_300a.json 2>/dev/null; show gpurun_out/r02_v_300a.json "300A 1e7"
timeout 300 python bench.py --workload hanford300a_mr --steps 3 --warmup 3 --cells 1000000 --no-extra > gpurun_out/r02_v_mr.json 2>/dev/null; show gpurun_out/r02_v_mr.json "mr 1e6"
timeout 300 python bench.py --workload calcite --steps 10 --warmup 3 --no-extra > gpurun_out/r02_v_calcite.json 2>/dev/null; show gpurun_out/r02_v_calcite.json "calcite 1e6"
timeout 300 python bench.py --workload calcite --steps 10 --warmup 3 --no-extra --cells 4000000 > gpurun_out/r02_v_calcite4.json 2>/dev/null; show gpurun_out/r02_v_calcite4.json "calcite 4e6"
timeout 300 python bench.py --workload hpt_calcite --steps 10 --warmup 3 --no-extra > gpurun_out/r02_v_hpt.json 2>/dev/null; show gpurun_out/r02_v_hpt.json "hpt 1e6"
timeout 300 python bench.py --mode gi --steps 5 --warmup 3 > gpurun_out/r02_v_gi_hpt.json 2>/dev/null; show gpurun_out/r02_v_gi_hpt.json "gi hpt"
timeout 300 python bench.py --mode gi --workload hanford300a_eq --steps 5 --warmup 3 > gpurun_out/r02_v_gi_300a.json 2>/dev/null; show gpurun_out/r02_v_gi_300a.json "gi 300A"
