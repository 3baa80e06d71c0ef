#!/bin/bash
# round 2, call R: hpt logK with exact corrected divisions (GPU parity + timing), sanitizer over the flux-by-columns and coupler kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flux.py -m gpu -q -k "hpt or fit5 or flux or boundary" 2>&1 | tail -4
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2: %.1f M/s e2e %.1f kernel_ms %.3f %s frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac']))
except Exception as e: print('$2 failed', e)
PY
}
timeout 300 python bench.py --mode gi --steps 5 --warmup 3 > gpurun_out/r02_r_gi_hpt.json 2> gpurun_out/r02_r_gi_hpt.err; show gpurun_out/r02_r_gi_hpt.json "gi hpt"
timeout 300 python bench.py --workload hpt_calcite --steps 5 --warmup 3 > gpurun_out/r02_r_react_hpt.json 2> gpurun_out/r02_r_react_hpt.err; show gpurun_out/r02_r_react_hpt.json "react hpt"
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 10 python profiles/sanitize_run.py flux > gpurun_out/r02_r_sanitizer_flux_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -3 gpurun_out/r02_r_sanitizer_flux_$tool.log
done
