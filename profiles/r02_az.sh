#!/bin/bash
# round 2, call AZ: parity soak of the global-implicit entry points; second RReact soak with another seed
mkdir -p gpurun_out
timeout 1800 python profiles/parity_soak_gi.py 77002 1.0 > gpurun_out/r02_az_parity_soak_gi.jsonl 2> gpurun_out/r02_az_parity_soak_gi.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_az_parity_soak_gi.jsonl'):
    d=json.loads(l); print('%-22s %8d cells  state %.2e  accum %.2e  residual %.2e  jacobian %.2e' % (d['workload'], d['cells'], d['max_err_state_after_RTUpdateAuxVars'], d['max_err_fixed_accumulation'], d['max_err_residual_blocks'], d['max_err_jacobian_blocks']))
PY
tail -3 gpurun_out/r02_az_parity_soak_gi.err
timeout 1800 python profiles/parity_soak.py 88003 1.0 > gpurun_out/r02_az_parity_soak_seed2.jsonl 2> gpurun_out/r02_az_parity_soak_seed2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_az_parity_soak_seed2.jsonl'):
    d=json.loads(l); print('%-26s %8d cells  mismatches %d  max rel err %s  >1e-10: %d' % (d['workload'], d['cells'], d['iteration_or_flag_mismatches'], d['max_rel_err_free_ion'], d['cells_above_1e-10']))
PY
