#!/bin/bash
# round 2, call AM: parity soak on fresh random cells (profiles/parity_soak.py)
mkdir -p gpurun_out
timeout 2400 python profiles/parity_soak.py 77001 1.0 > gpurun_out/r02_am_parity_soak.jsonl 2> gpurun_out/r02_am_parity_soak.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_am_parity_soak.jsonl'):
    d=json.loads(l); print('%-26s %8d cells  mismatches %d  nonref %d  max rel err %s  >1e-10: %d  its mean %.2f max %d  oracle %.1fs' % (d['workload'], d['cells'], d['iteration_or_flag_mismatches'], d['cells_with_nonreference_flags'], d['max_rel_err_free_ion'], d['cells_above_1e-10'], d['mean_newton_iterations'], d['max_newton_iterations'], d['oracle_s']))
PY
tail -3 gpurun_out/r02_am_parity_soak.err
