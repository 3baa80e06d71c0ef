#!/bin/bash
for g in "64 64 32" "64 64 64" "128 64 64" "128 128 64"; do
  timeout 300 python profiles/bench_flux.py hanford300a_eq $g 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['grid'], d['cells'], {k: (round(d[k]['kernel_ms'], 3), round(d[k]['hbm_frac'], 3)) for k in ('flux_coefs', 'flux_residual', 'flux_jacobian')})
    else: print(l.rstrip())
"
done
