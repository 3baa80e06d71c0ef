#!/bin/bash
for lib in librxn_b200.so librxn_b200_alt.so librxn_b200_alt2.so; do
  export RXN_B200_LIB=$PWD/pflotran_b200/$lib
  timeout 300 python profiles/bench_flux.py hanford300a_eq 100 100 100 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', d['grid'], d['cells'], {k: (round(d[k]['kernel_ms'], 3), round(d[k]['hbm_frac'], 3)) for k in ('flux_coefs', 'flux_residual', 'flux_jacobian')})
    else: print(l.rstrip())
"
done
