#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flux.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_flux.log
for g in "128 128 64" "100 100 100" "125 125 67"; do
  timeout 300 python profiles/bench_flux.py hanford300a_eq $g 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['grid'], {k: (round(d[k]['kernel_ms'], 3), round(d[k]['hbm_frac'], 3)) for k in ('flux_coefs', 'flux_residual', 'flux_jacobian')})
    else: print(l.rstrip())
"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux_jacobian -s 1 -c 1 -o gpurun_out/flux_jac \
  python profiles/bench_flux.py hanford300a_eq 100 100 50 > gpurun_out/ncu_flux.log 2>&1
cat gpurun_out/pytest_flux.log
