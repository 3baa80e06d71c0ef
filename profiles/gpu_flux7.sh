#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_flux.py -x -q 2>&1 | tail -3
timeout 300 python profiles/bench_flux.py hanford300a_eq 128 128 64 > gpurun_out/bench_flux_300a.json 2> gpurun_out/bench_flux_300a.err
timeout 300 python profiles/bench_flux.py hanford300a_eq 100 100 100 > gpurun_out/bench_flux_300a_100.json 2> gpurun_out/bench_flux_300a_100.err
timeout 300 python profiles/bench_flux.py calcite 100 100 100 > gpurun_out/bench_flux_calcite_100.json 2> gpurun_out/bench_flux_calcite_100.err
timeout 300 python profiles/bench_flux.py calcite 256 256 128 > gpurun_out/bench_flux_calcite.json 2> gpurun_out/bench_flux_calcite.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux_residual -s 1 -c 1 -o gpurun_out/flux_res \
  python profiles/bench_flux.py hanford300a_eq 100 100 50 > gpurun_out/ncu_flux_res.log 2>&1
for f in gpurun_out/bench_flux_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['workload'], d['grid'], {k:(round(d[k]['kernel_ms'],3), round(d[k]['hbm_frac'],3)) for k in ('flux_coefs','flux_residual','flux_jacobian')})"; done
