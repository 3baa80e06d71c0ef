#!/bin/bash
# round 2, call P: flux Jacobian by block columns: GPU flux tests (bitwise), timing against the row walk, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flux.py -m gpu -q 2>&1 | tail -5
for rows in 1 0; do
  RXN_FLUX_ROWS=$rows timeout 300 python profiles/bench_flux.py hanford300a_eq 100 100 100 > gpurun_out/r02_p_flux_rows$rows.json 2> gpurun_out/r02_p_flux_rows$rows.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_p_flux_rows$rows.json').read().strip().splitlines()[-1])
print('RXN_FLUX_ROWS=$rows', {k: (round(v['kernel_ms'],3), round(v['hbm_frac'],3)) for k,v in d.items() if isinstance(v, dict)}, 'checksum', d['checksum'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_flux_jacobian_cols -s 2 -c 1 -o gpurun_out/r02_p_flux_cols \
  python profiles/bench_flux.py hanford300a_eq 100 100 50 > gpurun_out/r02_p_ncu.log 2>&1
ls -la gpurun_out/r02_p*.ncu-rep
