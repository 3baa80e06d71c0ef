#!/bin/bash
# round 2, call K: ncu of the global-implicit kernels (config 4: hpt chemistry; 300A)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_auxvars|k_gi_lane" -s 2 -c 2 -o gpurun_out/r02_k_gi_hpt \
  python bench.py --mode gi --steps 1 --warmup 1 --cells 2000000 > gpurun_out/r02_k_ncu_gi_hpt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_auxvars|k_gi_lane" -s 2 -c 2 -o gpurun_out/r02_k_gi_300a \
  python bench.py --mode gi --workload hanford300a_eq --steps 1 --warmup 1 --cells 400000 > gpurun_out/r02_k_ncu_gi_300a.log 2>&1
ls -la gpurun_out/*.ncu-rep
