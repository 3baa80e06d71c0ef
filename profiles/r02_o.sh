#!/bin/bash
# round 2, call O: C driver on the GPU; ncu of the small-chemistry RReact kernel (config 2: calcite) and of the hpt one
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c_driver.py -m gpu -q 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_lane -s 2 -c 1 -o gpurun_out/r02_o_lane_calcite \
  python bench.py --workload calcite --steps 1 --warmup 1 --cells 4000000 > gpurun_out/r02_o_ncu_calcite.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_lane -s 2 -c 1 -o gpurun_out/r02_o_lane_hpt \
  python bench.py --workload hpt_calcite --steps 1 --warmup 1 --cells 4000000 > gpurun_out/r02_o_ncu_hpt.log 2>&1
ls -la gpurun_out/r02_o*.ncu-rep
