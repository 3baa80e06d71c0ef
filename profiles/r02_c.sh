#!/bin/bash
# round 2, call C: GPU tests of the TMEM kernel + full ncu capture with source for G=2 and G=4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "react" 2>&1 | tail -5 > gpurun_out/r02_c_pytest_react.log
cat gpurun_out/r02_c_pytest_react.log
for g in 2 4; do
RXN_TM_G=$g timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_react_tm -s 1 -c 1 -o gpurun_out/r02_c_tm_g$g \
  python bench.py --steps 1 --warmup 1 --cells 400000 > gpurun_out/r02_c_ncu_tm_g$g.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
