#!/bin/bash
# round 2, call BB: ncu --set full of the headline kernel at the final source state (600 000 cells: above 16 generations of resident cells, index order)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_tm -s 1 -c 1 -f -o gpurun_out/r02_bb_tm_g3 \
  python bench.py --steps 1 --warmup 1 --cells 600000 --no-extra > gpurun_out/r02_bb_ncu_tm.log 2>&1
tail -2 gpurun_out/r02_bb_ncu_tm.log; ls -la gpurun_out/r02_bb_tm_g3.ncu-rep
