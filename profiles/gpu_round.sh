#!/bin/bash
mkdir -p gpurun_out
export RXN_LANE_VERBOSE=1
RXN_LANE_G=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "test_react or inactive or full_size" 2>&1 | tail -4 > gpurun_out/pytest_react_g2.log
for g in 2 4 1; do
  RXN_LANE_G=$g timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --cells 2000000 > gpurun_out/bench_300a_g$g.json 2> gpurun_out/bench_300a_g$g.err
done
for g in 2 1; do
RXN_B200_LIB=$PWD/pflotran_b200/librxn_b200_alt.so RXN_LANE_G=$g timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --cells 2000000 > gpurun_out/bench_300a_alt_g$g.json 2> gpurun_out/bench_300a_alt_g$g.err
done
for g in 2 4; do
RXN_LANE_G=$g timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload hanford300a_mr --cells 1000000 > gpurun_out/bench_mr_g$g.json 2> gpurun_out/bench_mr_g$g.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload calcite --cells 4000000 > gpurun_out/bench_calcite_k3.json 2> gpurun_out/bench_calcite_k3.err
RXN_B200_LIB=$PWD/pflotran_b200/librxn_b200_alt.so timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload calcite --cells 4000000 > gpurun_out/bench_calcite_alt.json 2> gpurun_out/bench_calcite_alt.err
RXN_LANE_G=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_react_lane -s 2 -c 1 -o gpurun_out/lane_15_64_2 \
  python bench.py --steps 1 --warmup 1 --kernel 3 --cells 600000 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_react_g*.log
