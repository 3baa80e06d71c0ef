#!/bin/bash
mkdir -p gpurun_out
export RXN_LANE_VERBOSE=1
RXN_LANE_G=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "test_react or inactive or full_size" 2>&1 | tail -4 > gpurun_out/pytest_react_g2.log
for cfg in "RXN_LANE_G=2" "RXN_LANE_G=1" "RXN_LANE_G=4 RXN_LANE_CPB=60"; do
  tag=$(echo $cfg | tr -d ' =' | sed 's/RXN_LANE_//g')
  env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --cells 2000000 > gpurun_out/bench_300a_$tag.json 2> gpurun_out/bench_300a_$tag.err
  env $cfg RXN_B200_LIB=$PWD/pflotran_b200/librxn_b200_alt.so timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --cells 2000000 > gpurun_out/bench_300a_alt_$tag.json 2> gpurun_out/bench_300a_alt_$tag.err
done
for g in 2 4; do
RXN_LANE_G=$g timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload hanford300a_mr --cells 1000000 > gpurun_out/bench_mr_G$g.json 2> gpurun_out/bench_mr_G$g.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload calcite --cells 4000000 > gpurun_out/bench_calcite_k3.json 2> gpurun_out/bench_calcite_k3.err
cat gpurun_out/pytest_react_g*.log
