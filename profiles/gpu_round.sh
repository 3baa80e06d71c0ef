#!/bin/bash
# one gpurun call: parity tests of the resident-lane kernel, bench comparison, ncu captures
mkdir -p gpurun_out
export RXN_LANE_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_react or inactive" 2>&1 | tail -15 > gpurun_out/pytest_react.log
for k in 3 2; do
  timeout 300 python bench.py --steps 3 --warmup 3 --kernel $k --cells 2000000 > gpurun_out/bench_300a_k$k.json 2> gpurun_out/bench_300a_k$k.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload hanford300a_mr --cells 1000000 > gpurun_out/bench_mr_k3.json 2> gpurun_out/bench_mr_k3.err
timeout 300 python bench.py --steps 3 --warmup 3 --kernel 3 --workload calcite --cells 4000000 > gpurun_out/bench_calcite_k3.json 2> gpurun_out/bench_calcite_k3.err
timeout 300 python bench.py --steps 3 --warmup 3 --kernel 2 --workload calcite --cells 4000000 > gpurun_out/bench_calcite_k2.json 2> gpurun_out/bench_calcite_k2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_react_lane -s 2 -c 1 -o gpurun_out/lane_15_64 \
  python bench.py --steps 1 --warmup 1 --kernel 3 --cells 300000 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/*.json gpurun_out/pytest_react.log
