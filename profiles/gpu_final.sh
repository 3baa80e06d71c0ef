#!/bin/bash
# round-end evidence: GPU test suite, smoke, bench lines (ours + reference arm), ncu launch list and full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_300a.json 2> gpurun_out/bench_300a.err
timeout 300 python bench.py --workload calcite > gpurun_out/bench_calcite.json 2> gpurun_out/bench_calcite.err
timeout 300 python bench.py --workload hanford300a_mr > gpurun_out/bench_mr.json 2> gpurun_out/bench_mr.err
timeout 300 python bench.py --workload hpt_calcite > gpurun_out/bench_hpt.json 2> gpurun_out/bench_hpt.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --cells 2000000 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_react_lane -s 2 -c 1 -o gpurun_out/lane_final \
  python bench.py --steps 1 --warmup 1 --cells 600000 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
