#!/bin/bash
# round-end evidence: GPU test suite, smoke, bench lines (ours + reference arm), ncu launch list and full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_300a.json 2> gpurun_out/bench_300a.err
timeout 300 python bench.py --workload calcite > gpurun_out/bench_calcite.json 2> gpurun_out/bench_calcite.err
timeout 300 python profiles/bench_flux.py hanford300a_eq 128 128 64 > gpurun_out/bench_flux_300a.json 2> gpurun_out/bench_flux_300a.err
timeout 300 python profiles/bench_flux.py hanford300a_eq 100 100 100 > gpurun_out/bench_flux_300a_100.json 2> gpurun_out/bench_flux_300a_100.err
timeout 300 python profiles/bench_flux.py calcite 100 100 100 > gpurun_out/bench_flux_calcite_100.json 2> gpurun_out/bench_flux_calcite_100.err
timeout 300 python profiles/bench_flux.py calcite 256 256 128 > gpurun_out/bench_flux_calcite.json 2> gpurun_out/bench_flux_calcite.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_flux.csv \
  python profiles/bench_flux.py hanford300a_eq 96 96 48 > gpurun_out/ncu_launches_flux.log 2>&1
cat gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
