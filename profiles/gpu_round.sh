#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "equilibrate" 2>&1 | tail -4 | tee gpurun_out/pytest_eq.log
for wl in hanford300a_eq calcite; do timeout 200 python profiles/bench_equilibrate.py $wl 200000 2>&1 | tail -1 | tee gpurun_out/bench_eq_$wl.json; done
