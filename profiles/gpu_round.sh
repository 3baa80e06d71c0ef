#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -u profiles/dbg_gi.py calcite 2000 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "global_implicit or test_react" 2>&1 | tail -6 | tee gpurun_out/pytest_gi.log
for wl in hanford300a_eq hanford300a_mr calcite; do timeout 120 python profiles/bench_gi.py $wl 500000 2>&1 | tail -1 | tee gpurun_out/bench_gi_$wl.json; done
