#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "full_size or test_react or inactive" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_300a.json 2> gpurun_out/bench_300a.err
RXN_NO_PIPELINE=1 timeout 600 python bench.py --steps 3 > gpurun_out/bench_300a_nopipe.json 2> gpurun_out/bench_300a_nopipe.err
timeout 300 python bench.py --workload calcite > gpurun_out/bench_calcite.json 2> gpurun_out/bench_calcite.err
