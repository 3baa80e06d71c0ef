#!/bin/bash
# round 2, call AR: compute-sanitizer memcheck over the chunked host-buffer RTReact (two compute streams); full GPU suite at the final source state
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python profiles/sanitize_run.py pipeline > gpurun_out/r02_ar_sanitizer_pipeline_memcheck.log 2>&1; tail -4 gpurun_out/r02_ar_sanitizer_pipeline_memcheck.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_ar_pytest_gpu.log
cat gpurun_out/r02_ar_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_ar_smoke.log 2>&1; tail -3 gpurun_out/r02_ar_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ar_launches.csv python bench.py --steps 2 --warmup 1 --cells 1000000 --no-extra > /dev/null 2>&1
