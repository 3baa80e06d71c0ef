#!/bin/bash
# round 2, call AJ: compute-sanitizer memcheck + racecheck over the kernels added late in round 2 (profiles/sanitize_run.py new)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python profiles/sanitize_run.py new 20000 > gpurun_out/r02_aj_sanitizer_new_memcheck.log 2>&1; tail -4 gpurun_out/r02_aj_sanitizer_new_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python profiles/sanitize_run.py new 20000 > gpurun_out/r02_aj_sanitizer_new_racecheck.log 2>&1; tail -4 gpurun_out/r02_aj_sanitizer_new_racecheck.log
