#!/bin/bash
# round 2, call E: full GPU suite + smoke after the fixture / reaction-type additions
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_e_pytest_gpu.log
cat gpurun_out/r02_e_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_e_smoke.log 2>&1; tail -4 gpurun_out/r02_e_smoke.log
