#!/bin/bash
# flux-side kernels: GPU tests, kernel times, ncu capture of the Jacobian kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flux.py tests/test_gpu_parity.py::test_global_implicit_device_resident_entry_points -x -q 2>&1 | tail -15 > gpurun_out/pytest_flux.log
timeout 300 python profiles/bench_flux.py hanford300a_eq 128 128 64 > gpurun_out/bench_flux_300a.json 2> gpurun_out/bench_flux_300a.err
timeout 300 python profiles/bench_flux.py calcite 256 256 128 > gpurun_out/bench_flux_calcite.json 2> gpurun_out/bench_flux_calcite.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux_jacobian -s 1 -c 1 -o gpurun_out/flux_jac \
  python profiles/bench_flux.py hanford300a_eq 96 96 48 > gpurun_out/ncu_flux.log 2>&1
cat gpurun_out/pytest_flux.log; cat gpurun_out/bench_flux_300a.json; tail -2 gpurun_out/bench_flux_300a.err; cat gpurun_out/bench_flux_calcite.json
