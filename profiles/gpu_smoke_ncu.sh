#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux_residual -s 1 -c 1 -o gpurun_out/flux_res \
  python profiles/bench_flux.py hanford300a_eq 100 100 50 > gpurun_out/ncu_flux_res.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux_coefs -s 1 -c 1 -o gpurun_out/flux_coefs \
  python profiles/bench_flux.py hanford300a_eq 100 100 50 > gpurun_out/ncu_flux_coefs.log 2>&1
ls -la gpurun_out/*.ncu-rep
