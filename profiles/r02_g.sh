#!/bin/bash
# round 2, call G: full GPU suite (no -x), compute-sanitizer memcheck / racecheck / synccheck over every kernel family,
# ncu capture of the default (G = 3) tensor-memory kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_g_pytest_gpu.log
cat gpurun_out/r02_g_pytest_gpu.log
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python profiles/sanitize_run.py all 40000 > gpurun_out/r02_g_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -v "^react\|^gi\|^flux" gpurun_out/r02_g_sanitizer_$tool.log | tail -6
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_tm -s 1 -c 1 -o gpurun_out/r02_g_tm_g3 \
  python bench.py --steps 1 --warmup 1 --cells 600000 > gpurun_out/r02_g_ncu_tm.log 2>&1
ls -la gpurun_out/*.ncu-rep
