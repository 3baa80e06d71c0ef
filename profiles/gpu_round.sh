#!/bin/bash
mkdir -p gpurun_out
for lib in main alt; do
  if [ $lib = alt ]; then export RXN_B200_LIB=$PWD/pflotran_b200/librxn_b200_alt.so; fi
  timeout 200 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/bench_300a_$lib.json 2> gpurun_out/bench_300a_$lib.err
  timeout 200 python bench.py --steps 3 --warmup 3 --workload calcite --cells 4000000 > gpurun_out/bench_calcite_$lib.json 2> gpurun_out/bench_calcite_$lib.err
done
