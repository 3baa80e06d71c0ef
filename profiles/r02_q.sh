#!/bin/bash
# round 2, call Q: tensor-memory kernel variants (term-stream unroll 1 / 2 / 4, LU first-row preload), kinetic-state kernel
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2: %.2f M/s kernel_ms %.3f frac %.4f' % (d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e: print('$2 failed', e)
PY
}
for v in main u4 u1 pl main; do
  if [ $v = main ]; then unset RXN_B200_LIB; else export RXN_B200_LIB=$PWD/pflotran_b200/librxn_b200_$v.so; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --cells 2000000 --no-extra > gpurun_out/r02_q_300a_$v.json 2> gpurun_out/r02_q_300a_$v.err; show gpurun_out/r02_q_300a_$v.json "300A 2e6 $v"
done
unset RXN_B200_LIB
timeout 300 python profiles/bench_kinstate.py hanford300a_mr 500000 | tee gpurun_out/r02_q_kinstate_mr.json
timeout 300 python profiles/bench_kinstate.py hanford300a_eq 2000000 | tee gpurun_out/r02_q_kinstate_300a.json
