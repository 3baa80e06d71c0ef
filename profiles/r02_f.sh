#!/bin/bash
# round 2, call F: full GPU suite + smoke at the current source state, G sweep of the tensor-memory kernel, the BASELINE
# configurations, launch list and one full ncu capture (with source) of the default kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_f_pytest_gpu.log
cat gpurun_out/r02_f_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_f_smoke.log 2>&1; tail -4 gpurun_out/r02_f_smoke.log
show() { python -c "
import json,sys
try:
  d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', '%.2f M/s e2e %.2f kernel_ms %.2f frac %.4f bad %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['cells_with_nonreference_flags']), d['config']['kernel'][:60])
except Exception as e: print('$2 failed', e)
"; }
for g in 2 3 4; do
  RXN_TM_G=$g timeout 200 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/r02_f_300a_tm_g$g.json 2> gpurun_out/r02_f_300a_tm_g$g.err; show gpurun_out/r02_f_300a_tm_g$g.json "300A 2e6 TM G=$g"
done
for g in 3 4; do
  RXN_TM_G=$g timeout 200 python bench.py --steps 3 --warmup 3 --workload hanford300a_mr --cells 1000000 > gpurun_out/r02_f_mr_tm_g$g.json 2> gpurun_out/r02_f_mr_tm_g$g.err; show gpurun_out/r02_f_mr_tm_g$g.json "mr 1e6 TM G=$g"
done
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_f_bench_300a.json 2> gpurun_out/r02_f_bench_300a.err; show gpurun_out/r02_f_bench_300a.json "300A 1e7 default"
RXN_TM_G=3 timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_f_bench_300a_g3.json 2> gpurun_out/r02_f_bench_300a_g3.err; show gpurun_out/r02_f_bench_300a_g3.json "300A 1e7 G=3"
timeout 300 python bench.py --steps 5 --warmup 3 --workload calcite > gpurun_out/r02_f_bench_calcite.json 2> gpurun_out/r02_f_bench_calcite.err; show gpurun_out/r02_f_bench_calcite.json "calcite 1e6"
timeout 300 python bench.py --steps 5 --warmup 3 --workload hpt_calcite > gpurun_out/r02_f_bench_hpt.json 2> gpurun_out/r02_f_bench_hpt.err; show gpurun_out/r02_f_bench_hpt.json "hpt 1e6"
# launch list of the default bench command (per-launch times are cold-cache and serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_f_launches.csv \
  python bench.py --steps 2 --warmup 1 --cells 1000000 > gpurun_out/r02_f_launches.log 2>&1
# full capture with source of the default kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_react_tm -s 1 -c 1 -o gpurun_out/r02_f_tm \
  python bench.py --steps 1 --warmup 1 --cells 600000 > gpurun_out/r02_f_ncu_tm.log 2>&1
ls -la gpurun_out/*.ncu-rep
