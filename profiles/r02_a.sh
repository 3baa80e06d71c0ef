#!/bin/bash
# round 2, call A: TMEM microbenchmark; react parity with the transposed lane mapping; G sweep; bank-conflict counters
mkdir -p gpurun_out
timeout 120 ./profiles/ubench_tmem > gpurun_out/r02_ubench_tmem.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "react" 2>&1 | tail -5 > gpurun_out/r02_a_pytest_react.log
for g in 2 4 1; do
  RXN_LANE_G=$g timeout 200 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/r02_a_bench_300a_g$g.json 2> gpurun_out/r02_a_bench_300a_g$g.err
done
RXN_LANE_G=2 timeout 300 ncu --clock-control none -k regex:k_react_lane -s 1 -c 1 --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio \
  --csv --log-file gpurun_out/r02_a_ncu_g2.csv python bench.py --steps 1 --warmup 1 --cells 600000 > gpurun_out/r02_a_ncu_g2.log 2>&1
cat gpurun_out/r02_ubench_tmem.txt; cat gpurun_out/r02_a_pytest_react.log
for g in 2 4 1; do python -c "
import json,sys
try:
  d=json.loads(open('gpurun_out/r02_a_bench_300a_g$g.json').read().strip().splitlines()[-1]); print('G=$g', d['value']/1e6, 'M/s kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], d['config']['kernel'])
except Exception as e: print('G=$g failed', e)
"; done
tail -12 gpurun_out/r02_a_ncu_g2.csv | cut -c1-400
