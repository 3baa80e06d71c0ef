#!/bin/bash
# round 2, call B: first run of the tensor-memory RReact kernel: parity, G sweep, ncu counters
mkdir -p gpurun_out
RXN_LANE_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "react" 2>&1 | tail -15 > gpurun_out/r02_b_pytest_react.log
cat gpurun_out/r02_b_pytest_react.log
for g in 2 4 1; do
  RXN_TM_G=$g timeout 200 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/r02_b_bench_300a_tm_g$g.json 2> gpurun_out/r02_b_bench_300a_tm_g$g.err
  python -c "
import json
try:
  d=json.loads(open('gpurun_out/r02_b_bench_300a_tm_g$g.json').read().strip().splitlines()[-1]); print('TM G=$g', d['value']/1e6, 'M/s kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'bad', d['config']['cells_with_nonreference_flags'], d['config']['kernel'])
except Exception as e: print('G=$g failed', e); print(open('gpurun_out/r02_b_bench_300a_tm_g$g.err').read()[-2000:])
"
done
RXN_TM_G=2 timeout 300 ncu --clock-control none -k regex:k_react_tm -s 1 -c 1 --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
  --csv --log-file gpurun_out/r02_b_ncu_tm_g2.csv python bench.py --steps 1 --warmup 1 --cells 600000 > gpurun_out/r02_b_ncu_tm_g2.log 2>&1
grep k_react_tm gpurun_out/r02_b_ncu_tm_g2.csv | awk -F'","' '{print $(NF-2), $(NF)}'
