#!/bin/bash
# round 2, call I: full GPU suite, synccheck of the tensor-memory kernel with the tcgen05.alloc result slot moved (alt build),
# default bench line with the extra configurations, --mode gi lines
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_i_pytest_gpu.log
cat gpurun_out/r02_i_pytest_gpu.log
CS=/usr/local/cuda/bin/compute-sanitizer
RXN_B200_LIB=$PWD/pflotran_b200/librxn_b200_alt.so timeout 900 $CS --tool synccheck --print-limit 2 python profiles/sanitize_run.py react 20000 > gpurun_out/r02_i_synccheck_tm_slot2.log 2>&1
grep -m3 "Barrier is located" gpurun_out/r02_i_synccheck_tm_slot2.log
( time timeout 900 python bench.py ) > gpurun_out/r02_i_bench_default.json 2> gpurun_out/r02_i_bench_default.err; tail -3 gpurun_out/r02_i_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_i_bench_default.json').read().strip().splitlines()[-1])
print('headline %.2f M/s e2e %.2f frac %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac']))
for o in d.get('other_configs', []):
    print(json.dumps(o)[:700])
PY
timeout 600 python bench.py --mode gi --steps 5 --warmup 3 > gpurun_out/r02_i_bench_gi_hpt.json 2> gpurun_out/r02_i_bench_gi_hpt.err; cut -c1-1500 gpurun_out/r02_i_bench_gi_hpt.json; tail -3 gpurun_out/r02_i_bench_gi_hpt.err
timeout 600 python bench.py --mode gi --workload hanford300a_eq --steps 5 --warmup 3 > gpurun_out/r02_i_bench_gi_300a.json 2> gpurun_out/r02_i_bench_gi_300a.err; cut -c1-1500 gpurun_out/r02_i_bench_gi_300a.json; tail -3 gpurun_out/r02_i_bench_gi_300a.err
timeout 600 python bench.py --mode gi --workload hanford300a_mr --steps 5 --warmup 3 > gpurun_out/r02_i_bench_gi_mr.json 2> gpurun_out/r02_i_bench_gi_mr.err; cut -c1-1500 gpurun_out/r02_i_bench_gi_mr.json; tail -3 gpurun_out/r02_i_bench_gi_mr.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_i_gi_launches.csv python bench.py --mode gi --workload hanford300a_eq --steps 2 --warmup 1 --cells 400000 > /dev/null 2>&1
