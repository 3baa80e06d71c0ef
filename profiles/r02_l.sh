#!/bin/bash
# round 2, call L: global-implicit loops on the tensor-memory layout: GPU tests, bench --mode gi before/after, sanitizer
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "global_implicit or time_stepped or gold" 2>&1 | tail -25 > gpurun_out/r02_l_pytest.log
cat gpurun_out/r02_l_pytest.log
for wl in hanford300a_eq hanford300a_mr; do
  for k in 0 2 1; do
    RXN_GI_KERNEL=$k timeout 300 python bench.py --mode gi --workload $wl --steps 5 --warmup 3 > gpurun_out/r02_l_gi_${wl}_k$k.json 2> gpurun_out/r02_l_gi_${wl}_k$k.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_l_gi_${wl}_k$k.json').read().strip().splitlines()[-1])
    print('$wl gi_kernel=$k: %.1f M blocks/s e2e %.1f kernel_ms %.3f %s frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['bound'], d['roofline']['frac']))
except Exception as e: print('$wl $k failed', e)
PY
    tail -2 gpurun_out/r02_l_gi_${wl}_k$k.err
  done
done
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 10 python profiles/sanitize_run.py gi 40000 > gpurun_out/r02_l_sanitizer_gi_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -3 gpurun_out/r02_l_sanitizer_gi_$tool.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_l_gi_launches.csv python bench.py --mode gi --workload hanford300a_eq --steps 2 --warmup 1 --cells 400000 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gi_tm -s 2 -c 2 -o gpurun_out/r02_l_gi_tm \
  python bench.py --mode gi --workload hanford300a_eq --steps 1 --warmup 1 --cells 400000 > gpurun_out/r02_l_ncu.log 2>&1
