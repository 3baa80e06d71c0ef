#!/bin/bash
# round 2, call AW: work order under imperfect prediction at rank-sized batches and on the tail-bound chemistry
mkdir -p gpurun_out
for a in "hanford300a_eq 50000" "hanford300a_eq 100000" "hanford300a_eq 300000" "hanford300a_eq 1000000" "calcite 100000" "ascem 200000" "ascem 1000000"; do
  set -- $a
  timeout 900 python profiles/bench_order_prediction.py $1 $2 > gpurun_out/r02_aw_order_prediction_$1_$2.json 2>> gpurun_out/r02_aw.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_aw_order_prediction_$1_$2.json').read().strip().splitlines()[-1])
print('$1 $2: unordered %.3f ms  by other realisation %.3f (x%.3f)  by same inputs %.3f (x%.3f)  equal counts %.2f' % (d['kernel_ms_unordered'], d['kernel_ms_ordered_by_other_realisation'], d['gain_other'], d['kernel_ms_ordered_by_same_inputs'], d['gain_same'], d['cells_with_equal_iteration_count_in_A_and_B']))
PY
done
tail -2 gpurun_out/r02_aw.err
