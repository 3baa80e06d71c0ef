"""How much of the work order's gain survives an IMPERFECT prediction.  The library orders a call by the Newton iteration counts of the
previous call; bench.py repeats the same inputs every step, so there the prediction is exact.  Here consecutive steps alternate between two
noise realisations A / B of the same cells (same reaction-front membership, porosity and minerals; independent N(0,1) draws for every
total: synth.make_cells(variant=1)), so every call is ordered by the counts of DIFFERENT inputs - the situation of a transport run
whose totals change from step to step.  Prints one JSON line: kernel ms per step unordered / ordered from the other realisation / ordered
from the same inputs.   usage: python profiles/bench_order_prediction.py [workload] [cells]"""
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pflotran_b200 import abi, synth, reactive_transport as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'hanford300a_eq'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
w = synth.Workload(name)
t = w.tables
A = synth.make_cells(w, 0, n)
B = synth.make_cells(w, 0, n, variant=1)
assert (A['porosity'] == B['porosity']).all() and not (A['tran_xx'] == B['tran_xx']).all()
rx = rt.Reaction(t)
rz = rt.Realization(rx, n)
RESET = ['PRI_MOLAL', 'PRI_ACT_COEF', 'SEC_MOLAL', 'SEC_ACT_COEF', 'LN_ACT_H2O', 'TOTAL_SORB_EQ', 'FREE_SITE_CONC', 'EQIONX_REF_CATION_SORBED_CONC']
for f, v in w.base.items():
    rz.broadcast(f, v)
rz.set_cell_scalars(porosity=A['porosity'], temp=A['temp'], pres=A['pres'])
if t.nkinmnrl:
    rz.upload('MNRL_VOLFRAC', A['volfrac'])
nb = n * t.ncomp * 8
d_in = [rz.device_alloc(nb), rz.device_alloc(nb)]
rz.device_copy(d_in[0], A['tran_xx'], nb, 0)
rz.device_copy(d_in[1], B['tran_xx'], nb, 0)
d_xx = rz.device_alloc(nb); d_it = rz.device_alloc(n * 4); d_fl = rz.device_alloc(n * 4)
it = [np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)]


def step(k):
    for f in RESET:
        if rx.field_rows(f):
            rz.broadcast(f, w.base[f])
    rz.device_copy(d_xx, d_in[k], nb, 2)
    rz.RTReact_device(d_xx, n, 3600.0, abi.RXN_DT_CONSISTENT, 0, d_it, d_fl)
    return rz.last_kernel_ms()


def run(seq, steps=8):
    for k in seq[:3]:
        step(k)
    return statistics.mean(step(seq[i % len(seq)]) for i in range(steps))


os.environ['RXN_NO_REACT_ORDER'] = '1'
ms_off = run([0, 1])
step(0); rz.device_copy(it[0], d_it, n * 4, 1)
step(1); rz.device_copy(it[1], d_it, n * 4, 1)
del os.environ['RXN_NO_REACT_ORDER']
ms_other = run([0, 1])          # every call ordered by the counts of the other realisation
ms_same = run([0])              # ordered by the counts of the same inputs (bench.py's situation)
print(json.dumps({'workload': name, 'cells': n, 'kernel_ms_unordered': ms_off, 'kernel_ms_ordered_by_other_realisation': ms_other,
                  'kernel_ms_ordered_by_same_inputs': ms_same, 'gain_other': ms_off / ms_other, 'gain_same': ms_off / ms_same,
                  'cells_with_equal_iteration_count_in_A_and_B': float((it[0] == it[1]).mean()),
                  'mean_abs_count_difference': float(np.abs(it[0] - it[1]).mean()), 'mean_newton_iterations': float(it[0].mean())}))
