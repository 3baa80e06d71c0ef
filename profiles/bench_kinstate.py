"""Kernel time of RTUpdateKineticState (rxn_update_kinetic_state_batch: mineral volume fractions + the multirate sorbed totals,
reactive_transport.F90:692-705, reaction.F90:5320-5429) against its algorithmic HBM bytes.  usage: python profiles/bench_kinstate.py [workload] [cells]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pflotran_b200 import synth, reactive_transport as rt

name = sys.argv[1] if len(sys.argv) > 1 else 'hanford300a_mr'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
w = synth.Workload(name)
t = w.tables
cells = synth.make_cells(w, 0, n)
rx = rt.Reaction(t)
rz = rt.Realization(rx, n)
for f, v in w.base.items():
    rz.broadcast(f, v)
rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
if t.nkinmnrl:
    rz.upload('MNRL_VOLFRAC', cells['volfrac'])
xx = np.ascontiguousarray(np.tile(w.base['PRI_MOLAL'] * 1.02, (n, 1)))
rz.RTUpdateAuxVars(xx, True)
rz.RTResidualJacobianNonFlux(1800.0, jacobian=False)
ms = []
for _ in range(6):
    rz.RTUpdateKineticState(1800.0)
    ms.append(rz.last_kernel_ms())
ms = float(np.median(ms[2:]))
naq, nkin = t.naqcomp, t.nkinmnrl
nrate = t.kinmr_max_nrate if t.nkinmrsrfcplxrxn else 0
# reads: m, gamma (ln a for the mineral rates), volfrac, area, rate, scalars; multirate: S_0 and every S_r; writes: volfrac, every S_r
nbytes = 8.0 * n * (2 * naq + 4 * nkin + 7 + t.nkinmrsrfcplxrxn * naq * (1 + 2 * nrate))
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6535.7
print(json.dumps({'workload': name, 'cells': n, 'kernel_ms': ms, 'cells_per_s': n / (ms * 1e-3), 'algorithmic_bytes': nbytes,
                  'gbs': nbytes / (ms * 1e-3) / 1e9, 'hbm_frac': nbytes / (ms * 1e-3) / 1e9 / peak}))
