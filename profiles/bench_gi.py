"""Kernel time of the global-implicit residual/Jacobian block entry point (rxn_residual_jacobian_blocks_batch):
resident-lane layout vs thread-per-cell (RXN_GI_KERNEL=1).  usage: python profiles/bench_gi.py [workload] [cells]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pflotran_b200 import synth, reactive_transport as rt

name = sys.argv[1] if len(sys.argv) > 1 else 'hanford300a_eq'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
w = synth.Workload(name)
cells = synth.make_cells(w, 0, n)
out = {'workload': name, 'cells': n}
for label, env in (('resident_lane', '0'), ('thread_per_cell', '1')):
    os.environ['RXN_GI_KERNEL'] = env
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    for f, v in w.base.items():
        rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    if w.tables.nkinmnrl:
        rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
    rz.RTUpdateAuxVars(xx, True)
    ms = []
    for _ in range(4):
        r, j = rz.RTResidualJacobianNonFlux(1800.0)
        ms.append(rz.last_kernel_ms())
    out[label] = {'kernel_ms': min(ms[1:]), 'blocks_per_s': n / (min(ms[1:]) * 1e-3), 'checksum': float(np.abs(j).sum())}
    del rz, rx
print(json.dumps(out))
