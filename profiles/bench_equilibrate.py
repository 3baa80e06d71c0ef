"""Throughput of rxn_equilibrate_constraint_batch (ReactionEquilibrateConstraint per cell) on the GPU against the CPU
oracle.  usage: python profiles/bench_equilibrate.py [workload] [cells]"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from pflotran_b200 import abi, synth, reactive_transport as rt
from oracle.pyoracle import Oracle
import kat

name = sys.argv[1] if len(sys.argv) > 1 else 'hanford300a_eq'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
w = synth.Workload(name); t = w.tables
ctype, conc, cid, guess, vf, area = kat.fixture_constraint(w)
rng = np.random.default_rng(5)
st = abi.HostState(t, n)
kat.fill_scalars(st, t, 0.25)
st['DEN_KG'][0] = t.reference_water_density * (1.0 + 0.01 * rng.standard_normal(n))
st['MNRL_VOLFRAC'][:] = vf[:, None]; st['MNRL_AREA'][:] = area[:, None]
concs = np.tile(conc, (n, 1))
lin = np.isin(ctype, [0, 1, 2, 7, 9])
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
concs[:, lin] *= np.exp(sigma * rng.standard_normal((n, t.naqcomp)))[:, lin]
rx = rt.Reaction(t); rz = rt.Realization(rx, n)
ms = []
for _ in range(3):
    rz.upload_host_state(st)
    basis, it, status = rz.ReactionEquilibrateConstraint(ctype, concs, cid, guess, False, bool(t.initialize_with_molality))
    ms.append(rz.last_kernel_ms())
orc = Oracle(t); m = 2000; st_o = st.copy()
t0 = time.perf_counter()
cpu_failed = 0
for c in range(m):
    try:
        orc.equilibrate(st_o, c, ctype, concs[c], cid, guess, use_prev=False)
    except RuntimeError:        # the reference's fatal errors (singular Newton matrix, zero concentration) for this constraint
        cpu_failed += 1
cpu = m / (time.perf_counter() - t0)
print(json.dumps({'workload': name, 'cells': n, 'sigma': sigma, 'kernel_ms': min(ms[1:]), 'gpu_cells_per_s': n / (min(ms[1:]) * 1e-3),
                  'mean_iterations': float(it.mean()), 'failed': int((status != 0).sum()), 'failed_status_histogram': {int(k): int(v) for k, v in zip(*np.unique(status, return_counts=True))},
                  'cpu_oracle_cells_per_s_1_thread': cpu, 'cpu_failed_of_%d' % m: cpu_failed}))
