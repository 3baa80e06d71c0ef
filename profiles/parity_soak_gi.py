"""Parity soak of the global-implicit entry points: RTUpdateAuxVars (+ activity update), RTUpdateFixedAccumulation and the accumulation /
reaction residual and Jacobian blocks on the GPU (through the C ABI) against the CPU oracle, on fresh random cells (a seed the
test-suite does not use) at sizes well beyond the tests' batches.  Errors are measured on the scales of tests/common.py (the terms of
each difference).  Test infrastructure (imports oracle/).   usage: python profiles/parity_soak_gi.py [seed] [scale]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from pflotran_b200 import synth, reactive_transport as rt  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from common import accumulation_scale, residual_scale, jacobian_scale, STATE_FIELDS  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 77002
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
PLAN = [('hanford300a_eq', 500_000), ('hanford300a_mr', 150_000), ('hpt_calcite', 2_000_000), ('calcite', 2_000_000), ('scco2_brine', 1_000_000),
        ('surface_complexation', 300_000), ('ion_exchange', 300_000), ('abcd_microbial', 1_000_000), ('general_reaction', 1_000_000),
        ('hanford300a_kinsrf', 150_000)]
threads = os.cpu_count() or 1
for name, n in PLAN:
    n = max(4096, int(n * scale))
    w = synth.Workload(name)
    cells = synth.make_cells(w, 0, n, seed=seed)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    rz.upload_host_state(st_g)
    orc = Oracle(w.tables)
    rng = np.random.default_rng(seed)
    xx = np.ascontiguousarray(w.base_solution()[None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
    orc.update_auxvars(st_o, xx, True, nthreads=threads)
    rz.RTUpdateAuxVars(xx, True)
    rz.download_host_state(st_g)
    worst_state = 0.0
    for f in ('PRI_MOLAL', 'PRI_ACT_COEF', 'SEC_ACT_COEF', 'TOTAL_SORB_EQ', 'FREE_SITE_CONC', 'IMMOBILE'):
        a, b = st_g[f], st_o[f]
        if a.size:
            sc = np.maximum(np.abs(b), 1e-13 * np.abs(b).max(axis=1, keepdims=True))
            worst_state = max(worst_state, float((np.abs(a - b) / np.maximum(sc, 1e-300)).max()))
    a_o = orc.fixed_accum(st_o, xx, nthreads=threads)
    a_g = rz.RTUpdateFixedAccumulation(xx)
    e_acc = float((np.abs(a_g - a_o) / np.maximum(accumulation_scale(st_o, w.tables, a_o), 1e-300)).max())
    r_o, j_o = orc.residual_jacobian(st_o, 1800.0, nthreads=threads)
    r_g, j_g = rz.RTResidualJacobianNonFlux(1800.0)
    e_res = float((np.abs(r_g - r_o) / np.maximum(residual_scale(st_o, w.tables, r_o, a_o, 1800.0), 1e-300)).max())
    e_jac = float((np.abs(j_g - j_o) / np.maximum(jacobian_scale(st_o, j_o, w.ncomp), 1e-300)).max())
    print(json.dumps({'workload': name, 'seed': seed, 'cells': n, 'max_err_state_after_RTUpdateAuxVars': worst_state, 'max_err_fixed_accumulation': e_acc,
                      'max_err_residual_blocks': e_res, 'max_err_jacobian_blocks': e_jac, 'bar': 1e-10}), flush=True)
    del rz, rx
