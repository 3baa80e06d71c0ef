"""Parity soak: RTReact on the GPU (through the C ABI) against the CPU oracle on fresh random cells (a seed the test-suite does not
use) at sizes well beyond the tests' batches.  Per workload: cells compared, cells whose Newton iteration count or exit flags differ,
largest relative deviation of the converged free-ion molalities and of the mineral volume fractions.  Test infrastructure (imports
oracle/); prints one JSON line per workload.   usage: python profiles/parity_soak.py [seed] [scale]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from pflotran_b200 import abi, synth, reactive_transport as rt  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 77001
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
only = set(sys.argv[3:])
PLAN = [('hanford300a_eq', 2_000_000), ('hanford300a_mr', 300_000), ('calcite', 4_000_000), ('hpt_calcite', 4_000_000),
        ('scco2_brine', 2_000_000), ('surface_complexation', 500_000), ('ion_exchange', 500_000), ('hanford300a_stoich', 300_000),
        ('hanford300a_kinsrf', 300_000), ('hanford300a_act_newton', 200_000), ('mineral_prefactor', 200_000),
        ('abcd_microbial', 2_000_000), ('abcd_microbial_inhibition', 1_000_000), ('general_reaction', 2_000_000), ('ascem', 60_000)]
threads = os.cpu_count() or 1
for name, n in PLAN:
    if only and name not in only:
        continue
    n = max(4096, int(n * scale))
    w = synth.Workload(name)
    cells = synth.make_cells(w, 0, n, seed=seed)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    rz.upload_host_state(st_g)
    xg = cells['tran_xx'].copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0, abi.RXN_DT_CONSISTENT)
    info = rz.react_kernel_info()
    rz.download_host_state(st_g)
    xo = cells['tran_xx'].copy()
    t0 = time.perf_counter()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=threads)
    el = time.perf_counter() - t0
    # as tests/common.py iteration_parity: a cell whose arithmetic left the finite range in BOTH (the reference would spin or abort there)
    # is compared on the iteration count and the NONFINITE flag only (whether the closing free-site loop then also hits its guard
    # depends on which garbage value the singular step produced)
    nf = ((fl_g & abi.RXN_FLAG_NONFINITE) != 0) & ((fl_o & abi.RXN_FLAG_NONFINITE) != 0)
    same = (it_g == it_o) & ((fl_g == fl_o) | (nf & (((fl_g ^ fl_o) & ~abi.RXN_FLAG_CAPPED) == 0)))
    strict = (it_g == it_o) & (fl_g == fl_o)
    ok = ((fl_o & ~3) == 0) & same
    err = np.abs(xg[ok] - xo[ok]) / np.abs(xo[ok])
    out = {'workload': name, 'seed': seed, 'cells': n, 'iteration_or_flag_mismatches': int((~same).sum()), 'nonfinite_in_both_with_other_guard_bit': int((same & ~strict).sum()),
           'cells_with_nonreference_flags': int(((fl_o & ~3) != 0).sum()), 'converged_cells_compared': int(ok.sum()),
           'max_rel_err_free_ion': float(err.max()) if err.size else None,
           'cells_above_1e-10': int((err.max(axis=1) > 1e-10).sum()) if err.size else 0,
           'mean_newton_iterations': float(it_o.mean()), 'max_newton_iterations': int(it_o.max()), 'kernel': info[:70], 'oracle_s': round(el, 1)}
    if out['iteration_or_flag_mismatches'] or out['cells_above_1e-10']:
        # the yardstick of tests/common.py (PerturbedOracle): the oracle itself on inputs changed in the last bit
        sign = np.where(np.random.default_rng(11).random(cells['tran_xx'].shape) < 0.5, -1.0, 1.0)
        st_p = synth.host_state(w, cells)
        xp = cells['tran_xx'] * (1.0 + 2.2e-16 * sign)
        it_p, fl_p = Oracle(w.tables).react(st_p, xp, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=threads)
        same_p = (it_p == it_o) & (fl_p == fl_o)
        okp = ((fl_o & ~3) == 0) & same_p
        errp = np.abs(xp[okp] - xo[okp]) / np.abs(xo[okp])
        out['oracle_vs_last_bit_perturbed_oracle'] = {'iteration_or_flag_mismatches': int((~same_p).sum()),
                                                      'max_rel_err_free_ion': float(errp.max()) if errp.size else None,
                                                      'cells_above_1e-10': int((errp.max(axis=1) > 1e-10).sum()) if errp.size else 0}
    if w.tables.nkinmnrl:
        a, b = st_g['MNRL_VOLFRAC'][:, ok], st_o['MNRL_VOLFRAC'][:, ok]
        m = b != 0
        out['max_rel_err_mnrl_volfrac'] = float((np.abs(a - b)[m] / np.abs(b[m])).max()) if m.any() else 0.0
    print(json.dumps(out), flush=True)
    del rz, rx
