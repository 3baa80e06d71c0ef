"""GPU parity tests: the CUDA library, called through the C ABI, against the CPU oracle on the
same seeded synthetic inputs.  Bar (BASELINE.json north_star): relative 1e-10 on converged
free-ion and mineral concentrations, identical Newton iteration counts and exit flags."""
import numpy as np
import pytest

from pflotran_b200 import abi, synth, reactive_transport as rt
from oracle.pyoracle import Oracle
from common import PerturbedOracle, iteration_parity, free_ion_parity, assert_state_close, workload_cells, RTOL, rel_err, total_magnitude, accumulation_scale, residual_scale, jacobian_scale

pytestmark = pytest.mark.gpu

WORKLOADS = ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite', 'ion_exchange', 'surface_complexation',
             'calcite_kinetics', 'kd_wo_mineral']
# fixtures that reach the branches no reference batch deck exercises (tests/golden/make_fixtures.py: VARIANTS and the
# prefactor / non-isothermal / 22-primary decks): NEWTON activity algorithm + activity of water, free-site inner Newton,
# Langmuir / Freundlich isotherms, Temkin / scale factor / affinity power / threshold / rate limiter / Arrhenius, mineral
# prefactors, 5-term logK fit per cell, BASELINE config 1 (22 primaries / 164 complexes), general (forward / backward rate)
# reactions, radioactive decay, kinetic surface complexation
BRANCH_WORKLOADS = ['hanford300a_act_newton', 'hanford300a_stoich', 'kd_langmuir', 'kd_freundlich', 'calcite_rate_laws', 'mineral_prefactor', 'calcite_fit5', 'ascem', 'general_reaction', 'decay_ab', 'hanford300a_kinsrf',
                    'abcd_microbial', 'abcd_microbial_act_high', 'ab_microbial_linear', 'scco2_brine', 'abcd_microbial_inhibition']   # RMicrobial, immobile dofs, RImmobileDecay
WORKLOADS = WORKLOADS + BRANCH_WORKLOADS
GI_WORKLOADS = ['calcite', 'hanford300a_mr', 'hpt_calcite', 'ion_exchange', 'surface_complexation'] + BRANCH_WORKLOADS


def _gpu_state(w, st):
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, st.ncells)
    rz.upload_host_state(st)
    return rx, rz


@pytest.mark.parametrize('name', WORKLOADS)
@pytest.mark.parametrize('dt,mode', [(3600.0, abi.RXN_DT_CONSISTENT), (1.0, abi.RXN_DT_AS_WRITTEN)])
@pytest.mark.parametrize('kernel', [1, 3])
def test_react(name, dt, mode, kernel):
    n = 5000
    w, cells = workload_cells(name, n)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    try:
        rz.set_react_kernel(kernel)
        xo = cells['tran_xx'].copy()
        xg = xo.copy()
        it_g, fl_g = rz.RTReact(xg, dt, mode)
    except rt.RxnError as e:
        if kernel == 3 and e.status == abi.RXN_ERR_UNSUPPORTED:
            pytest.skip('shared-memory kernel not available for these tables: %s' % e)
        raise
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, dt, mode, maxit=10000, nthreads=8)
    rz.download_host_state(st_g)
    pert = PerturbedOracle(w, cells, dt, mode)
    same = iteration_parity(it_g, fl_g, it_o, fl_o, pert)
    ok = ((fl_o & ~3) == 0) & same
    good = free_ion_parity(xg, xo, ok, pert)
    assert_state_close(st_g, st_o, cells=good, what=name, tables=w.tables)


@pytest.mark.parametrize('name,G', [('hanford300a_eq', 1), ('hanford300a_eq', 2), ('hanford300a_eq', 4), ('hanford300a_mr', 2)])
def test_react_resident_lane_group_widths(name, G, monkeypatch):
    """Resident-lane kernel with 1, 2 and 4 lanes per cell (RXN_LANE_G picks the compiled shape; the multirate vectors fit
    the G = 2 shapes only since the ablation shapes were pruned), enough cells that every lane group takes several cells from
    the work counter."""
    monkeypatch.setenv('RXN_LANE_G', str(G))
    monkeypatch.setenv('RXN_TM', '0')          # the shared-memory-J kernel (the tensor-memory kernel is the default for N <= 15)
    n = 30000
    w, cells = workload_cells(name, n)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    rz.set_react_kernel(3)
    assert 'lanes/cell=%d' % G in rz.react_kernel_info()
    xo = cells['tran_xx'].copy()
    xg = xo.copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0, abi.RXN_DT_CONSISTENT)
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=8)
    rz.download_host_state(st_g)
    assert (it_o == it_g).all() and (fl_o == fl_g).all()
    ok = (fl_o & ~3) == 0
    assert rel_err(xg[ok], xo[ok]).max() <= RTOL
    assert_state_close(st_g, st_o, cells=np.where(ok)[0], what=name, tables=w.tables)


@pytest.mark.parametrize('G', [1, 2, 3, 4])
@pytest.mark.parametrize('name', ['hanford300a_eq', 'hanford300a_mr'])
def test_react_tensor_memory_kernel(name, G, monkeypatch):
    """Tensor-memory kernel (J in TMEM, rxn_tm_dev.cuh) with 1 to 4 member warps per cell; enough cells that every lane
    takes several cells from the work counter and cells of one warp sit in different Newton iterations."""
    if G == 1 and name == 'hanford300a_mr':
        pytest.skip('G = 1 is compiled for 128 cells per CTA only (ablation shape); the multirate vectors need 96')
    monkeypatch.setenv('RXN_TM_G', str(G))
    n = 40000
    w, cells = workload_cells(name, n)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    rz.set_react_kernel(3)
    assert 'tensor-memory' in rz.react_kernel_info() and 'warps/cell=%d' % G in rz.react_kernel_info()
    xo = cells['tran_xx'].copy()
    xg = xo.copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0, abi.RXN_DT_CONSISTENT)
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=8)
    rz.download_host_state(st_g)
    assert (it_o == it_g).all() and (fl_o == fl_g).all()
    ok = (fl_o & ~3) == 0
    assert rel_err(xg[ok], xo[ok]).max() <= RTOL
    assert_state_close(st_g, st_o, cells=np.where(ok)[0], what=name, tables=w.tables)


@pytest.mark.parametrize('name,N', [('hanford300a_eq', 16), ('hanford300a_eq', 24), ('calcite', 8), ('calcite', 12)])
def test_react_resident_lane_padded_shapes(name, N, monkeypatch):
    """naq smaller than the compiled matrix dimension (RXN_LANE_N forces a padded shape): same answers."""
    monkeypatch.setenv('RXN_LANE_N', str(N))
    n = 20000
    w, cells = workload_cells(name, n)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    rz.set_react_kernel(3)
    assert 'N=%d ' % N in rz.react_kernel_info()
    xo = cells['tran_xx'].copy()
    xg = xo.copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0, abi.RXN_DT_CONSISTENT)
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=8)
    rz.download_host_state(st_g)
    assert (it_o == it_g).all() and (fl_o == fl_g).all()
    ok = (fl_o & ~3) == 0
    assert rel_err(xg[ok], xo[ok]).max() <= RTOL
    assert_state_close(st_g, st_o, cells=np.where(ok)[0], what=name, tables=w.tables)


@pytest.mark.parametrize('kernel', [1, 3])
def test_react_iteration_cap_takes_the_closing_pass(kernel, monkeypatch):
    """Abnormal exit (GPU-only iteration cap, where the reference would spin): pri_molal moved after the last RTotal,
    so the closing RTAuxVarCompute has to redo the speciation; every kernel must agree with the oracle's capped run."""
    monkeypatch.setenv('RXN_MAX_NEWTON_ITERATIONS', '5')
    n = 6000
    w, cells = workload_cells('hanford300a_eq', n)
    st_o = synth.host_state(w, cells)
    st_o.active[::11] = 0
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    rz.set_cell_scalars(active=st_g.active)
    rz.set_react_kernel(kernel)
    xo = cells['tran_xx'].copy()
    xg = xo.copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0, abi.RXN_DT_CONSISTENT)
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=5, nthreads=8)
    rz.download_host_state(st_g)
    assert (fl_g & abi.RXN_FLAG_CAPPED).any() and (fl_g[::11] == abi.RXN_FLAG_INACTIVE).all()
    assert (it_o == it_g).all() and (fl_o == fl_g).all()
    # cells that converged within the cap: the parity bar; cells stopped mid-transient (5 Newton steps from the initial
    # guess, Jacobians far from the solution): the same algorithm, but rounding differences are amplified by the
    # transient's conditioning (measured up to 6e-6), so only a sanity bound applies to their unconverged iterate
    conv = np.where((st_o.active != 0) & ((fl_o == abi.RXN_EXIT_RESIDUAL) | (fl_o == abi.RXN_EXIT_REL_CHANGE)))[0]
    capped = np.where((st_o.active != 0) & (fl_o == abi.RXN_FLAG_CAPPED))[0]
    assert len(conv) > 100 and len(capped) > 100
    assert rel_err(xg[conv], xo[conv]).max() <= RTOL
    assert_state_close(st_g, st_o, cells=conv, what='converged under the cap', tables=w.tables)
    assert rel_err(xg[capped], xo[capped]).max() <= 1.0e-3


@pytest.mark.parametrize('name', GI_WORKLOADS)
def test_global_implicit_entry_points(name):
    n = 3000
    w, cells = workload_cells(name, n)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    orc = Oracle(w.tables)
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base_solution()[None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
    orc.update_auxvars(st_o, xx, True, nthreads=8)
    rz.RTUpdateAuxVars(xx, True)
    rz.download_host_state(st_g)
    assert_state_close(st_g, st_o, what=name + ' RTUpdateAuxVars', tables=w.tables)
    a_o = orc.fixed_accum(st_o, xx, nthreads=8)
    a_g = rz.RTUpdateFixedAccumulation(xx)
    # accumulation = phi*s*1000*V*total (+ sorbed*V), reaction.F90:5072-5148: compared on the scale of total's terms
    a_scale = accumulation_scale(st_o, w.tables, a_o)
    assert (np.abs(a_g - a_o) / np.maximum(a_scale, 1e-300)).max() <= RTOL
    r_o, j_o = orc.residual_jacobian(st_o, 1800.0, nthreads=8)
    r_g, j_g = rz.RTResidualJacobianNonFlux(1800.0)
    rs = residual_scale(st_o, w.tables, r_o, a_o, 1800.0)
    assert (np.abs(r_g - r_o) / np.maximum(rs, 1e-300)).max() <= RTOL
    js = jacobian_scale(st_o, j_o, w.ncomp)
    assert (np.abs(j_g - j_o) / np.maximum(js, 1e-300)).max() <= RTOL
    orc.update_kinetic_state(st_o, 1800.0, nthreads=8)
    rz.RTUpdateKineticState(1800.0)
    rz.download_host_state(st_g)
    assert_state_close(st_g, st_o, what=name + ' RTUpdateKineticState', tables=w.tables, kinetic_dt=1800.0)


@pytest.mark.parametrize('name', ['hanford300a_eq', 'hanford300a_mr', 'hanford300a_stoich', 'calcite'])
@pytest.mark.parametrize('gi_kernel', [0, 1])
def test_global_implicit_auxvars_with_derivative_blocks(name, gi_kernel, monkeypatch):
    """RTUpdateAuxVars with dtotal / dtotal_sorb_eq materialised (what the flux Jacobian consumes), with the activity update and
    with the state's lagged activity coefficients, then the fixed accumulation through an l2g map: the tensor-memory layout
    (gi_kernel 0, where the tables allow it: k_gi_tm) and the thread-per-cell kernels (1) against the oracle."""
    monkeypatch.setenv('RXN_GI_KERNEL', str(gi_kernel))
    n = 20000
    w, cells = workload_cells(name, n)
    st_o = synth.host_state(w, cells)
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    rz.materialize('DTOTAL')
    if rx.field_rows('DTOTAL_SORB_EQ'):
        rz.materialize('DTOTAL_SORB_EQ')
    orc = Oracle(w.tables)
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
    for update, x in ((True, xx), (False, np.ascontiguousarray(xx * np.exp(0.05 * rng.standard_normal(xx.shape))))):
        orc.update_auxvars(st_o, x, update, nthreads=8)
        rz.RTUpdateAuxVars(x, update)
        rz.download_host_state(st_g)
        assert_state_close(st_g, st_o, what='%s RTUpdateAuxVars(%s)' % (name, update), tables=w.tables)
        for f in ('DTOTAL', 'DTOTAL_SORB_EQ'):
            if rx.field_rows(f):
                d_g, d_o = rz.download(f), st_o[f]
                sc = np.maximum(np.abs(d_o), 1e-12 * np.abs(d_o).max(axis=0, keepdims=True))
                assert (np.abs(d_g - d_o) / np.maximum(sc, 1e-300)).max() <= RTOL, f
    l2g = np.ascontiguousarray(rng.permutation(n)[:n // 3].astype(np.int32))
    xa = np.ascontiguousarray(xx[l2g])
    a_g = rz.RTUpdateFixedAccumulation(xa, l2g)
    st_sub = st_o.copy()
    a_o = orc.fixed_accum(st_sub, np.ascontiguousarray(np.where(np.isin(np.arange(n), l2g)[:, None], xx, st_o['PRI_MOLAL'].T)), nthreads=8)
    a_scale = np.maximum(np.abs(a_o), (st_sub['POROSITY'] * st_sub['SAT'] * 1000.0 * st_sub['VOLUME'] * total_magnitude(st_sub, w.tables)).T)
    assert (np.abs(a_g - a_o[l2g]) / np.maximum(a_scale[l2g], 1e-300)).max() <= RTOL


@pytest.mark.parametrize('gi_kernel', [0, 2, 1])
def test_global_implicit_blocks_inactive_cells_and_l2g(gi_kernel, monkeypatch):
    """Residual / Jacobian blocks through an l2g map with inactive cells, on the tensor-memory layout (0), the resident lanes (2)
    and one thread per cell (1): blocks of inactive cells stay as the call zeroed them, every other block meets the oracle; a batch
    that does not fill the last round of a CTA (lanes beyond the batch walk a valid cell with their stores off)."""
    monkeypatch.setenv('RXN_GI_KERNEL', str(gi_kernel))
    n = 128 * 148 + 77                                     # one full round of k_gi_tm on a B200 + a ragged tail
    w, cells = workload_cells('hanford300a_eq', n)
    st_o = synth.host_state(w, cells)
    st_o.active = np.ones(n, dtype=np.uint8)
    st_o.active[::13] = 0
    st_g = st_o.copy()
    rx, rz = _gpu_state(w, st_g)
    rz.set_cell_scalars(active=st_o.active)
    orc = Oracle(w.tables)
    rng = np.random.default_rng(5)
    xx = np.ascontiguousarray(w.base_solution()[None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
    orc.update_auxvars(st_o, xx, True, nthreads=8)
    rz.RTUpdateAuxVars(xx, True)
    l2g = np.ascontiguousarray(rng.permutation(n)[:n - 1000].astype(np.int32))
    r_g, j_g = rz.RTResidualJacobianNonFlux(900.0, l2g=l2g)
    a_o = orc.fixed_accum(st_o.copy(), xx, nthreads=8)
    r_o, j_o = orc.residual_jacobian(st_o, 900.0, nthreads=8)
    dead = st_o.active[l2g] == 0
    assert dead.any() and (r_g[dead] == 0).all() and (j_g[dead] == 0).all()
    live = ~dead
    rs = residual_scale(st_o, w.tables, r_o, a_o, 900.0)
    assert (np.abs(r_g[live] - r_o[l2g[live]]) / np.maximum(rs[l2g[live]], 1e-300)).max() <= RTOL
    js = jacobian_scale(st_o, j_o, w.ncomp)
    assert (np.abs(j_g[live] - j_o[l2g[live]]) / np.maximum(js[l2g[live]], 1e-300)).max() <= RTOL


def test_global_implicit_entry_points_report_failed_cells():
    """A cell whose global-implicit evaluation is not finite (or raises a flag the reference stops on) makes the entry point
    return RXN_ERR_CELL_FAILED instead of handing NaN residuals to the caller; an out-of-range l2g entry is RXN_ERR_INVALID."""
    w, cells = workload_cells('calcite', 256)
    st = synth.host_state(w, cells)
    rx, rz = _gpu_state(w, st)
    xx = np.ascontiguousarray(np.tile(w.base['PRI_MOLAL'], (256, 1)))
    rz.RTUpdateAuxVars(xx, True)                       # healthy state: no error
    rz.RTResidualJacobianNonFlux(1800.0)
    xx[17, 0] = np.nan
    with pytest.raises(rt.RxnError) as e:
        rz.RTUpdateAuxVars(xx, True)
    assert e.value.status == abi.RXN_ERR_CELL_FAILED and 'non-finite' in str(e.value)
    with pytest.raises(rt.RxnError) as e:
        rz.RTResidualJacobianNonFlux(1800.0)
    assert e.value.status == abi.RXN_ERR_CELL_FAILED
    xx[17, 0] = w.base['PRI_MOLAL'][0]
    rz.upload_host_state(st)                           # the NaN cell poisoned its lagged sec_molal (as it would in the reference)
    rz.RTUpdateAuxVars(xx, True)                       # the flag word is per call
    l2g = np.array([0, 5, 256], dtype=np.int32)
    with pytest.raises(rt.RxnError) as e:
        rz.RTUpdateFixedAccumulation(np.ascontiguousarray(xx[:3]), l2g)
    assert e.value.status == abi.RXN_ERR_INVALID and 'l2g' in str(e.value)


def test_state_roundtrip_layouts():
    w, cells = workload_cells('hanford300a_eq', 1000)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, 1000)
    rng = np.random.default_rng(3)
    a = rng.random((w.ncomp, 1000))
    rz.upload('PRI_MOLAL', a)
    np.testing.assert_array_equal(rz.download('PRI_MOLAL'), a)
    np.testing.assert_array_equal(rz.download_aos('PRI_MOLAL'), a.T)
    rz.upload_aos('TOTAL', np.ascontiguousarray(a.T))
    np.testing.assert_array_equal(rz.download('TOTAL'), a)
    rz.broadcast('SEC_MOLAL', w.base['SEC_MOLAL'])
    s = rz.download('SEC_MOLAL')
    assert (s == w.base['SEC_MOLAL'][:, None]).all()
    # reference initial values (reactive_transport_aux.F90:213-400)
    rz2 = rt.Realization(rx, 33)
    assert (rz2.download('PRI_ACT_COEF') == 1.0).all() and (rz2.download('FREE_SITE_CONC') == 1.0e-9).all()
    assert (rz2.download('PRI_MOLAL') == 0.0).all()


def test_inactive_cells_l2g_and_empty():
    w, cells = workload_cells('calcite', 64)
    st = synth.host_state(w, cells)
    st.active[::5] = 0
    st_o = st.copy()
    rx, rz = _gpu_state(w, st)
    xg = cells['tran_xx'].copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0)
    xo = cells['tran_xx'].copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0)
    assert (it_o == it_g).all() and (fl_o == fl_g).all()
    assert (fl_g[::5] == abi.RXN_FLAG_INACTIVE).all()
    np.testing.assert_array_equal(xg[::5], cells['tran_xx'][::5])
    # l2g map + empty batch
    st2 = synth.host_state(w, cells)
    rx2, rz2 = _gpu_state(w, st2)
    l2g = np.arange(55, 39, -1, dtype=np.int32)
    x2 = np.ascontiguousarray(cells['tran_xx'][l2g])
    it2, fl2 = rz2.RTReact(x2, 3600.0, l2g=l2g)
    st_r = synth.host_state(w, cells)
    xr = cells['tran_xx'].copy()
    itr, flr = Oracle(w.tables).react(st_r, xr, 3600.0)
    assert (it2 == itr[l2g]).all() and rel_err(x2, xr[l2g]).max() <= RTOL
    e = np.zeros((0, w.ncomp))
    it0, fl0 = rz2.RTReact(e, 3600.0)
    assert it0.shape == (0,)


def test_unsupported_tables_rejected():
    w = synth.Workload('calcite')
    d = abi.make_desc(w.tables)
    d.nactive_gas = 1                                  # a reaction type outside the path (RTotalGas / RTotalCO2)
    with pytest.raises(rt.RxnError) as e:
        rt.Reaction(d)
    assert e.value.status == abi.RXN_ERR_UNSUPPORTED
    d.nactive_gas = 0
    d.ngeneral_rxn = 1                                 # a supported reaction type without its tables
    with pytest.raises(rt.RxnError) as e:
        rt.Reaction(d)
    assert e.value.status == abi.RXN_ERR_INVALID


@pytest.mark.parametrize('name,n', [('calcite', 1_000_000), ('hanford300a_eq', 200_000)])
def test_full_size_sampled_against_oracle(name, n):
    """BASELINE-size batches: the whole batch runs on the GPU, a seeded sample of cells is
    re-run by the oracle (cells are independent, so any subset must agree), and the size-
    independent property holds everywhere: free-ion output re-speciates to the stored totals
    and every cell reports a reference exit reason."""
    w, cells = workload_cells(name, n)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    for f, v in w.base.items():
        rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    if w.tables.nkinmnrl:
        rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    xg = cells['tran_xx'].copy()
    it_g, fl_g = rz.RTReact(xg, 3600.0)
    # a handful of "reaction front" cells drive the reference's Newton iteration itself to NaN (the oracle agrees
    # cell by cell, see the flag comparison below); everything else must leave through a reference exit
    conv = (fl_g == abi.RXN_EXIT_RESIDUAL) | (fl_g == abi.RXN_EXIT_REL_CHANGE)
    assert conv.mean() >= 0.9999 and ((fl_g[~conv] & abi.RXN_FLAG_NONFINITE) != 0).all()
    assert it_g.min() >= 1
    pm = rz.download('PRI_MOLAL')
    np.testing.assert_array_equal(pm.T, xg)                      # tran_xx out == stored pri_molal
    sample = np.sort(np.union1d(np.random.default_rng(11).choice(n, 3000, replace=False), np.where(~conv)[0][:64]))
    sub = {k: (v[sample] if v.ndim == 1 else (v[sample] if k == 'tran_xx' else v[:, sample])) for k, v in cells.items()}
    st_o = synth.host_state(w, sub)
    xo = sub['tran_xx'].copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, nthreads=8)
    assert (it_o == it_g[sample]).all() and (fl_o == fl_g[sample]).all()
    ok = conv[sample]
    assert rel_err(xg[sample][ok], xo[ok]).max() <= RTOL
    tot = rz.download('TOTAL')
    assert_state_close({'TOTAL': tot[:, sample], 'SEC_MOLAL': rz.download('SEC_MOLAL')[:, sample]},
                       st_o, fields=['TOTAL', 'SEC_MOLAL'], cells=np.where(ok)[0], what=name + ' full size', tables=w.tables)


@pytest.mark.parametrize('name,n', [('hanford300a_eq', 1_000_000), ('hpt_calcite', 2_000_000)])
def test_full_size_global_implicit_sampled_against_oracle(name, n):
    """BASELINE-size global-implicit step (config 4 and its 300A counterpart): RTUpdateAuxVars with activity update + the residual
    and Jacobian blocks of every cell on the GPU (device-resident outputs), a seeded sample of cells re-run by the oracle, and the
    size-independent properties everywhere: finite blocks, no failed cell, a positive diagonal (d accumulation_i / d m_i dominates)."""
    w, cells = workload_cells(name, n)
    t = w.tables
    nc = t.ncomp
    rx = rt.Reaction(t)
    rz = rt.Realization(rx, n)
    for f, v in w.base.items():
        rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    if t.nkinmnrl:
        rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * (cells['tran_xx'] / w.base['TOTAL'][None, :]))
    rz.RTUpdateAuxVars(xx, True)
    d_res = rz.device_alloc(n * nc * 8)
    d_jac = rz.device_alloc(n * nc * nc * 8)
    rz.RTResidualJacobianNonFlux_device(n, 1800.0, d_res, d_jac)       # raises RXN_ERR_CELL_FAILED on any non-finite cell
    sample = np.sort(np.random.default_rng(13).choice(n, 2000, replace=False))
    # the sampled rows of the device arrays (contiguous chunks around the sample would need n round trips: copy everything once)
    res = np.empty((n, nc)); jac = np.empty((n, nc * nc))
    rz.device_copy(res, d_res, res.nbytes, 1)
    rz.device_copy(jac, d_jac, jac.nbytes, 1)
    assert np.isfinite(res).all() and np.isfinite(jac).all()
    diag = jac.reshape(n, nc, nc)[:, np.arange(nc), np.arange(nc)]
    assert (diag[:, :t.naqcomp] > 0).all()                             # d(accumulation_i)/d m_i dominates every diagonal entry
    sub = {k: (v[sample] if v.ndim == 1 else (v[sample] if k == 'tran_xx' else v[:, sample])) for k, v in cells.items()}
    st_o = synth.host_state(w, sub)
    orc = Oracle(t)
    orc.update_auxvars(st_o, np.ascontiguousarray(xx[sample]), True, nthreads=8)
    a_o = orc.fixed_accum(st_o.copy(), np.ascontiguousarray(xx[sample]), nthreads=8)
    r_o, j_o = orc.residual_jacobian(st_o, 1800.0, nthreads=8)
    rs = residual_scale(st_o, t, r_o, a_o, 1800.0)
    assert (np.abs(res[sample] - r_o) / np.maximum(rs, 1e-300)).max() <= RTOL
    js = jacobian_scale(st_o, j_o, nc)
    assert (np.abs(jac[sample] - j_o) / np.maximum(js, 1e-300)).max() <= RTOL
    rz.device_free(d_res); rz.device_free(d_jac)


# ---- ReactionEquilibrateConstraint batched on the GPU -----------------------------------------------
class _GpuBackend:
    """equilibrate / update_auxvars of the KAT start-up sequence through the C ABI."""

    def __init__(self, t):
        self.t = t
        self.rx = rt.Reaction(t)

    def equilibrate(self, st, ctype, conc, cid, guess):
        rz = rt.Realization(self.rx, st.ncells)
        rz.upload_host_state(st)
        basis, it, status = rz.ReactionEquilibrateConstraint(ctype, conc, cid, guess, False, bool(self.t.initialize_with_molality))
        assert status[0] == abi.RXN_EQ_OK
        rz.download_host_state(st)
        return basis[0], int(it[0])

    def update_auxvars(self, st, xx, act):
        rz = rt.Realization(self.rx, st.ncells)
        rz.upload_host_state(st)
        rz.RTUpdateAuxVars(xx, act)
        rz.download_host_state(st)


@pytest.mark.parametrize('name', ['carbonate_unit', 'carbonate_dh', 'ca_carbonate_unit', 'ca_carbonate_dh', 'ion_exchange',
                                  'surface_complexation'])
def test_equilibrate_constraint_gpu_hits_reference_gold(name):
    """rxn_equilibrate_constraint_batch + rxn_update_auxvars_batch reproduce the reference's own regression gold files
    (regression_tests/ascem/batch/*.regression.gold): the GPU path pinned directly to reference output."""
    import kat
    w = synth.Workload(name)
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_GpuBackend(w.tables))
    t2, orc, st_o, xx_o, nit_o, cst_o = kat.initial_cell_from_fixture(w)
    assert nit == nit_o
    out = kat.outputs(t, st)
    checked = 0
    for var, vals in w.gold.items():
        if var in ('Transport', 'Material ID') or var.endswith('Site Density'):
            continue
        g = vals['1']
        assert abs(out[var] - g) <= RTOL * max(1.0, abs(g)), '%s %s: %.14e gold %.14e' % (name, var, out[var], g)
        checked += 1
    assert checked >= 4


def test_ascem_speciation_gpu_hits_reference_kat():
    """rxn_equilibrate_constraint_batch on BASELINE config 1 (22 primaries / 164 complexes): the reference's own printed
    speciation (example_problems/ascem_chemistry/pflotran.out:5811-5870), `iterations: 179` included."""
    import kat
    w = synth.Workload('ascem')
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_GpuBackend(w.tables))
    assert kat.check_speciation_kat(w, t, cst, nit) == 2 * 22 + 157


@pytest.mark.parametrize('name', ['calcite_kinetics', 'calcite_kinetics_vf', 'kd_w_mineral', 'kd_wo_mineral', 'general_reaction',
                                  'abcd_microbial', 'abcd_microbial_act_high', 'abcd_microbial_act_low'])
def test_time_stepped_gpu_hits_reference_gold(name):
    """rxn_fixed_accum_batch -> rxn_update_auxvars_batch -> rxn_residual_jacobian_blocks_batch -> block solve ->
    rxn_update_kinetic_state_batch, stepped as the reference's 1-cell global-implicit run (tests/gi_driver.py), reproduce the
    reference's time-stepped gold files (calcite-kinetics: 500 steps / 1000 Newton iterations, ...) at the reference's
    own 1e-12.  Two identical cells: cell 1 checks that nothing leaks between cells."""
    import gi_driver
    import kat
    w = synth.Workload(name)
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_GpuBackend(w.tables))
    st2 = abi.HostState(t, 2)
    for f in abi.FIELDS:
        if st2[f].shape[0]:
            st2[f][:] = st[f][:, :1]
    xx2 = np.ascontiguousarray(np.tile(xx, (2, 1)))
    rx = rt.Reaction(t)
    rz = rt.Realization(rx, 2)
    dev = gi_driver.DeviceGI(rz, st2)
    assert gi_driver.check_time_stepped_gold(w, dev, t, xx2, tol=1.0e-12) >= 1
    s = dev.state()
    for f in ('PRI_MOLAL', 'TOTAL', 'MNRL_VOLFRAC', 'MNRL_RATE', 'IMMOBILE'):
        if s[f].shape[0]:
            assert (s[f][:, 0] == s[f][:, 1]).all()


def test_radioactive_decay_gpu_closed_form():
    """RRadioactiveDecay on the GPU through the C ABI: 500 backward-Euler steps hit A_0 / (1 + k dt)^500 at 1e-12."""
    import gi_driver
    import kat
    w = synth.Workload('decay_ab')
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_GpuBackend(w.tables))
    rx = rt.Reaction(t)
    rz = rt.Realization(rx, 1)
    gi_driver.check_decay_closed_form(w, gi_driver.DeviceGI(rz, st), t, xx)


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite'])
def test_equilibrate_constraint_gpu_batch(name):
    import kat
    w = synth.Workload(name)
    t = w.tables
    ctype, conc, cid, guess, vf, area = kat.fixture_constraint(w)
    n = 500
    rng = np.random.default_rng(5)
    st = abi.HostState(t, n)
    kat.fill_scalars(st, t, 0.25)
    st['DEN_KG'][0] = t.reference_water_density * (1.0 + 0.01 * rng.standard_normal(n))
    if t.logK_mode != 0:
        st['TEMP'][0] = 25.0 + 100.0 * rng.random(n)
    st['MNRL_VOLFRAC'][:] = vf[:, None]
    st['MNRL_AREA'][:] = area[:, None]
    concs = np.tile(conc, (n, 1))
    scale = np.exp(0.05 * rng.standard_normal((n, t.naqcomp)))
    lin = np.isin(ctype, [0, 1, 2, 7, 9])
    concs[:, lin] *= scale[:, lin]
    st_o = st.copy()
    rx = rt.Reaction(t)
    rz = rt.Realization(rx, n)
    rz.upload_host_state(st)
    basis_g, it_g, status = rz.ReactionEquilibrateConstraint(ctype, concs, cid, guess, False, bool(t.initialize_with_molality))
    rz.download_host_state(st)
    assert (status == abi.RXN_EQ_OK).all()
    orc = Oracle(t)
    it_o = []
    for c in range(0, n, 7):
        b, it = orc.equilibrate(st_o, c, ctype, concs[c], cid, guess, use_prev=False)
        it_o.append(it)
        assert rel_err(basis_g[c], b).max() <= RTOL
    assert_state_close(st, st_o, cells=np.arange(0, n, 7), what=name + ' equilibrated batch', tables=t)
    # The 300A constraint (two mineral constraints + charge balance) takes 90-200 Newton iterations along a path that is
    # sensitive to the last bit (measured: the iteration count differs in a third of the cells while the converged
    # molarities agree to 2e-14); the short, well-conditioned solves must also agree in their iteration counts.
    if np.mean(it_o) < 40:
        assert (np.array(it_o) == it_g[::7]).all()


def test_react_chunked_host_path_with_l2g():
    """rxn_react_batch moves large batches in chunks (PCIe copies overlapping the kernel); with a local->ghosted map the
    chunks address the state through l2g, without one through a chunk base: both must give the single-launch answer."""
    n = 300_000
    w, cells = workload_cells('calcite', n)

    def run(l2g, pipeline, monkey_env):
        import os
        if pipeline:
            os.environ.pop('RXN_NO_PIPELINE', None)
        else:
            os.environ['RXN_NO_PIPELINE'] = '1'
        rx = rt.Reaction(w.tables)
        rz = rt.Realization(rx, n)
        for f, v in w.base.items():
            rz.broadcast(f, v)
        rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
        rz.upload('MNRL_VOLFRAC', cells['volfrac'])
        xx = cells['tran_xx'].copy() if l2g is None else np.ascontiguousarray(cells['tran_xx'][l2g])
        it, fl = rz.RTReact(xx, 3600.0, l2g=l2g)
        os.environ.pop('RXN_NO_PIPELINE', None)
        return xx, it, fl, rz.download('TOTAL')

    x0, it0, fl0, tot0 = run(None, False, None)
    x1, it1, fl1, tot1 = run(None, True, None)
    np.testing.assert_array_equal(x1, x0)
    np.testing.assert_array_equal(it1, it0)
    np.testing.assert_array_equal(tot1, tot0)
    perm = np.arange(n - 1, -1, -1, dtype=np.int32)
    x2, it2, fl2, tot2 = run(perm, True, None)
    np.testing.assert_array_equal(x2, x0[perm])
    np.testing.assert_array_equal(it2, it0[perm])
    np.testing.assert_array_equal(fl2, fl0[perm])
    np.testing.assert_array_equal(tot2, tot0)


def test_global_implicit_device_resident_entry_points():
    """rxn_update_auxvars_batch_device / rxn_residual_jacobian_blocks_batch_device (device pointers) give
    bit for bit what the host-buffer entry points give."""
    n = 20000
    w, cells = workload_cells('hanford300a_eq', n)
    st = synth.host_state(w, cells)
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
    nc = w.ncomp
    rx, rz = _gpu_state(w, st.copy())
    rz.RTUpdateAuxVars(xx, True)
    r_h, j_h = rz.RTResidualJacobianNonFlux(1800.0)
    rx2, rz2 = _gpu_state(w, st.copy())
    d_xx = rz2.device_alloc(n * nc * 8)
    d_res = rz2.device_alloc(n * nc * 8)
    d_jac = rz2.device_alloc(n * nc * nc * 8)
    rz2.device_copy(d_xx, xx, n * nc * 8, 0)
    rz2.RTUpdateAuxVars_device(d_xx, True)
    rz2.RTResidualJacobianNonFlux_device(n, 1800.0, d_res, d_jac)
    r_d = np.zeros((n, nc))
    j_d = np.zeros((n, nc * nc))
    rz2.device_copy(r_d, d_res, n * nc * 8, 1)
    rz2.device_copy(j_d, d_jac, n * nc * nc * 8, 1)
    np.testing.assert_array_equal(r_d, r_h)
    np.testing.assert_array_equal(j_d, j_h)
    np.testing.assert_array_equal(rz2.download('TOTAL'), rz.download('TOTAL'))
    for p in (d_xx, d_res, d_jac):
        rz2.device_free(p)


@pytest.mark.parametrize('name,n', [('ascem', 6000), ('hanford300a_eq', 30000), ('hanford300a_mr', 20000), ('scco2_brine', 40000),
                                    ('hanford300a_eq', 300000)])     # the last one: the chunked host-buffer call (never ordered; repeatable)
def test_react_work_order_is_bitwise_neutral(name, n):
    """Chemistries on the N = 24 shapes (ascem: damped redox cells with thousands of Newton iterations), and chemistries with at least
    8 primaries on batches below 16 generations of resident cells, are handed to the lanes (resident-lane and tensor-memory kernel)
    in the order of the previous call's iteration counts, slowest first (rxn_b200.cu: react_order, react_ordered).  Cells are
    independent, so the second (ordered) call must reproduce the first (unordered) one bit for bit - with and without a
    local-to-ghosted map."""
    w, cells = workload_cells(name, n)
    st0 = synth.host_state(w, cells)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    assert 'work order' in rz.react_kernel_info()
    out = []
    for call in range(3):
        rz.upload_host_state(st0)
        xx = cells['tran_xx'].copy()
        it, fl = rz.RTReact(xx, 3600.0, abi.RXN_DT_CONSISTENT)
        st = st0.copy()
        rz.download_host_state(st)
        out.append((xx, it.copy(), fl.copy(), st))
    if name == 'ascem':
        assert out[0][1].max() > 100                  # the tail the order was made for
    for k in (1, 2):
        np.testing.assert_array_equal(out[k][0], out[0][0])
        np.testing.assert_array_equal(out[k][1], out[0][1])
        np.testing.assert_array_equal(out[k][2], out[0][2])
        for f in ('PRI_MOLAL', 'TOTAL', 'SEC_MOLAL', 'MNRL_RATE'):
            np.testing.assert_array_equal(out[k][3][f], out[0][3][f])
    # local -> ghosted map: a third of the cells at reversed ghosted slots, twice (the second call ordered)
    l2g = np.arange(2 * n // 3 - 1, n // 3 - 1, -1, dtype=np.int32)
    res = []
    for call in range(2):
        rz.upload_host_state(st0)
        x2 = np.ascontiguousarray(cells['tran_xx'][l2g])
        it2, fl2 = rz.RTReact(x2, 3600.0, abi.RXN_DT_CONSISTENT, l2g=l2g)
        res.append((x2, it2.copy(), fl2.copy()))
    np.testing.assert_array_equal(res[1][0], res[0][0])
    np.testing.assert_array_equal(res[1][1], res[0][1])
    np.testing.assert_array_equal(res[0][1], out[0][1][l2g])
    np.testing.assert_array_equal(res[0][0], out[0][0][l2g])
