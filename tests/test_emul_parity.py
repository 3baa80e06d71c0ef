"""CPU-only parity of the CUDA device routines (compiled for the host by tests/emul) and of
the table packer against the oracle, on the seeded synthetic workloads.  This is the logic
check that runs in the build container; the same comparisons run against the real CUDA
library in test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from pflotran_b200 import abi, synth
from oracle.pyoracle import Oracle
from emulator import Emulator, pack_status
from common import PerturbedOracle, iteration_parity, free_ion_parity, assert_state_close, workload_cells, RTOL, rel_err, total_magnitude, accumulation_scale, residual_scale, jacobian_scale

WORKLOADS = ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite', 'ion_exchange', 'surface_complexation',
             'calcite_kinetics', 'kd_wo_mineral']
# fixtures that reach the branches no reference batch deck exercises (tests/golden/make_fixtures.py: VARIANTS and the
# prefactor / non-isothermal / 22-primary decks): NEWTON activity algorithm + activity of water, free-site inner Newton,
# Langmuir / Freundlich isotherms, Temkin / scale factor / affinity power / threshold / rate limiter / Arrhenius, mineral
# prefactors, 5-term logK fit per cell, BASELINE config 1 (22 primaries / 164 complexes), general (forward / backward rate)
# reactions, radioactive decay, kinetic surface complexation, microbial reactions (Monod / inverse-Monod terms, biomass as an
# immobile dof, Arrhenius factor) with immobile decay, and a microbial reaction without biomass in the linear formulation
BRANCH_WORKLOADS = ['hanford300a_act_newton', 'hanford300a_stoich', 'kd_langmuir', 'kd_freundlich', 'calcite_rate_laws', 'mineral_prefactor', 'calcite_fit5', 'ascem', 'general_reaction', 'decay_ab', 'hanford300a_kinsrf',
                    'abcd_microbial', 'abcd_microbial_act_high', 'ab_microbial_linear', 'scco2_brine', 'abcd_microbial_inhibition']
WORKLOADS = WORKLOADS + BRANCH_WORKLOADS
GI_WORKLOADS = ['calcite', 'hanford300a_mr', 'hpt_calcite', 'ion_exchange', 'surface_complexation'] + BRANCH_WORKLOADS


@pytest.mark.parametrize('name', WORKLOADS)
@pytest.mark.parametrize('dt,mode', [(3600.0, abi.RXN_DT_CONSISTENT), (1.0, abi.RXN_DT_AS_WRITTEN)])
def test_react(name, dt, mode):
    w, cells = workload_cells(name, 600)
    st_o = synth.host_state(w, cells)
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, dt, mode, maxit=10000)
    it_e, fl_e = Emulator(w.tables).react(st_e, xe, dt, mode)
    pert = PerturbedOracle(w, cells, dt, mode, nthreads=1)
    same = iteration_parity(it_e, fl_e, it_o, fl_o, pert)
    ok = ((fl_o & ~3) == 0) & same
    good = free_ion_parity(xe, xo, ok, pert)
    assert_state_close(st_e, st_o, cells=good, what=name)


LANE_WORKLOADS = ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite', 'surface_complexation', 'calcite_kinetics', 'scco2_brine']


@pytest.mark.parametrize('name', LANE_WORKLOADS)
@pytest.mark.parametrize('dt,mode', [(3600.0, abi.RXN_DT_CONSISTENT), (1.0, abi.RXN_DT_AS_WRITTEN)])
@pytest.mark.parametrize('G', [1, 2, 4])
def test_react_resident_lane(name, dt, mode, G):
    """Per-lane routines of the resident-lane kernel (rxn_lane_dev.cuh: term streams, ln-m Jacobian, LU in the
    cell-fastest layout, closing pass) with G lanes per cell (host threads meeting at barriers where the device
    has __syncwarp / shuffles), driven cell by cell on the host, against the oracle."""
    w, cells = workload_cells(name, 300 if G > 1 else 600)
    st_o = synth.host_state(w, cells)
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, dt, mode, maxit=10000)
    it_e, fl_e = Emulator(w.tables).react_lane(st_e, xe, dt, mode, G=G)
    assert (it_o == it_e).all() and (fl_o == fl_e).all()
    ok = (fl_o & ~3) == 0
    assert rel_err(xe[ok], xo[ok]).max() <= RTOL
    assert_state_close(st_e, st_o, cells=np.where(ok)[0], what=name, tables=w.tables)


@pytest.mark.parametrize('name', LANE_WORKLOADS)
@pytest.mark.parametrize('dt,mode', [(3600.0, abi.RXN_DT_CONSISTENT), (1.0, abi.RXN_DT_AS_WRITTEN)])
@pytest.mark.parametrize('G', [1, 2, 3, 4])
def test_react_tensor_memory_routines(name, dt, mode, G):
    """Routines of the tensor-memory kernel (rxn_tm_dev.cuh: J rows in the emulated TMEM lane, k-unrolled LU with select
    swaps, predicated trips, exchange-slot reductions) with G member warps per cell (host threads, one lane each),
    against the oracle.  Chemistries with fewer than 13 primaries run on the padded N = 12 shape."""
    w, cells = workload_cells(name, 300 if G > 1 else 600)
    st_o = synth.host_state(w, cells)
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, dt, mode, maxit=10000)
    it_e, fl_e = Emulator(w.tables).react_tm(st_e, xe, dt, mode, G=G)
    assert (it_o == it_e).all() and (fl_o == fl_e).all()
    ok = (fl_o & ~3) == 0
    assert rel_err(xe[ok], xo[ok]).max() <= RTOL
    assert_state_close(st_e, st_o, cells=np.where(ok)[0], what=name, tables=w.tables)


@pytest.mark.parametrize('G', [1, 2, 3, 4])
def test_tensor_memory_iteration_cap_and_inactive(G):
    """Abnormal exit (iteration cap) in the predicated trip: the lane turns `closing`, redoes RTotal, then finishes."""
    w, cells = workload_cells('hanford300a_eq', 200)
    st_o = synth.host_state(w, cells)
    st_o.active[::7] = 0
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=3)
    it_e, fl_e = Emulator(w.tables).react_tm(st_e, xe, 3600.0, abi.RXN_DT_CONSISTENT, maxit=3, G=G)
    assert (fl_e & abi.RXN_FLAG_CAPPED).any() and (fl_e[::7] == abi.RXN_FLAG_INACTIVE).all()
    assert (it_o == it_e).all() and (fl_o == fl_e).all()
    act = np.where(st_o.active != 0)[0]
    assert rel_err(xe[act], xo[act]).max() <= RTOL
    assert_state_close(st_e, st_o, cells=act, what='capped', tables=w.tables)


@pytest.mark.parametrize('G', [1, 4])
def test_resident_lane_iteration_cap_and_inactive(G):
    """Abnormal exit (iteration cap): pri_molal moved after the last RTotal, so the closing pass must redo it."""
    w, cells = workload_cells('hanford300a_eq', 200)
    st_o = synth.host_state(w, cells)
    st_o.active[::7] = 0
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=3)
    it_e, fl_e = Emulator(w.tables).react_lane(st_e, xe, 3600.0, abi.RXN_DT_CONSISTENT, maxit=3, G=G)
    assert (fl_e & abi.RXN_FLAG_CAPPED).any() and (fl_e[::7] == abi.RXN_FLAG_INACTIVE).all()
    assert (it_o == it_e).all() and (fl_o == fl_e).all()
    act = np.where(st_o.active != 0)[0]
    assert rel_err(xe[act], xo[act]).max() <= RTOL
    assert_state_close(st_e, st_o, cells=act, what='capped', tables=w.tables)


@pytest.mark.parametrize('name,N,G', [('hanford300a_eq', 16, 2), ('hanford300a_eq', 24, 4), ('hanford300a_eq', 24, 8), ('calcite', 8, 1), ('calcite', 12, 2)])
def test_resident_lane_padded_shapes(name, N, G):
    """naq smaller than the compiled matrix dimension: the padding rows (m = 1, zero residual, decoupled) must not change
    anything - iteration counts, flags and values as with the exact shape."""
    w, cells = workload_cells(name, 200)
    st_o = synth.host_state(w, cells)
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000)
    em = Emulator(w.tables)
    it_e, fl_e = em.react_lane(st_e, xe, 3600.0, abi.RXN_DT_CONSISTENT, G=G, N=N)
    assert em.lane_stats['N'] == N
    assert (it_o == it_e).all() and (fl_o == fl_e).all()
    ok = (fl_o & ~3) == 0
    assert rel_err(xe[ok], xo[ok]).max() <= RTOL
    assert_state_close(st_e, st_o, cells=np.where(ok)[0], what=name, tables=w.tables)


def test_resident_lane_rejects_what_it_does_not_cover():
    for name in ['ion_exchange', 'kd_wo_mineral']:
        w, cells = workload_cells(name, 8)
        st = synth.host_state(w, cells)
        with pytest.raises(NotImplementedError):
            Emulator(w.tables).react_lane(st, cells['tran_xx'].copy(), 3600.0)


@pytest.mark.parametrize('name', GI_WORKLOADS)
def test_global_implicit_entry_points(name):
    w, cells = workload_cells(name, 300)
    st_o = synth.host_state(w, cells)
    orc, emu = Oracle(w.tables), Emulator(w.tables)
    # a perturbed iterate: free-ion molalities around the base state
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base_solution()[None, :] * np.exp(0.1 * rng.standard_normal((300, w.ncomp))))
    st_e = st_o.copy()
    orc.update_auxvars(st_o, xx, True)
    emu.update_auxvars(st_e, xx, True)
    assert_state_close(st_e, st_o, what=name + ' update_auxvars')
    a_o = orc.fixed_accum(st_o, xx)
    a_e = emu.fixed_accum(st_e, xx)
    # accumulation = phi*s*1000*V*total (+ sorbed*V), reaction.F90:5072-5148: compared on the scale of total's terms
    a_scale = accumulation_scale(st_o, w.tables, a_o)
    assert (np.abs(a_e - a_o) / np.maximum(a_scale, 1e-300)).max() <= RTOL
    r_o, j_o = orc.residual_jacobian(st_o, 1800.0)
    r_e, j_e = emu.residual_jacobian(st_e, 1800.0)
    n = w.ncomp
    rs = residual_scale(st_o, w.tables, r_o, a_o, 1800.0)
    assert (np.abs(r_e - r_o) / np.maximum(rs, 1e-300)).max() <= RTOL
    js = jacobian_scale(st_o, j_o, w.ncomp)
    assert (np.abs(j_e - j_o) / np.maximum(js, 1e-300)).max() <= RTOL
    orc.update_kinetic_state(st_o, 1800.0)
    emu.update_kinetic_state(st_e, 1800.0)
    assert_state_close(st_e, st_o, what=name + ' kinetic state', tables=w.tables, kinetic_dt=1800.0)


@pytest.mark.parametrize('name,G', [(n, g) for n in ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite', 'surface_complexation']
                                    for g in [1, 2, 4]] + [('ascem', 8)])     # ascem: the library's N = 24, 8-lanes-per-cell shape
def test_global_implicit_blocks_resident_lane(name, G):
    """Residual / Jacobian blocks through the resident-lane routines (lane_gi_cell: ln-m Jacobian divided by m_j on the
    way out, activity coefficients taken from the state) against the oracle, including the state side effects."""
    w, cells = workload_cells(name, 200)
    st_o = synth.host_state(w, cells)
    orc, emu = Oracle(w.tables), Emulator(w.tables)
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((200, w.ncomp))))
    orc.update_auxvars(st_o, xx, True)
    st_e = st_o.copy()
    a_o = orc.fixed_accum(st_o, xx)
    emu.fixed_accum(st_e, xx)
    r_o, j_o = orc.residual_jacobian(st_o, 1800.0)
    r_e, j_e = emu.residual_jacobian_lane(st_e, 1800.0, G=G)
    rs = np.maximum(np.maximum(np.abs(r_o), np.abs(a_o) / 1800.0), 1e-12 * np.abs(r_o).max(axis=1, keepdims=True))
    assert (np.abs(r_e - r_o) / np.maximum(rs, 1e-300)).max() <= RTOL
    js = jacobian_scale(st_o, j_o, w.ncomp)
    assert (np.abs(j_e - j_o) / np.maximum(js, 1e-300)).max() <= RTOL
    assert_state_close(st_e, st_o, what=name + ' residual/Jacobian state', tables=w.tables)


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite', 'surface_complexation', 'hanford300a_stoich',
                                  'calcite_rate_laws', 'calcite_fit5'])
@pytest.mark.parametrize('G', [1, 2, 3, 4])
def test_global_implicit_tensor_memory_routines(name, G):
    """The global-implicit cell loops on the tensor-memory layout (tm_gi_cell: one pass per cell, class-based Debye-Hueckel
    update or the state's per-species activity coefficients, dtotal / dtotal_sorb_eq blocks from the ln-m Jacobian in the
    cell's TMEM lane) against the oracle: RTUpdateAuxVars with and without activity update, RTUpdateFixedAccumulation, residual
    and Jacobian blocks, including every state side effect."""
    nc = 150
    w, cells = workload_cells(name, nc)
    st_o = synth.host_state(w, cells)
    orc, emu = Oracle(w.tables), Emulator(w.tables)
    rng = np.random.default_rng(7)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((nc, w.ncomp))))
    st_e = st_o.copy()
    orc.update_auxvars(st_o, xx, True)
    emu.gi_tm(st_e, 1, G=G, update_act=True, xx=xx)
    assert_state_close(st_e, st_o, what=name + ' update_auxvars(act)', tables=w.tables)
    for f in ('DTOTAL', 'DTOTAL_SORB_EQ'):
        if st_o[f].size:
            sc = np.maximum(np.abs(st_o[f]), 1e-12 * np.abs(st_o[f]).max(axis=0, keepdims=True))
            assert (np.abs(st_e[f] - st_o[f]) / np.maximum(sc, 1e-300)).max() <= RTOL, f
    # a second iterate with the activity coefficients of the state (lagged), then the accumulation from a third
    xx2 = np.ascontiguousarray(xx * np.exp(0.05 * rng.standard_normal(xx.shape)))
    orc.update_auxvars(st_o, xx2, False)
    emu.gi_tm(st_e, 1, G=G, update_act=False, xx=xx2)
    assert_state_close(st_e, st_o, what=name + ' update_auxvars(lagged)', tables=w.tables)
    a_o = orc.fixed_accum(st_o, xx)
    a_e = emu.gi_tm(st_e, 1, G=G, xx=xx, xx_by_item=True, want_accum=True)
    a_scale = accumulation_scale(st_o, w.tables, a_o)
    assert (np.abs(a_e - a_o) / np.maximum(a_scale, 1e-300)).max() <= RTOL
    assert_state_close(st_e, st_o, what=name + ' fixed accumulation state', tables=w.tables)
    r_o, j_o = orc.residual_jacobian(st_o, 1800.0)
    r_e, j_e = emu.gi_tm(st_e, 2, G=G, dt=1800.0)
    rs = residual_scale(st_o, w.tables, r_o, a_o, 1800.0)
    assert (np.abs(r_e - r_o) / np.maximum(rs, 1e-300)).max() <= RTOL
    js = jacobian_scale(st_o, j_o, w.ncomp)
    assert (np.abs(j_e - j_o) / np.maximum(js, 1e-300)).max() <= RTOL
    assert_state_close(st_e, st_o, what=name + ' residual/Jacobian state', tables=w.tables)


def test_global_implicit_tensor_memory_inactive_and_l2g():
    w, cells = workload_cells('hanford300a_eq', 40)
    st_o = synth.host_state(w, cells)
    st_o.active = np.ones(40, dtype=np.uint8); st_o.active[[3, 17]] = 0
    orc, emu = Oracle(w.tables), Emulator(w.tables)
    xx = np.ascontiguousarray(np.tile(w.base['PRI_MOLAL'] * 1.05, (40, 1)))
    st_e = st_o.copy()
    orc.update_auxvars(st_o, xx, True)
    emu.gi_tm(st_e, 1, G=3, update_act=True, xx=xx)
    assert_state_close(st_e, st_o, what='inactive', tables=w.tables)
    assert (st_e['PRI_MOLAL'][:, 3] == w.base['PRI_MOLAL']).all()          # untouched
    l2g = np.array([5, 2, 17, 30], dtype=np.int32)
    r_e, j_e = emu.gi_tm(st_e, 2, G=3, dt=900.0, l2g=l2g)
    r_o, j_o = orc.residual_jacobian(st_o, 900.0)
    live = [0, 1, 3]
    assert (np.abs(r_e[live] - r_o[l2g[live]]) <= 1e-9 * np.abs(r_o[l2g[live]]).max()).all()
    assert (r_e[2] == 0).all() and (j_e[2] == 0).all()                      # inactive cell: block left as the caller zeroed it


def test_inactive_cells_and_l2g():
    w, cells = workload_cells('calcite', 64)
    st_o = synth.host_state(w, cells)
    st_o.active[::5] = 0
    st_e = st_o.copy()
    xo = cells['tran_xx'].copy()
    xe = xo.copy()
    it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0)
    it_e, fl_e = Emulator(w.tables).react(st_e, xe, 3600.0)
    assert (fl_e[::5] == abi.RXN_FLAG_INACTIVE).all() and (it_e[::5] == 0).all()
    assert (it_o == it_e).all() and (fl_o == fl_e).all()
    np.testing.assert_array_equal(xe[::5], cells['tran_xx'][::5])     # skipped cells untouched
    # local->ghosted map: 16 local cells living at ghosted slots 40..55 in reverse order
    st2 = synth.host_state(w, cells)
    l2g = np.arange(55, 39, -1, dtype=np.int32)
    x2 = np.ascontiguousarray(cells['tran_xx'][l2g])
    it2, fl2 = Emulator(w.tables).react(st2, x2, 3600.0, l2g=l2g)
    st_ref = synth.host_state(w, cells)
    xr = cells['tran_xx'].copy()
    itr, flr = Oracle(w.tables).react(st_ref, xr, 3600.0)
    assert (it2 == itr[l2g]).all()
    assert rel_err(x2, xr[l2g]).max() <= RTOL


def test_packer_rejects_unsupported():
    w = synth.Workload('calcite')
    d = abi.make_desc(w.tables)
    for fld in ['nactive_gas', 'ncoll', 'has_sandbox', 'has_clm', 'has_solid_solution', 'co2_flow_mode',
                'numerical_derivatives']:
        setattr(d, fld, 1)
        rc, msg = pack_status(d)
        assert rc == abi.RXN_ERR_UNSUPPORTED and 'outside the B200 path' in msg, fld
        setattr(d, fld, 0)
    # counts of the supported rate reactions without their tables, and more kinetic surface complexation reactions than the
    # reference's state holds
    for fld in ['ngeneral_rxn', 'nradiodecay_rxn', 'nkinsrfcplxrxn', 'nmicrobial_rxn', 'nimmobile_decay_rxn']:
        setattr(d, fld, 1)
        assert pack_status(d)[0] == abi.RXN_ERR_INVALID, fld
        setattr(d, fld, 0)
    d.nimmobile = 1                       # immobile dofs must be counted in ncomp
    assert pack_status(d)[0] == abi.RXN_ERR_UNSUPPORTED
    d.nimmobile = 0
    d.nkinsrfcplxrxn = 2
    assert pack_status(d)[0] == abi.RXN_ERR_UNSUPPORTED
    d.nkinsrfcplxrxn = 0
    assert pack_status(d)[0] == abi.RXN_OK
    d.struct_size = 8
    assert pack_status(d)[0] == abi.RXN_ERR_INVALID


# ---- ReactionEquilibrateConstraint on the device routines -------------------------------------------
class _EmuBackend:
    """equilibrate / update_auxvars of the KAT start-up sequence through the host compilation of the device code."""

    def __init__(self, t):
        self.emu = Emulator(t)
        self.t = t

    def equilibrate(self, st, ctype, conc, cid, guess):
        basis, it, status = self.emu.equilibrate_batch(st, ctype, conc, cid, guess, False, bool(self.t.initialize_with_molality))
        assert status[0] == 0
        return basis[0], int(it[0])

    def update_auxvars(self, st, xx, act):
        self.emu.update_auxvars(st, xx, act)


@pytest.mark.parametrize('name', ['carbonate_unit', 'carbonate_dh', 'ca_carbonate_unit', 'ca_carbonate_dh', 'ion_exchange',
                                  'surface_complexation'])
def test_equilibrate_constraint_device_code_hits_reference_gold(name):
    """cell_equilibrate (rxn_device.cuh) + RTUpdateAuxVars reproduce the reference's 14-digit gold files (free, pH,
    charge-balance, mineral and total constraints), with the oracle's iteration count."""
    import kat
    w = synth.Workload(name)
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_EmuBackend(w.tables))
    t2, orc, st_o, xx_o, nit_o, cst_o = kat.initial_cell_from_fixture(w)
    assert nit == nit_o
    out = kat.outputs(t, st)
    for var, vals in w.gold.items():
        if var in ('Transport', 'Material ID') or var.endswith('Site Density'):
            continue
        g = vals['1']
        assert abs(out[var] - g) <= 1.0e-12 * max(1.0, abs(g)), '%s %s: %.14e gold %.14e' % (name, var, out[var], g)
    assert_state_close(cst, cst_o, what=name + ' constraint cell', tables=w.tables)


def test_ascem_speciation_device_code_hits_reference_kat():
    """cell_equilibrate (N = 24 variant) on the 22 / 164 chemistry: the reference's 179 iterations and printed speciation."""
    import kat
    w = synth.Workload('ascem')
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_EmuBackend(w.tables))
    assert kat.check_speciation_kat(w, t, cst, nit) == 2 * 22 + 157


class _EmuGI:
    """global-implicit entry points of the host-compiled device routines (tests/gi_driver.py backend)"""

    def __init__(self, t, st):
        self.emu, self.st = Emulator(t), st

    def fixed_accum(self, xx):
        return self.emu.fixed_accum(self.st, xx)

    def update_auxvars(self, xx, act):
        self.emu.update_auxvars(self.st, xx, act)

    def residual_jacobian(self, dt):
        return self.emu.residual_jacobian(self.st, dt)

    def update_kinetic_state(self, dt):
        self.emu.update_kinetic_state(self.st, dt)

    def state(self):
        return self.st


@pytest.mark.parametrize('name', ['calcite_kinetics', 'calcite_kinetics_vf', 'kd_w_mineral', 'kd_wo_mineral', 'general_reaction',
                                  'abcd_microbial', 'abcd_microbial_act_high', 'abcd_microbial_act_low'])
def test_time_stepped_device_code_hits_reference_gold(name):
    """The device routines behind rxn_fixed_accum / update_auxvars / residual_jacobian_blocks / update_kinetic_state
    (host compilation) driven through the reference's 1-cell global-implicit time loop reproduce
    calcite-kinetics(.volume-fractions).regression.gold, solute_KD_{w,wo}_mineral.regression.gold and
    ABCD_microbial(_activation_{high,low}).regression.gold (microbial reaction with biomass as an immobile dof, immobile
    decay) at 1e-12 with the reference's own time-step and Newton-iteration counts."""
    import gi_driver
    import kat
    w = synth.Workload(name)
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_EmuBackend(w.tables))
    assert gi_driver.check_time_stepped_gold(w, _EmuGI(t, st), t, xx) >= 1


def test_radioactive_decay_device_code_closed_form():
    import gi_driver
    import kat
    w = synth.Workload('decay_ab')
    t, be, st, xx, nit, cst = kat.initial_cell_from_fixture(w, backend=_EmuBackend(w.tables))
    gi_driver.check_decay_closed_form(w, _EmuGI(t, st), t, xx)


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite'])
def test_equilibrate_constraint_batch_per_cell_concentrations(name):
    """A batch of cells with their own constraint concentrations, water density and temperature against the oracle."""
    import kat
    w = synth.Workload(name)
    t = w.tables
    ctype, conc, cid, guess, vf, area = kat.fixture_constraint(w)
    n = 24
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
    lin = np.isin(ctype, [0, 1, 2, 7, 9])                         # concentrations (not pH / log / mineral ids): perturb
    concs[:, lin] *= scale[:, lin]
    st_o = st.copy()
    basis_e, it_e, status = Emulator(t).equilibrate_batch(st, ctype, concs, cid, guess, False, bool(t.initialize_with_molality))
    orc = Oracle(t)
    for c in range(n):
        b, it = orc.equilibrate(st_o, c, ctype, concs[c], cid, guess, use_prev=False)
        assert it == it_e[c] and status[c] == 0
        assert rel_err(basis_e[c], b).max() <= RTOL
    assert_state_close(st, st_o, what=name + ' equilibrated batch', tables=t)


def test_packer_validates_microbial_and_immobile_tables():
    """Range checks of the microbial / immobile tables (rxn_pack.h): ids that would index past the per-cell arrays are rejected with
    RXN_ERR_INVALID, an inhibition type RMicrobial has no branch for with RXN_ERR_UNSUPPORTED, ncomp must count the immobile dofs."""
    import copy
    w = synth.Workload('abcd_microbial')
    assert pack_status(abi.make_desc(w.tables))[0] == abi.RXN_OK

    def status(mutate):
        t = copy.deepcopy(w.tables)
        mutate(t)
        return pack_status(abi.make_desc(t))[0]

    def set_(name, idx, val):
        def f(t):
            a = getattr(t, name).copy(); a[idx] = val; setattr(t, name, a)
        return f
    assert status(set_('microbial_specid', (0, 1), 9)) == abi.RXN_ERR_INVALID            # species id beyond ncomp = 4
    assert status(set_('microbial_biomassid', 0, 2)) == abi.RXN_ERR_INVALID              # one immobile species only
    assert status(set_('microbial_monod_specid', 0, 4)) == abi.RXN_ERR_INVALID           # Monod terms act on aqueous species
    assert status(set_('microbial_monodid', (0, 1), 3)) == abi.RXN_ERR_INVALID           # two Monod terms exist
    assert status(set_('microbial_inhibition_specid', 0, 0)) == abi.RXN_ERR_INVALID
    assert status(set_('microbial_inhibition_type', 0, 2)) == abi.RXN_ERR_UNSUPPORTED     # INHIBITION_THERMODYNAMIC
    assert status(set_('immobile_decayspecid', 0, 2)) == abi.RXN_ERR_INVALID
    assert status(lambda t: setattr(t, 'ncomp', 3)) == abi.RXN_ERR_UNSUPPORTED            # ncomp must be naqcomp + nimmobile
    assert status(lambda t: setattr(t, 'nimmobile', 5)) == abi.RXN_ERR_UNSUPPORTED
