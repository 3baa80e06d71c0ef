"""Flux side of the global-implicit path on the CPU: the oracle's restatement of TFluxCoef / TFlux /
TFluxDerivative and of the RTResidualFlux / RTJacobianFlux interior loops against hand-computed connections and
conservation, and the row view of rxn_flux.h compiled for the host — the structure builder the library runs and the per-row
arithmetic the kernels of rxn_flux.cuh implement (the kernels themselves are compared in test_gpu_flux.py) — against the oracle, bit for bit.
The reference has no unit test or gold file for these routines alone: parity of this part is pinned by these properties."""
import numpy as np
import pytest

from pflotran_b200 import synth
from oracle.pyoracle import Oracle
import emulator
from common import workload_cells
from flux_common import structured_connections, random_connections, boundary_connections, source_sinks


def _state_with_totals(name, n, seed=11):
    w, cells = workload_cells(name, n)
    st = synth.host_state(w, cells)
    rng = np.random.default_rng(seed)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.2 * rng.standard_normal((n, w.ncomp))))
    Oracle(w.tables).update_auxvars(st, xx, True)     # total + dtotal of every cell (RTUpdateAuxVars)
    return w, st


def test_tfluxcoef_by_hand():
    conn = {'id_up': np.array([0, 0, 1], dtype=np.int32), 'id_dn': np.array([1, 1, 0], dtype=np.int32),
            'area': np.array([2.0, 2.0, 0.5]), 'velocity': np.array([3.0e-6, -3.0e-6, 0.0]),
            'disp': np.array([[1.0e-7, 2.0e-7], [1.0e-7, 2.0e-7], [4.0e-7, 5.0e-7]]), 'fraction_upwind': np.array([0.25, 0.25, 0.5])}
    Tu, Td = Oracle.flux_coefs(conn, 2, use_upwinding=True)
    # transport.F90:794-799, 813-814
    np.testing.assert_array_equal(Tu[0], (conn['disp'][0] + 3.0e-6) * 2.0 * 1000.0)
    np.testing.assert_array_equal(Td[0], (-conn['disp'][0]) * 2.0 * 1000.0)
    np.testing.assert_array_equal(Tu[1], conn['disp'][1] * 2.0 * 1000.0)
    np.testing.assert_array_equal(Td[1], (-conn['disp'][1] + -3.0e-6) * 2.0 * 1000.0)
    np.testing.assert_array_equal(Td[2], (-conn['disp'][2] + 0.0) * 0.5 * 1000.0)
    Tu, Td = Oracle.flux_coefs(conn, 2, use_upwinding=False)
    # transport.F90:804-807
    np.testing.assert_array_equal(Tu[0], (conn['disp'][0] + (1.0 - 0.25) * 3.0e-6) * 2.0 * 1000.0)
    np.testing.assert_array_equal(Td[0], (-conn['disp'][0] + 0.25 * 3.0e-6) * 2.0 * 1000.0)


def test_two_cell_flux_by_hand():
    """One connection 0 -> 1: Res = T_up total_0 + T_dn total_1 goes +Res to cell 0 and -Res to cell 1; the four Jacobian blocks
    are +Jup, +Jdn on row 0 and -Jdn (diagonal), -Jup on row 1."""
    w, st = _state_with_totals('calcite', 2)
    n = w.tables.naqcomp
    conn = {'id_up': np.array([0], dtype=np.int32), 'id_dn': np.array([1], dtype=np.int32), 'area': np.array([1.5]),
            'velocity': np.array([2.0e-6]), 'disp': np.full((1, n), 3.0e-8), 'fraction_upwind': np.array([0.5])}
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(conn, n)
    r = o.flux_residual(st, conn, Tu, Td, 2)
    res = Tu[0] * st['TOTAL'][:, 0] + Td[0] * st['TOTAL'][:, 1]
    np.testing.assert_array_equal(r[0], res)
    np.testing.assert_array_equal(r[1], -res)
    row_ptr, col, val = o.flux_jacobian(st, conn, Tu, Td, 2)
    np.testing.assert_array_equal(row_ptr, [0, 2, 4])
    np.testing.assert_array_equal(col, [0, 1, 1, 0])
    D0 = st['DTOTAL'][:, 0].reshape(n, n)   # [j, i]
    D1 = st['DTOTAL'][:, 1].reshape(n, n)
    Jup = (D0 * Tu[0][None, :]).ravel()
    Jdn = (D1 * Td[0][None, :]).ravel()
    np.testing.assert_array_equal(val[0], Jup)
    np.testing.assert_array_equal(val[1], Jdn)
    np.testing.assert_array_equal(val[2], -Jdn)
    np.testing.assert_array_equal(val[3], -Jup)


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq'])
def test_flux_conservation_and_jacobian_consistency(name):
    """All cells local: the interior fluxes cancel in the sum over cells (to rounding), and the Jacobian is the derivative of
    the residual: J . d(free ion) = d(residual) to first order."""
    nx, ny, nz = 5, 4, 3
    w, st = _state_with_totals(name, nx * ny * nz)
    n = w.tables.naqcomp
    conn, nghosted, nlocal, active = structured_connections(nx, ny, nz, n)
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(conn, n)
    r = o.flux_residual(st, conn, Tu, Td, nlocal)
    scale = np.abs(Tu).max() * np.abs(st['TOTAL']).max(axis=1)
    assert (np.abs(r.sum(axis=0)) <= 1e-12 * scale * len(conn['id_up'])).all()
    row_ptr, col, val = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    rng = np.random.default_rng(5)
    m0 = st['PRI_MOLAL'].T.copy()
    dm = 1e-7 * m0 * rng.standard_normal(m0.shape)
    st2 = st.copy()
    o.update_auxvars(st2, np.ascontiguousarray(m0 + dm), False)
    r2 = o.flux_residual(st2, conn, Tu, Td, nlocal)
    jd = np.zeros_like(r)
    for row in range(nlocal):
        for s in range(row_ptr[row], row_ptr[row + 1]):
            jd[row] += val[s].reshape(n, n).T @ dm[col[s]]
    err = np.abs((r2 - r) - jd).max(axis=0)
    assert (err <= 1e-5 * np.abs(jd).max(axis=0) + 1e-30).all()


@pytest.mark.parametrize('name,ghost,inactive,upwind', [('calcite', 0, 0.0, True), ('calcite', 1, 0.1, False),
                                                        ('hanford300a_eq', 1, 0.05, True), ('hanford300a_eq', 0, 0.0, False)])
def test_row_view_matches_connection_loop(name, ghost, inactive, upwind):
    """rxn_flux.h (structure builder + per-row sums, the arithmetic the kernels implement) against the oracle's connection loop: same
    block-CSR structure, residual and Jacobian identical bit for bit (the row view adds in the reference's order)."""
    nx, ny, nz = 7, 5, 4
    g = ghost
    nghost_cells = (nx + 2 * g) * (ny + 2 * g) * (nz + 2 * g)
    w, st = _state_with_totals(name, nghost_cells)
    n = w.tables.naqcomp
    conn, nghosted, nlocal, active = structured_connections(nx, ny, nz, n, ghost_layers=g, inactive_fraction=inactive)
    assert nghosted == nghost_cells
    st.active = active
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(conn, n, use_upwinding=upwind)
    r_o = o.flux_residual(st, conn, Tu, Td, nlocal)
    rp_o, col_o, val_o = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    rp_e, col_e, r_e, val_e = emulator.flux(st, conn, nlocal, use_upwinding=upwind)
    np.testing.assert_array_equal(rp_e, rp_o)
    np.testing.assert_array_equal(col_e, col_o)
    np.testing.assert_array_equal(r_e, r_o)
    np.testing.assert_array_equal(val_e, val_o)
    assert np.abs(r_o).max() > 0 and np.abs(val_o).max() > 0
    # rows of inactive cells take no flux
    if inactive > 0:
        l2g = np.where(conn['g2l'] >= 0)[0] if conn['g2l'] is not None else np.arange(nlocal)
        dead = np.where(active[l2g] == 0)[0]
        assert len(dead) > 0 and (r_o[dead] == 0).all()


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq'])
def test_row_view_unstructured_long_rows(name):
    """Rows with many more connections than a structured grid's six (hub cells), repeated pairs and ghost cells."""
    ncells = 300
    w, st = _state_with_totals(name, ncells)
    n = w.tables.naqcomp
    conn, nghosted, nlocal, active = random_connections(ncells, 1500, n, nghost=20)
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(conn, n)
    r_o = o.flux_residual(st, conn, Tu, Td, nlocal)
    rp_o, col_o, val_o = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    rp_e, col_e, r_e, val_e = emulator.flux(st, conn, nlocal)
    assert np.diff(rp_o).max() > 50                      # the hub rows
    np.testing.assert_array_equal(rp_e, rp_o)
    np.testing.assert_array_equal(col_e, col_o)
    np.testing.assert_array_equal(r_e, r_o)
    np.testing.assert_array_equal(val_e, val_o)


def test_connection_set_rejects_bad_maps():
    w, st = _state_with_totals('calcite', 8)
    conn, _, nlocal, _ = structured_connections(2, 2, 2, w.tables.naqcomp)
    bad = dict(conn)
    bad['id_dn'] = conn['id_dn'].copy()
    bad['id_dn'][0] = 99
    with pytest.raises(ValueError):
        emulator.flux(st, bad, nlocal)


# ---- boundary conditions and source/sinks (coupler connections) ---------------------------------------------------------------

def test_boundary_and_source_sink_by_hand():
    """One boundary face and one well on a 2-cell state, against the formulas of reactive_transport.F90:2369-2382 / 3201-3217
    (TFluxCoef with fraction_upwind 0.5, TFlux, r_p -= Res, diagonal -= dtotal coef_dn) and :2646-2661 / 3425-3430
    (TSrcSinkCoef, Res = coef_in total + coef_out total_ss, r_p += Res, diagonal += coef_in dtotal)."""
    w, st = _state_with_totals('calcite', 2)
    n = w.tables.naqcomp
    o = Oracle(w.tables)
    ext = np.ascontiguousarray(st['TOTAL'][:, :1].T * 1.3)                      # boundary auxvar totals
    bc = {'id_up': np.array([0], dtype=np.int32), 'id_dn': np.array([1], dtype=np.int32), 'area': np.array([1.5]),
          'velocity': np.array([-2.0e-6]), 'disp': np.full((1, n), 3.0e-8), 'fraction_upwind': np.array([0.5])}
    cu, cd = Oracle.flux_coefs(bc, n, use_upwinding=True)
    np.testing.assert_array_equal(cu[0], bc['disp'][0] * 1.5 * 1000.0)          # q <= 0 branch, transport.F90:797-799
    np.testing.assert_array_equal(cd[0], (-bc['disp'][0] + -2.0e-6) * 1.5 * 1000.0)
    r = np.full((2, n), 0.25)
    flux = o.coupler_residual(st, 0, bc['id_dn'], ext, cu, cd, 2, r, want_flux=True)
    res = cu[0] * ext[0] + cd[0] * st['TOTAL'][:, 1]
    np.testing.assert_array_equal(r[1], 0.25 - res)
    np.testing.assert_array_equal(r[0], 0.25)
    np.testing.assert_array_equal(flux[0], -res)
    diag = np.zeros((2, n * n))
    o.coupler_jacobian(st, 0, bc['id_dn'], cd, 2, diag)
    D1 = st['DTOTAL'][:, 1].reshape(n, n)                                       # [j, i]
    np.testing.assert_array_equal(diag[1], -(D1 * cd[0][None, :]).ravel())
    assert (diag[0] == 0).all()
    # source/sink: extraction (qsrc < 0) takes the cell's total, injection brings the constraint's
    tin, tout = Oracle.ss_coefs(np.array([-4.0e-5, 4.0e-5, 1.0, 1.0]), np.array([0, 0, 12, 7], dtype=np.int32))
    np.testing.assert_array_equal(tin, [4.0e-5 * 1000.0, 0.0, 1.0e-3, 0.0])
    np.testing.assert_array_equal(tout, [0.0, -4.0e-5 * 1000.0, -1.0e-3, -1.0])
    ids = np.array([0, 0], dtype=np.int32)
    c_in = np.repeat(tin[:2, None], n, axis=1)
    c_out = np.repeat(tout[:2, None], n, axis=1)
    ext2 = np.ascontiguousarray(np.tile(ext, (2, 1)))
    r = np.zeros((2, n))
    o.coupler_residual(st, 1, ids, ext2, c_out, c_in, 2, r)
    np.testing.assert_array_equal(r[0], (0.0 + tin[0] * st['TOTAL'][:, 0]) + (tout[1] * ext[0]))
    diag = np.zeros((2, n * n))
    o.coupler_jacobian(st, 1, ids, c_in, 2, diag)
    np.testing.assert_array_equal(diag[0], (tin[0] * st['DTOTAL'][:, 0]))


@pytest.mark.parametrize('name,ghost,inactive,upwind', [('calcite', 0, 0.0, True), ('calcite', 1, 0.1, False), ('hanford300a_eq', 1, 0.05, True)])
def test_coupler_row_view_matches_connection_loop(name, ghost, inactive, upwind):
    """rxn_flux.h's coupler row view (what k_coupler_residual / k_coupler_jacobian implement) against the oracle's boundary and
    source/sink loops, applied on top of the interior-flux result: bit for bit, corner cells with three faces and two wells in
    one cell included."""
    nx, ny, nz = 6, 5, 4
    g = ghost
    nghosted = (nx + 2 * g) * (ny + 2 * g) * (nz + 2 * g)
    w, st = _state_with_totals(name, nghosted)
    n = w.tables.naqcomp
    conn, _, nlocal, active = structured_connections(nx, ny, nz, n, ghost_layers=g, inactive_fraction=inactive)
    st.active = active
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(conn, n, use_upwinding=upwind)
    r0 = o.flux_residual(st, conn, Tu, Td, nlocal)
    rp, col, val = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    bc = boundary_connections(nx, ny, nz, n, ghost_layers=g)
    nb = len(bc['id_dn'])
    rng = np.random.default_rng(31)
    ext = np.ascontiguousarray(st['TOTAL'][:, bc['id_dn']].T * np.exp(0.3 * rng.standard_normal((nb, n))))
    cu, cd = Oracle.flux_coefs({**bc, 'fraction_upwind': np.full(nb, 0.5)}, n, use_upwinding=upwind)
    r_o = r0.copy()
    f_o = o.coupler_residual(st, 0, bc['id_dn'], ext, cu, cd, nlocal, r_o, g2l=conn['g2l'], want_flux=True)
    d_o = np.ascontiguousarray(val[rp[:-1]])
    o.coupler_jacobian(st, 0, bc['id_dn'], cd, nlocal, d_o, g2l=conn['g2l'])
    r_e = r0.copy()
    d_e = np.ascontiguousarray(val[rp[:-1]])
    f_e = emulator.coupler(st, 0, bc['id_dn'], nlocal, ext, g2l=conn['g2l'], area=bc['area'], velocity=bc['velocity'], disp=bc['disp'],
                           use_upwinding=upwind, res=r_e, diag=d_e, want_flux=True)
    np.testing.assert_array_equal(r_e, r_o)
    np.testing.assert_array_equal(f_e, f_o)
    np.testing.assert_array_equal(d_e, d_o)
    assert (r_o != r0).any() and np.abs(f_o).max() > 0
    # source/sinks on top
    local_g = np.where(conn['g2l'] >= 0)[0] if conn['g2l'] is not None else np.arange(nlocal)
    ss = source_sinks(local_g, n)
    tin, tout = Oracle.ss_coefs(ss['qsrc'], ss['ss_type'])
    ext_s = np.ascontiguousarray(np.tile(w.base['TOTAL'] * 0.7, (len(tin), 1)))
    o.coupler_residual(st, 1, ss['id_dn'], ext_s, np.repeat(tout[:, None], n, 1).copy(), np.repeat(tin[:, None], n, 1).copy(), nlocal, r_o, g2l=conn['g2l'])
    o.coupler_jacobian(st, 1, ss['id_dn'], np.repeat(tin[:, None], n, 1).copy(), nlocal, d_o, g2l=conn['g2l'])
    emulator.coupler(st, 1, ss['id_dn'], nlocal, ext_s, g2l=conn['g2l'], qsrc=ss['qsrc'], ss_type=ss['ss_type'], res=r_e, diag=d_e)
    np.testing.assert_array_equal(r_e, r_o)
    np.testing.assert_array_equal(d_e, d_o)


def test_boundary_closes_the_mass_balance():
    """With every boundary face a coupler connection, the interior fluxes cancel and what is left of the summed residual is
    exactly the boundary fluxes: sum_cells r = sum_faces boundary_tran_fluxes (to rounding)."""
    nx, ny, nz = 5, 4, 3
    w, st = _state_with_totals('calcite', nx * ny * nz)
    n = w.tables.naqcomp
    conn, _, nlocal, _ = structured_connections(nx, ny, nz, n)
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(conn, n)
    r = o.flux_residual(st, conn, Tu, Td, nlocal)
    bc = boundary_connections(nx, ny, nz, n)
    nb = len(bc['id_dn'])
    ext = np.ascontiguousarray(st['TOTAL'][:, bc['id_dn']].T * 1.1)
    cu, cd = Oracle.flux_coefs({**bc, 'fraction_upwind': np.full(nb, 0.5)}, n)
    flux = o.coupler_residual(st, 0, bc['id_dn'], ext, cu, cd, nlocal, r, want_flux=True)
    scale = np.abs(cu).max() * np.abs(st['TOTAL']).max(axis=1) * (nb + len(conn['id_up']))
    assert (np.abs(r.sum(axis=0) - flux.sum(axis=0)) <= 1e-12 * scale).all()


def test_coupler_set_rejects_ghost_cells():
    w, st = _state_with_totals('calcite', 4 * 4 * 4)
    n = w.tables.naqcomp
    conn, _, nlocal, _ = structured_connections(2, 2, 2, n, ghost_layers=1)
    ghost = np.where(conn['g2l'] < 0)[0][:1].astype(np.int32)
    with pytest.raises(ValueError):
        emulator.coupler(st, 1, ghost, nlocal, np.zeros((1, n)), g2l=conn['g2l'], qsrc=np.zeros(1), ss_type=np.zeros(1, dtype=np.int32),
                         res=np.zeros((nlocal, n)))
