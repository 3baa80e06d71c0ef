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
from flux_common import structured_connections, random_connections


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
