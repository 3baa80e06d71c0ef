"""GPU tests of the flux side: rxn_connset_* / rxn_flux_*_batch through the C ABI against the oracle's
restatement of the RTResidualFlux / RTJacobianFlux interior loops.  With the same total / dtotal in the state the kernels add
in the reference's order, so the comparison is bit for bit."""
import numpy as np
import pytest

from pflotran_b200 import synth, reactive_transport as rt
from oracle.pyoracle import Oracle
from common import workload_cells, RTOL
from flux_common import structured_connections, random_connections, boundary_connections, source_sinks
from pflotran_b200 import abi

pytestmark = pytest.mark.gpu


def _setup(name, nx, ny, nz, ghost, inactive, seed=11):
    g = ghost
    ncell = (nx + 2 * g) * (ny + 2 * g) * (nz + 2 * g)
    w, cells = workload_cells(name, ncell)
    st = synth.host_state(w, cells)
    rng = np.random.default_rng(seed)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.2 * rng.standard_normal((ncell, w.ncomp))))
    conn, nghosted, nlocal, active = structured_connections(nx, ny, nz, w.tables.naqcomp, ghost_layers=g, inactive_fraction=inactive)
    st.active = active
    return w, st, xx, conn, nlocal


@pytest.mark.parametrize('name,dims,ghost,inactive,upwind', [
    ('calcite', (37, 9, 5), 0, 0.0, True), ('calcite', (33, 7, 3), 1, 0.1, False),
    ('hanford300a_eq', (19, 6, 5), 1, 0.05, True), ('hanford300a_eq', (65, 3, 2), 0, 0.0, False),
    ('hanford300a_mr', (8, 4, 3), 0, 0.0, True), ('hpt_calcite', (31, 1, 1), 0, 0.0, True)])
@pytest.mark.parametrize('generic', [0, 1])
def test_flux_residual_and_jacobian_bitwise(name, dims, ghost, inactive, upwind, generic, monkeypatch):
    """generic = 1: the run-time-n Jacobian kernel (any chemistry) instead of the compile-time-n instantiation."""
    monkeypatch.setenv('RXN_FLUX_GENERIC', str(generic))
    w, st, xx, conn, nlocal = _setup(name, *dims, ghost, inactive)
    o = Oracle(w.tables)
    o.update_auxvars(st, xx, True)
    n = w.tables.naqcomp
    Tu, Td = o.flux_coefs(conn, n, use_upwinding=upwind)
    r_o = o.flux_residual(st, conn, Tu, Td, nlocal)
    rp_o, col_o, val_o = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, st.ncells)
    rz.upload_host_state(st)
    rz.upload('DTOTAL', st['DTOTAL'])
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal, conn['g2l'], st.active)
    rp, col = cs.structure()
    np.testing.assert_array_equal(rp, rp_o)
    np.testing.assert_array_equal(col, col_o)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'], conn['fraction_upwind'], upwind)
    r_g = rz.RTResidualFlux(cs)
    val_g = rz.RTJacobianFlux(cs)
    np.testing.assert_array_equal(r_g, r_o)
    np.testing.assert_array_equal(val_g, val_o)
    # device-resident outputs (PETSc VECCUDA / MATSEQBAIJ value array on the GPU)
    d_r = rz.device_alloc(r_o.nbytes)
    d_v = rz.device_alloc(val_o.nbytes)
    rz.RTResidualFlux_device(cs, d_r)
    rz.RTJacobianFlux_device(cs, d_v)
    r_d = np.zeros_like(r_o); v_d = np.zeros_like(val_o)
    rz.device_copy(r_d, d_r, r_o.nbytes, 1)
    rz.device_copy(v_d, d_v, val_o.nbytes, 1)
    np.testing.assert_array_equal(r_d, r_o)
    np.testing.assert_array_equal(v_d, val_o)
    rz.device_free(d_r); rz.device_free(d_v)
    cs.close()


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq', 'hanford300a_mr'])
@pytest.mark.parametrize('generic', [0, 1])
def test_flux_unstructured_long_rows(name, generic, monkeypatch):
    """Unstructured connectivity with hub cells (rows of 100+ connections: beyond the kernels' register window of six entries, and
    tiles whose rows have very different lengths), repeated pairs and ghost cells; bit for bit against the oracle loop."""
    monkeypatch.setenv('RXN_FLUX_GENERIC', str(generic))
    ncells = 1000
    w, cells = workload_cells(name, ncells)
    st = synth.host_state(w, cells)
    n = w.tables.naqcomp
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.2 * np.random.default_rng(1).standard_normal((ncells, n))))
    o = Oracle(w.tables)
    o.update_auxvars(st, xx, True)
    conn, nghosted, nlocal, active = random_connections(ncells, 6000, n, nghost=37)
    Tu, Td = o.flux_coefs(conn, n)
    r_o = o.flux_residual(st, conn, Tu, Td, nlocal)
    rp_o, col_o, val_o = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    assert np.diff(rp_o).max() > 100
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, ncells)
    rz.upload_host_state(st)
    rz.upload('DTOTAL', st['DTOTAL'])
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal, conn['g2l'])
    rp, col = cs.structure()
    np.testing.assert_array_equal(rp, rp_o)
    np.testing.assert_array_equal(col, col_o)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'])
    np.testing.assert_array_equal(rz.RTResidualFlux(cs), r_o)
    np.testing.assert_array_equal(rz.RTJacobianFlux(cs), val_o)
    cs.close()


def test_flux_after_device_auxvars():
    """The live sequence of a Newton iteration: RTUpdateAuxVars on the device (total, dtotal), then the flux kernels, against the
    oracle doing the same on the host; compared on the scale of the flux terms (a cell's fluxes nearly cancel)."""
    w, st, xx, conn, nlocal = _setup('hanford300a_eq', 12, 6, 5, 0, 0.0)
    n = w.tables.naqcomp
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, st.ncells)
    rz.upload_host_state(st)
    rz.materialize('DTOTAL')
    rz.RTUpdateAuxVars(xx, True)
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'])
    r_g = rz.RTResidualFlux(cs)
    val_g = rz.RTJacobianFlux(cs)
    o = Oracle(w.tables)
    o.update_auxvars(st, xx, True)
    Tu, Td = o.flux_coefs(conn, n)
    r_o = o.flux_residual(st, conn, Tu, Td, nlocal)
    _, _, val_o = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    tmag = np.abs(st['PRI_MOLAL']).T + (np.abs(st['TOTAL']).T)      # scale of total_i (cancelling stoichiometries)
    rscale = 6 * np.abs(Tu).max(axis=0)[None, :] * tmag.max(axis=0)[None, :]
    assert (np.abs(r_g - r_o) <= RTOL * rscale).all()
    vscale = np.abs(val_o).max(axis=0)[None, :]
    assert (np.abs(val_g - val_o) <= RTOL * np.maximum(np.abs(val_o), vscale)).all()


def test_flux_entry_point_errors():
    w, st, xx, conn, nlocal = _setup('calcite', 4, 3, 2, 0, 0.0)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, st.ncells)
    rz.upload_host_state(st)
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal)
    with pytest.raises(rt.RxnError):            # coefficients not set
        rz.RTResidualFlux(cs)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'])
    with pytest.raises(rt.RxnError):            # dtotal not materialised
        rz.RTJacobianFlux(cs)
    bad = conn['id_dn'].copy(); bad[0] = 10 ** 6
    with pytest.raises(rt.RxnError):
        rt.ConnectionSet(rz, conn['id_up'], bad, nlocal)


def test_flux_full_size_properties():
    """BASELINE config 2 size (100 x 100 x 100, calcite): size-independent properties of the whole-grid result — interior
    fluxes cancel in the sum over cells; row r's diagonal block is minus the sum of the blocks its neighbours hold for
    column r (each connection contributes +J to one row and -J to the other); a seeded sample of rows equals the oracle
    loop restricted to those rows' connections bit for bit."""
    nx = ny = nz = 100
    n_cells = nx * ny * nz
    w, cells = workload_cells('calcite', n_cells)
    n = w.tables.naqcomp
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n_cells)
    for f, v in w.base.items():
        rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    rz.materialize('DTOTAL')
    rng = np.random.default_rng(3)
    xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.2 * rng.standard_normal((n_cells, n))))
    rz.RTUpdateAuxVars(xx, True)
    conn, nghosted, nlocal, active = structured_connections(nx, ny, nz, n)
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'])
    r = rz.RTResidualFlux(cs)
    val = rz.RTJacobianFlux(cs)
    row_ptr, col = cs.structure()
    assert cs.nnz_blocks == n_cells + 2 * len(conn['id_up'])
    tot = rz.download('TOTAL')
    Tmax = (np.abs(conn['velocity']).max() + conn['disp'].max()) * conn['area'].max() * 1000.0
    assert (np.abs(r.sum(axis=0)) <= 1e-9 * Tmax * np.abs(tot).max(axis=1) * np.sqrt(len(conn['id_up']))).all()
    # column sums of the block matrix vanish: sum over rows of block (row, c) = 0 for every column cell c
    colsum = np.zeros((n_cells, n * n))
    np.add.at(colsum, col, val)
    scale = np.zeros((n_cells, n * n))
    np.add.at(scale, col, np.abs(val))
    assert (np.abs(colsum) <= 1e-12 * scale + 1e-300).all()
    # sampled rows against the oracle's connection loop on the sub-list of connections that touch them
    st = synth.host_state(w, cells)
    st['TOTAL'][:] = tot
    st['DTOTAL'][:] = rz.download('DTOTAL')
    sample = np.sort(rng.choice(n_cells, 2000, replace=False))
    mark = np.zeros(n_cells, dtype=bool); mark[sample] = True
    keep = mark[conn['id_up']] | mark[conn['id_dn']]
    sub = {k: (v[keep] if isinstance(v, np.ndarray) and len(v) == len(keep) else v) for k, v in conn.items()}
    o = Oracle(w.tables)
    Tu, Td = o.flux_coefs(sub, n)
    r_o = o.flux_residual(st, sub, Tu, Td, n_cells)
    np.testing.assert_array_equal(r[sample], r_o[sample])
    rp_o, col_o, val_o = o.flux_jacobian(st, sub, Tu, Td, n_cells)
    for c in sample[:500]:
        np.testing.assert_array_equal(col[row_ptr[c]:row_ptr[c + 1]], col_o[rp_o[c]:rp_o[c + 1]])
        np.testing.assert_array_equal(val[row_ptr[c]:row_ptr[c + 1]], val_o[rp_o[c]:rp_o[c + 1]])
    cs.close()


@pytest.mark.parametrize('name,dims,ghost,inactive,upwind', [('calcite', (17, 9, 5), 0, 0.0, True), ('calcite', (13, 7, 3), 1, 0.1, False),
                                                             ('hanford300a_eq', (9, 6, 5), 1, 0.05, True)])
def test_boundary_and_source_sink_bitwise(name, dims, ghost, inactive, upwind):
    """Boundary-condition and source/sink connections (rxn_couplerset_* / rxn_coupler_*_batch) on top of the interior-flux
    result, against the oracle's restatement of the reference loops: bit for bit, corner cells (three faces on one cell), two
    wells in one cell and inactive cells included.  The boundary auxvars are a second realization with one cell per
    connection, built as RTUpdateAuxVars does (reactive_transport.F90:3851-4030: Dirichlet / zero gradient / Dirichlet-zero-
    gradient faces) and updated on the GPU; its totals are handed over on the device."""
    w, st, xx, conn, nlocal = _setup(name, *dims, ghost, inactive)
    n = w.tables.naqcomp
    o = Oracle(w.tables)
    o.update_auxvars(st, xx, True)
    Tu, Td = o.flux_coefs(conn, n, use_upwinding=upwind)
    r_o = o.flux_residual(st, conn, Tu, Td, nlocal)
    rp_o, col_o, val_o = o.flux_jacobian(st, conn, Tu, Td, nlocal)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, st.ncells)
    rz.upload_host_state(st)
    rz.upload('DTOTAL', st['DTOTAL'])
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal, conn['g2l'], st.active)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'], conn['fraction_upwind'], upwind)
    r_g = rz.RTResidualFlux(cs)
    val_g = rz.RTJacobianFlux(cs)
    # ---- boundary faces: the boundary realization
    bc = boundary_connections(*dims, n, ghost_layers=ghost)
    nb = len(bc['id_dn'])
    basis_molarity = np.tile(w.base['PRI_MOLAL'] * w.base['DEN_KG'][0] * 1.0e-3 * 1.2, (nb, 1))
    den_bc = np.full(nb, w.base['DEN_KG'][0])
    xxbc = rz.boundary_free_ion(bc['bc_type'], basis_molarity, den_bc, bc['velocity'], xx[bc['id_dn']])
    zg = (bc['bc_type'] == 4) | ((bc['bc_type'] == 3) & (bc['velocity'] < 0))
    assert (xxbc[zg] == xx[bc['id_dn']][zg]).all() and (xxbc[~zg] != xx[bc['id_dn']][~zg]).any()
    wb, cells_b = workload_cells(name, nb)
    st_b = synth.host_state(wb, cells_b)
    bz = rt.Realization(rx, nb)
    bz.upload_host_state(st_b)
    bz.RTUpdateAuxVars(xxbc, True)
    o.update_auxvars(st_b, xxbc, True)
    ext = np.ascontiguousarray(st_b['TOTAL'].T)
    np.testing.assert_allclose(bz.download('TOTAL').T, ext, rtol=1e-12)
    bs = rt.CouplerSet(rz, abi.RXN_COUPLER_BOUNDARY, bc['id_dn'], nlocal, conn['g2l'], st.active)
    bs.TFluxCoefBC(bc['area'], bc['velocity'], bc['disp'], upwind)
    bs.set_totals(ext)                                     # the oracle's totals, so that the comparison below is bit for bit
    cu, cd = Oracle.flux_coefs({**bc, 'fraction_upwind': np.full(nb, 0.5)}, n, use_upwinding=upwind)
    f_o = o.coupler_residual(st, 0, bc['id_dn'], ext, cu, cd, nlocal, r_o, g2l=conn['g2l'], want_flux=True)
    d_o = np.ascontiguousarray(val_o[rp_o[:-1]])
    o.coupler_jacobian(st, 0, bc['id_dn'], cd, nlocal, d_o, g2l=conn['g2l'])
    val_o[rp_o[:-1]] = d_o
    f_g = rz.RTResidualCoupler(bs, r_g, want_flux=True)
    rz.RTJacobianCoupler(bs, val_g, cs)
    np.testing.assert_array_equal(r_g, r_o)
    live = st.active[bc['id_dn']] != 0
    np.testing.assert_array_equal(f_g[live], f_o[live])
    np.testing.assert_array_equal(val_g, val_o)
    # the same faces with the totals taken from the boundary realization on the device: same result to rounding of its totals
    bs.totals_from_state(bz)
    r2 = rz.RTResidualFlux(cs)
    rz.RTResidualCoupler(bs, r2)
    scale = np.abs(cu).max() * np.abs(ext).max(axis=0)
    assert (np.abs(r2 - r_o) <= 1e-11 * scale[None, :]).all()
    # ---- source/sinks, onto plain diagonal blocks (the layout of rxn_jacobian_blocks_batch)
    local_g = np.where(conn['g2l'] >= 0)[0] if conn['g2l'] is not None else np.arange(nlocal)
    ss = source_sinks(local_g, n)
    tin, tout = Oracle.ss_coefs(ss['qsrc'], ss['ss_type'])
    ext_s = np.ascontiguousarray(np.tile(w.base['TOTAL'] * 0.7, (len(tin), 1)))
    c_in, c_out = np.repeat(tin[:, None], n, 1).copy(), np.repeat(tout[:, None], n, 1).copy()
    o.coupler_residual(st, 1, ss['id_dn'], ext_s, c_out, c_in, nlocal, r_o, g2l=conn['g2l'])
    diag_o = np.zeros((nlocal, n * n))
    o.coupler_jacobian(st, 1, ss['id_dn'], c_in, nlocal, diag_o, g2l=conn['g2l'])
    sk = rt.CouplerSet(rz, abi.RXN_COUPLER_SRC_SINK, ss['id_dn'], nlocal, conn['g2l'], st.active)
    sk.TSrcSinkCoef(ss['qsrc'], ss['ss_type'])
    sk.set_totals(ext_s)
    rz.RTResidualCoupler(sk, r_g)
    diag_g = np.zeros((nlocal, n * n))
    rz.RTJacobianCoupler(sk, diag_g)
    np.testing.assert_array_equal(r_g, r_o)
    np.testing.assert_array_equal(diag_g, diag_o)
    assert np.abs(diag_o).max() > 0
    # misuse: coefficients / totals missing, foreign state
    b2 = rt.CouplerSet(rz, abi.RXN_COUPLER_BOUNDARY, bc['id_dn'], nlocal, conn['g2l'], st.active)
    with pytest.raises(rt.RxnError) as e:
        rz.RTResidualCoupler(b2, r_g)
    assert e.value.status == abi.RXN_ERR_INVALID
    with pytest.raises(rt.RxnError) as e:
        bz.RTResidualCoupler(bs, np.zeros((nlocal, n)))
    assert e.value.status == abi.RXN_ERR_INVALID
    for x in (b2, sk, bs, cs):
        x.close()
