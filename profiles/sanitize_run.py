"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family of the library on batches
small enough for the tools' slowdown, large enough that persistent lanes take several cells from the work counter
(k_react_tm: 148 CTAs x 128 cells = 18 944 resident cells).  No oracle, no timing: the tools' reports are the result.

  compute-sanitizer --tool racecheck python profiles/sanitize_run.py [react|gi|flux|new|pipeline|all] [ncells]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from pflotran_b200 import abi, synth, reactive_transport as rt  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
ncells = int(sys.argv[2]) if len(sys.argv) > 2 else 60000


def react(name, n, kernel=3):
    w = synth.Workload(name)
    cells = synth.make_cells(w, 0, n)
    st = synth.host_state(w, cells)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    rz.upload_host_state(st)
    rz.set_react_kernel(kernel)
    xx = cells['tran_xx'].copy()
    it, fl = rz.RTReact(xx, 3600.0, abi.RXN_DT_CONSISTENT)
    print('react %-16s %6d cells  %s  mean its %.2f  non-reference flags %d' % (name, n, rz.react_kernel_info()[:60], it.mean(),
                                                                               int(((fl != 1) & (fl != 2)).sum())))
    return w, cells, st, rx, rz


if what in ('react', 'all'):
    react('hanford300a_eq', ncells)            # tensor-memory kernel (named barriers, TMEM, persistent lanes, atomicAdd work counter)
    react('hanford300a_mr', ncells // 2)       # + warp-cooperative multirate loads
    os.environ['RXN_TM'] = '0'
    react('hanford300a_eq', ncells // 2)       # resident-lane kernel, J in shared memory (sub-warp __syncwarp groups)
    del os.environ['RXN_TM']
    react('calcite', ncells)                   # resident-lane kernel G = 1, 448 cells per CTA
    react('ion_exchange', ncells // 4, 1)      # thread-per-cell kernel

if what in ('gi', 'all'):
    for name in ('hanford300a_eq', 'hanford300a_mr', 'calcite'):
        w = synth.Workload(name)
        n = ncells // 4
        cells = synth.make_cells(w, 0, n)
        st = synth.host_state(w, cells)
        rx = rt.Reaction(w.tables)
        rz = rt.Realization(rx, n)
        rz.upload_host_state(st)
        xx = np.ascontiguousarray(cells['tran_xx'])
        rz.RTUpdateAuxVars(st['PRI_MOLAL'].T.copy(), True)
        acc = rz.RTUpdateFixedAccumulation(None)
        res, jac = rz.RTResidualJacobianNonFlux(3600.0)
        rz.RTUpdateKineticState(3600.0)
        print('gi    %-16s %6d cells  |res| max %.3e' % (name, n, np.abs(res).max()))

if what in ('flux', 'all'):
    from flux_common import structured_connections
    w = synth.Workload('hanford300a_eq')
    nx, ny, nz = 24, 16, 12
    n = nx * ny * nz
    cells = synth.make_cells(w, 0, n)
    st = synth.host_state(w, cells)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    rz.upload_host_state(st)
    rz.materialize('DTOTAL')
    rz.RTUpdateAuxVars(st['PRI_MOLAL'].T.copy(), True)
    conn, _, nlocal, _ = structured_connections(nx, ny, nz, w.tables.naqcomp)
    cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal)
    cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp'])
    r = rz.RTResidualFlux(cs)
    v = rz.RTJacobianFlux(cs)
    print('flux  %d cells, %d blocks, |res| max %.3e' % (n, cs.nnz_blocks, np.abs(r).max()))
    # boundary faces and wells (coupler sets) on top
    from flux_common import boundary_connections, source_sinks
    bc = boundary_connections(nx, ny, nz, w.tables.naqcomp)
    bs = rt.CouplerSet(rz, abi.RXN_COUPLER_BOUNDARY, bc['id_dn'], nlocal)
    bs.TFluxCoefBC(bc['area'], bc['velocity'], bc['disp'])
    bs.set_totals(np.ascontiguousarray(np.tile(w.base['TOTAL'] * 1.1, (len(bc['id_dn']), 1))))
    rz.RTResidualCoupler(bs, r, want_flux=True)
    rz.RTJacobianCoupler(bs, v, cs)
    ss = source_sinks(np.arange(nlocal), w.tables.naqcomp)
    sk = rt.CouplerSet(rz, abi.RXN_COUPLER_SRC_SINK, ss['id_dn'], nlocal)
    sk.TSrcSinkCoef(ss['qsrc'], ss['ss_type'])
    sk.set_totals(np.ascontiguousarray(np.tile(w.base['TOTAL'] * 0.7, (len(ss['id_dn']), 1))))
    rz.RTResidualCoupler(sk, r)
    rz.RTJacobianCoupler(sk, v, cs)
    print('coupler %d boundary faces, %d wells, |res| max %.3e' % (len(bc['id_dn']), len(ss['id_dn']), np.abs(r).max()))
    bs.close(); sk.close()
    cs.close()
if what in ('pipeline',):
    react('calcite', 300000)                   # host-buffer RTReact above 262 144 cells: chunked copies, chunk kernels on two alternating streams
if what in ('new', 'all'):
    # kernels added late in round 2: resident-lane N = 24 with 8 lanes per cell (ascem), thread-per-cell microbial reactions with an
    # immobile dof (RReact and the global-implicit loops), the streaming multirate update k_kinmr_update
    react('ascem', max(2000, ncells // 20))
    react('abcd_microbial', ncells // 4, 1)
    for name in ('abcd_microbial', 'hanford300a_mr'):
        w = synth.Workload(name)
        n = ncells // 4
        cells = synth.make_cells(w, 0, n)
        st = synth.host_state(w, cells)
        rx = rt.Reaction(w.tables)
        rz = rt.Realization(rx, n)
        rz.upload_host_state(st)
        xx = np.ascontiguousarray(np.tile(w.base_solution() * 1.03, (n, 1)))
        rz.RTUpdateAuxVars(xx, True)
        acc = rz.RTUpdateFixedAccumulation(xx)
        res, jac = rz.RTResidualJacobianNonFlux(3600.0)
        rz.RTUpdateKineticState(3600.0)
        print('new   %-16s %6d cells  |res| max %.3e  blocks %s' % (name, n, np.abs(res).max(), jac.shape))
print('sanitize_run done')
