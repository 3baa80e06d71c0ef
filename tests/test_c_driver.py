"""A plain-C99 host (tests/c_driver/rxn_driver.c) walks the call sequence of the reference's call sites with nothing but
include/rxn_b200.h: CPU-only part = the header is valid C, the driver compiles with -pedantic and links
against librxn_b200.so; GPU part = it runs, and its results equal the same calls made through ctypes bit for bit (same
library, same inputs) and the oracle to the parity bar."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, 'tests', 'c_driver')
sys.path.insert(0, DRV)

from pflotran_b200 import abi, synth, reactive_transport as rt  # noqa: E402


def _build(tmp_path, name, ncells):
    import gen_case
    case = os.path.join(tmp_path, 'case_%s.h' % name)
    gen_case.write_case(name, ncells, case)
    exe = os.path.join(tmp_path, 'rxn_driver_%s' % name)
    libdir = os.path.dirname(rt.LIB_PATH)
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-O1', '-I', os.path.join(ROOT, 'include'),
                           '-DCASE_HEADER="%s"' % case, os.path.join(DRV, 'rxn_driver.c'), '-o', exe, '-L', libdir,
                           '-lrxn_b200', '-lm', '-Wl,-rpath,' + libdir])
    return exe


def test_header_is_c_and_driver_links(tmp_path):
    exe = _build(str(tmp_path), 'calcite', 8)
    out = subprocess.run([exe, os.path.join(str(tmp_path), 'o.bin'), 'link-check'], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert 'pflotran_b200' in out.stdout                      # rxn_version() through the C binding


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq'])
def test_c_driver_matches_ctypes_path_and_oracle(name, tmp_path):
    from oracle.pyoracle import Oracle
    n = 512
    exe = _build(str(tmp_path), name, n)
    outp = os.path.join(str(tmp_path), 'out.bin')
    r = subprocess.run([exe, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    w = synth.Workload(name)
    nc = w.ncomp
    raw = np.fromfile(outp)
    sizes = [n * nc, n * nc, n * nc * nc, n * nc, 2 * n]
    assert raw.size == sum(sizes)
    accum, res, jac, xx_c, itfl = np.split(raw, np.cumsum(sizes)[:-1])
    # the same sequence through ctypes
    cells = synth.make_cells(w, 0, n)
    rx = rt.Reaction(w.tables)
    rz = rt.Realization(rx, n)
    for f, v in w.base.items():
        if v.size:
            rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    if w.tables.nkinmnrl:
        rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    xx = np.ascontiguousarray(np.tile(w.base['PRI_MOLAL'] * 1.02, (n, 1)))
    rz.RTUpdateAuxVars(xx, True)
    a_p = rz.RTUpdateFixedAccumulation(xx)
    r_p, j_p = rz.RTResidualJacobianNonFlux(1800.0)
    rz.RTUpdateKineticState(1800.0)
    x_p = cells['tran_xx'].copy()
    it_p, fl_p = rz.RTReact(x_p, 3600.0, abi.RXN_DT_CONSISTENT)
    np.testing.assert_array_equal(accum.reshape(n, nc), a_p)
    np.testing.assert_array_equal(res.reshape(n, nc), r_p)
    np.testing.assert_array_equal(jac.reshape(n, nc * nc), j_p)
    np.testing.assert_array_equal(xx_c.reshape(n, nc), x_p)
    np.testing.assert_array_equal(itfl[:n].astype(np.int32), it_p)
    np.testing.assert_array_equal(itfl[n:].astype(np.int32), fl_p)
    # and the oracle on the operator-split result
    st = synth.host_state(w, cells)
    o = Oracle(w.tables)
    o.update_auxvars(st, xx, True)
    o.fixed_accum(st, xx)
    o.residual_jacobian(st, 1800.0)
    o.update_kinetic_state(st, 1800.0)
    xo = cells['tran_xx'].copy()
    it_o, fl_o = o.react(st, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=4)
    assert (it_o == it_p).all() and (fl_o == fl_p).all()
    ok = (fl_o & ~3) == 0
    assert (np.abs(x_p[ok] - xo[ok]) / np.abs(xo[ok])).max() <= 1e-10
