"""ctypes wrapper of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from pflotran_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, 'liborc.so')
    src = os.path.join(_HERE, 'rxn_oracle.cpp')
    hdr = os.path.join(_HERE, '..', 'include', 'rxn_b200.h')
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'liborc.so'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_create.argtypes = [C.POINTER(abi.RxnTablesDesc)]
        _LIB.orc_destroy.argtypes = [C.c_void_p]
    return _LIB


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _f64(a):
    return a.ctypes.data_as(abi.c_f64p) if a is not None else None


def _i32(a):
    return a.ctypes.data_as(abi.c_i32p) if a is not None else None


class Oracle:
    def __init__(self, tables):
        self.t = tables
        self.desc = abi.make_desc(tables)
        self.h = C.c_void_p(lib().orc_create(C.byref(self.desc)))

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
        except Exception:
            pass

    def react(self, st: abi.HostState, tran_xx: np.ndarray, dt: float, dt_mode: int = abi.RXN_DT_CONSISTENT,
              maxit: int = 1000, nthreads: int = 1):
        """RTReact loop.  tran_xx [ncells, ncomp] in: totals, out: free-ion (in place)."""
        assert tran_xx.flags.c_contiguous and tran_xx.dtype == np.float64
        n = st.ncells
        iters = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        v = st.view()
        rc = lib().orc_react_batch(self.h, C.byref(v), _f64(tran_xx), _u8(st.active), C.c_double(dt),
                                   C.c_int(dt_mode), _i32(iters), _i32(flags), C.c_int(maxit), C.c_int(nthreads))
        assert rc == 0
        return iters, flags

    def update_auxvars(self, st: abi.HostState, xx_loc: Optional[np.ndarray], update_act_coefs: bool,
                       nthreads: int = 1):
        v = st.view()
        rc = lib().orc_update_auxvars_batch(self.h, C.byref(v), _f64(xx_loc), _u8(st.active),
                                            C.c_int(int(update_act_coefs)), C.c_int(nthreads))
        assert rc == 0

    def fixed_accum(self, st: abi.HostState, xx: Optional[np.ndarray], nthreads: int = 1) -> np.ndarray:
        out = np.zeros((st.ncells, self.t.ncomp))
        v = st.view()
        rc = lib().orc_fixed_accum_batch(self.h, C.byref(v), _f64(xx), _u8(st.active), _f64(out), C.c_int(nthreads))
        assert rc == 0
        return out

    def residual_jacobian(self, st: abi.HostState, dt: float, nthreads: int = 1):
        n = self.t.ncomp
        res = np.zeros((st.ncells, n))
        jac = np.zeros((st.ncells, n * n))
        v = st.view()
        rc = lib().orc_residual_jacobian_batch(self.h, C.byref(v), _u8(st.active), C.c_double(dt), _f64(res),
                                               _f64(jac), C.c_int(nthreads))
        assert rc == 0
        return res, jac

    def update_kinetic_state(self, st: abi.HostState, dt: float, nthreads: int = 1):
        v = st.view()
        rc = lib().orc_update_kinetic_state_batch(self.h, C.byref(v), _u8(st.active), C.c_double(dt), C.c_int(nthreads))
        assert rc == 0

    def activity_coefficients(self, st: abi.HostState, nthreads: int = 1):
        v = st.view()
        assert lib().orc_activity_coefficients_batch(self.h, C.byref(v), C.c_int(nthreads)) == 0

    def equilibrate(self, st: abi.HostState, cell: int, ctype, conc, cid, free_ion_guess=None,
                    use_prev: bool = False, molal: Optional[bool] = None):
        naq = self.t.naqcomp
        ctype = np.ascontiguousarray(ctype, dtype=np.int32)
        conc = np.ascontiguousarray(conc, dtype=np.float64)
        cid = np.ascontiguousarray(cid, dtype=np.int32)
        fig = None if free_ion_guess is None else np.ascontiguousarray(free_ion_guess, dtype=np.float64)
        basis = np.zeros(naq)
        nit = C.c_int32(0)
        if molal is None:
            molal = bool(self.t.initialize_with_molality)
        v = st.view()
        rc = lib().orc_equilibrate_constraint(self.h, C.byref(v), C.c_int64(cell), _i32(ctype), _f64(conc),
                                              _i32(cid), _f64(fig), C.c_int(int(use_prev)), C.c_int(int(molal)),
                                              _f64(basis), C.byref(nit))
        if rc != 0:
            raise RuntimeError('ReactionEquilibrateConstraint failed rc=%d' % rc)
        return basis, nit.value

    # ---- flux side: TFluxCoef / RTResidualFlux / RTJacobianFlux interior loops
    @staticmethod
    def flux_coefs(conn, naq, use_upwinding=True):
        nconn = len(conn['id_up'])
        Tu = np.zeros((nconn, naq))
        Td = np.zeros((nconn, naq))
        assert lib().orc_flux_coefs(C.c_int(naq), C.c_int64(nconn), _f64(conn['area']), _f64(conn['velocity']), _f64(conn['disp']),
                                    _f64(conn['fraction_upwind']), C.c_int(int(use_upwinding)), _f64(Tu), _f64(Td)) == 0
        return Tu, Td

    def flux_residual(self, st: abi.HostState, conn, Tu, Td, nlocal):
        naq = self.t.naqcomp
        r = np.zeros((nlocal, naq))
        v = st.view()
        assert lib().orc_flux_residual(C.byref(v), _u8(st.active), C.c_int(naq), C.c_int64(len(conn['id_up'])), _i32(conn['id_up']),
                                       _i32(conn['id_dn']), _i32(conn.get('g2l')), _f64(Tu), _f64(Td), C.c_int64(nlocal), _f64(r)) == 0
        return r

    def flux_jacobian(self, st: abi.HostState, conn, Tu, Td, nlocal):
        naq = self.t.naqcomp
        v = st.view()
        L = lib()
        L.orc_flux_jacobian.restype = C.c_int64
        row_ptr = np.zeros(nlocal + 1, dtype=np.int32)
        a = lambda col, val: (C.byref(v), _u8(st.active), C.c_int(naq), C.c_int64(len(conn['id_up'])), _i32(conn['id_up']),
                              _i32(conn['id_dn']), _i32(conn.get('g2l')), _f64(Tu), _f64(Td), C.c_int64(nlocal),
                              C.c_int64(st.ncells), _i32(row_ptr), _i32(col), _f64(val))
        nnzb = L.orc_flux_jacobian(*a(None, None))
        col = np.zeros(nnzb, dtype=np.int32)
        val = np.zeros((nnzb, naq * naq))
        assert L.orc_flux_jacobian(*a(col, val)) == nnzb
        return row_ptr, col, val

    # ---- boundary-condition / source-sink connections (reactive_transport.F90:2347-2430, 3176-3240, 2623-2672, 3394-3436)
    @staticmethod
    def ss_coefs(qsrc, ss_type):
        n = len(qsrc)
        tin, tout = np.zeros(n), np.zeros(n)
        assert lib().orc_ss_coefs(C.c_int64(n), _f64(qsrc), _i32(ss_type), _f64(tin), _f64(tout)) == 0
        return tin, tout

    def coupler_residual(self, st: abi.HostState, kind, id_dn, ext_total, c_ext, c_cell, nlocal, res, g2l=None, want_flux=False):
        """res [nlocal, naq] updated in place; c_ext / c_cell [nconn, naq]; returns flux_out [nconn, naq] or None."""
        naq = self.t.naqcomp
        v = st.view()
        flux = np.zeros((len(id_dn), naq)) if want_flux else None
        assert lib().orc_coupler_residual(C.byref(v), _u8(st.active), C.c_int(kind), C.c_int(naq), C.c_int64(len(id_dn)), _i32(id_dn),
                                          _i32(g2l), _f64(ext_total), _f64(c_ext), _f64(c_cell), C.c_int64(nlocal), _f64(res),
                                          _f64(flux)) == 0
        return flux

    def coupler_jacobian(self, st: abi.HostState, kind, id_dn, c_cell, nlocal, diag, g2l=None):
        """diag [nlocal, naq*naq] (column-major blocks) updated in place."""
        naq = self.t.naqcomp
        v = st.view()
        assert lib().orc_coupler_jacobian(C.byref(v), _u8(st.active), C.c_int(kind), C.c_int(naq), C.c_int64(len(id_dn)), _i32(id_dn),
                                          _i32(g2l), _f64(c_cell), C.c_int64(nlocal), _f64(diag)) == 0

    @staticmethod
    def rsolve(res, jac, conc, use_log):
        n = len(res)
        res = np.array(res, dtype=np.float64)
        jac = np.array(jac, dtype=np.float64, order='F').ravel(order='F').copy()
        conc = np.array(conc, dtype=np.float64)
        upd = np.zeros(n)
        rc = lib().orc_rsolve(_f64(res), _f64(jac), _f64(conc), _f64(upd), C.c_int(n), C.c_int(int(use_log)))
        return rc, upd
