"""ctypes wrapper of tests/emul/libemul.so — the HOST compilation of the CUDA device routines.

TEST INFRASTRUCTURE (see tests/emul/emul.cpp): lets the CPU-only suite compare the device
code's logic and the table packer with the oracle.  Never imported by pflotran_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from pflotran_b200 import abi

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emul')
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'libemul.so')
    deps = [os.path.join(_HERE, 'emul.cpp')] + [os.path.join(_ROOT, 'pflotran_b200', 'csrc', f)
                                                 for f in ('rxn_device.cuh', 'rxn_pack.h', 'rxn_tab.h', 'rxn_lane.h', 'rxn_lane_dev.cuh', 'rxn_tm_dev.cuh', 'rxn_flux.h')] + \
        [os.path.join(_ROOT, 'include', 'rxn_b200.h')]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off', '-pthread',
                               '-o', so, os.path.join(_HERE, 'emul.cpp')])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.emu_create.restype = C.c_void_p
        _LIB.emu_create.argtypes = [C.POINTER(abi.RxnTablesDesc), C.c_char_p, C.c_int]
        _LIB.emu_destroy.argtypes = [C.c_void_p]
        _LIB.emu_set_maxit.argtypes = [C.c_void_p, C.c_int]
        _LIB.emu_pack_status.argtypes = [C.POINTER(abi.RxnTablesDesc), C.c_char_p, C.c_int]
    return _LIB


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def pack_status(tables_or_desc):
    d = tables_or_desc if isinstance(tables_or_desc, abi.RxnTablesDesc) else abi.make_desc(tables_or_desc)
    buf = C.create_string_buffer(1024)
    rc = lib().emu_pack_status(C.byref(d), buf, 1024)
    return rc, buf.value.decode()


class Emulator:
    def __init__(self, tables):
        self.t = tables
        self.desc = abi.make_desc(tables)
        buf = C.create_string_buffer(1024)
        h = lib().emu_create(C.byref(self.desc), buf, 1024)
        if not h:
            raise RuntimeError(buf.value.decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                lib().emu_destroy(self.h)
        except Exception:
            pass

    def react(self, st, tran_xx, dt, dt_mode=abi.RXN_DT_CONSISTENT, l2g=None, maxit=None):
        if maxit is not None:
            lib().emu_set_maxit(self.h, maxit)
        n = tran_xx.shape[0]
        iters = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        v = st.view()
        rc = lib().emu_react_batch(self.h, C.byref(v), _p(tran_xx, C.c_double), _p(st.active, C.c_uint8),
                                   _p(l2g, C.c_int32), C.c_int64(n), C.c_double(dt), C.c_int(dt_mode),
                                   _p(iters, C.c_int32), _p(flags, C.c_int32))
        assert rc == 0
        return iters, flags

    def react_lane(self, st, tran_xx, dt, dt_mode=abi.RXN_DT_CONSISTENT, l2g=None, maxit=None, G=1, N=0):
        """RReact through the resident-lane kernel's per-lane routines (rxn_lane_dev.cuh), G lanes per cell."""
        if maxit is not None:
            lib().emu_set_maxit(self.h, maxit)
        n = tran_xx.shape[0]
        iters = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        v = st.view()
        buf = C.create_string_buffer(512)
        stats = np.zeros(16, dtype=np.int32)
        rc = lib().emu_react_lane(self.h, C.byref(v), _p(tran_xx, C.c_double), _p(st.active, C.c_uint8),
                                  _p(l2g, C.c_int32), C.c_int64(n), C.c_double(dt), C.c_int(dt_mode),
                                  _p(iters, C.c_int32), _p(flags, C.c_int32), C.c_int(G), C.c_int(N), buf, 512, _p(stats, C.c_int32))
        if rc != 0:
            raise NotImplementedError(buf.value.decode())
        self.lane_stats = dict(zip(['N', 'blob_bytes', 'cell_bytes', 'terms_spec', 'steps_spec', 'terms_A', 'steps_A',
                                    'terms_B', 'steps_B', 'ncls'], stats.tolist()))
        return iters, flags

    def react_tm(self, st, tran_xx, dt, dt_mode=abi.RXN_DT_CONSISTENT, l2g=None, maxit=None, G=2, N=0):
        """RReact through the tensor-memory kernel's routines (rxn_tm_dev.cuh), G member warps per cell."""
        if maxit is not None:
            lib().emu_set_maxit(self.h, maxit)
        n = tran_xx.shape[0]
        iters = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        v = st.view()
        buf = C.create_string_buffer(512)
        rc = lib().emu_react_tm(self.h, C.byref(v), _p(tran_xx, C.c_double), _p(st.active, C.c_uint8),
                                _p(l2g, C.c_int32), C.c_int64(n), C.c_double(dt), C.c_int(dt_mode),
                                _p(iters, C.c_int32), _p(flags, C.c_int32), C.c_int(G), C.c_int(N), buf, 512)
        if rc != 0:
            raise NotImplementedError(buf.value.decode())
        return iters, flags

    def gi_tm(self, st, mode, G=3, update_act=False, xx=None, xx_by_item=False, dt=1.0, l2g=None, want_accum=False):
        """Global-implicit pass on the tensor-memory layout (tm_gi_cell, rxn_tm_dev.cuh): mode 1 = auxvar update / fixed
        accumulation, 2 = residual + Jacobian blocks.  Returns (accum | None) for mode 1, (res, jac) for mode 2."""
        n = st.ncells if l2g is None else len(l2g)
        nc = self.t.ncomp
        acc = np.zeros((n, nc)) if (mode == 1 and want_accum) else None
        res = np.zeros((n, nc)) if mode == 2 else None
        jac = np.zeros((n, nc * nc)) if mode == 2 else None
        v = st.view()
        buf = C.create_string_buffer(512)
        rc = lib().emu_gi_tm(self.h, C.byref(v), _p(st.active, C.c_uint8), _p(l2g, C.c_int32), C.c_int64(n), C.c_int(mode),
                             C.c_int(int(update_act)), _p(xx, C.c_double), C.c_int(int(xx_by_item)), _p(acc, C.c_double),
                             _p(res, C.c_double), _p(jac, C.c_double), C.c_double(dt), C.c_int(G), buf, 512)
        if rc != 0:
            raise NotImplementedError(buf.value.decode())
        return acc if mode == 1 else (res, jac)

    def update_auxvars(self, st, xx_loc, update_act_coefs):
        v = st.view()
        assert lib().emu_update_auxvars_batch(self.h, C.byref(v), _p(xx_loc, C.c_double), _p(st.active, C.c_uint8),
                                              C.c_int(int(update_act_coefs))) == 0

    def fixed_accum(self, st, xx, l2g=None):
        n = st.ncells if l2g is None else len(l2g)
        out = np.zeros((n, self.t.ncomp))
        v = st.view()
        assert lib().emu_fixed_accum_batch(self.h, C.byref(v), _p(xx, C.c_double), _p(st.active, C.c_uint8),
                                           _p(l2g, C.c_int32), C.c_int64(n), _p(out, C.c_double)) == 0
        return out

    def residual_jacobian(self, st, dt, l2g=None):
        n = st.ncells if l2g is None else len(l2g)
        nc = self.t.ncomp
        res = np.zeros((n, nc))
        jac = np.zeros((n, nc * nc))
        v = st.view()
        assert lib().emu_residual_jacobian_batch(self.h, C.byref(v), _p(st.active, C.c_uint8), _p(l2g, C.c_int32),
                                                 C.c_int64(n), C.c_double(dt), _p(res, C.c_double),
                                                 _p(jac, C.c_double)) == 0
        return res, jac

    def residual_jacobian_lane(self, st, dt, l2g=None, G=1):
        """Global-implicit blocks through the resident-lane routines (rxn_lane_dev.cuh: lane_gi_cell)."""
        n = st.ncells if l2g is None else len(l2g)
        nc = self.t.ncomp
        res = np.zeros((n, nc))
        jac = np.zeros((n, nc * nc))
        v = st.view()
        buf = C.create_string_buffer(512)
        rc = lib().emu_gi_lane(self.h, C.byref(v), _p(st.active, C.c_uint8), _p(l2g, C.c_int32), C.c_int64(n), C.c_double(dt),
                               _p(res, C.c_double), _p(jac, C.c_double), C.c_int(G), buf, 512)
        if rc != 0:
            raise NotImplementedError(buf.value.decode())
        return res, jac

    def equilibrate_batch(self, st, ctype, conc, cid, free_ion_guess=None, use_prev=False, molal=True):
        n = st.ncells
        ctype = np.ascontiguousarray(ctype, dtype=np.int32); cid = np.ascontiguousarray(cid, dtype=np.int32)
        conc = np.ascontiguousarray(conc, dtype=np.float64)
        stride = 0 if conc.ndim == 1 else conc.shape[1]
        guess = None if free_ion_guess is None else np.ascontiguousarray(free_ion_guess, dtype=np.float64)
        basis = np.zeros((n, self.t.naqcomp)); iters = np.zeros(n, dtype=np.int32); status = np.zeros(n, dtype=np.int32)
        v = st.view()
        assert lib().emu_equilibrate_batch(self.h, C.byref(v), _p(st.active, C.c_uint8), _p(ctype, C.c_int32), _p(conc, C.c_double),
                                           C.c_int64(stride), _p(cid, C.c_int32), _p(guess, C.c_double), C.c_int(int(use_prev)),
                                           C.c_int(int(molal)), C.c_int64(n), _p(basis, C.c_double), _p(iters, C.c_int32),
                                           _p(status, C.c_int32)) == 0
        return basis, iters, status

    def update_kinetic_state(self, st, dt):
        v = st.view()
        assert lib().emu_update_kinetic_state_batch(self.h, C.byref(v), _p(st.active, C.c_uint8), C.c_double(dt)) == 0


def flux(st, conn, nlocal, use_upwinding=True):
    """Flux residual / Jacobian through the structure builder and per-row arithmetic of rxn_flux.h (what the CUDA kernels implement).
    conn: dict(id_up, id_dn, g2l|None, area, velocity, disp [nconn, naq], fraction_upwind).  Returns row_ptr, col, res, val."""
    n = st.t.naqcomp
    nconn = len(conn['id_up'])
    v = st.view()
    L = lib()
    L.emu_flux.restype = C.c_int64
    row_ptr = np.zeros(nlocal + 1, dtype=np.int32)
    args = lambda col, res, val: (C.byref(v), _p(st.active, C.c_uint8), C.c_int(n), C.c_int64(nconn), _p(conn['id_up'], C.c_int32),
                                  _p(conn['id_dn'], C.c_int32), _p(conn.get('g2l'), C.c_int32), C.c_int64(nlocal),
                                  _p(conn['area'], C.c_double), _p(conn['velocity'], C.c_double), _p(conn['disp'], C.c_double),
                                  _p(conn['fraction_upwind'], C.c_double), C.c_int(int(use_upwinding)), _p(row_ptr, C.c_int32),
                                  _p(col, C.c_int32), _p(res, C.c_double), _p(val, C.c_double))
    nnzb = L.emu_flux(*args(None, None, None))
    if nnzb < 0:
        raise ValueError('connection set rejected')
    col = np.zeros(nnzb, dtype=np.int32)
    res = np.zeros((nlocal, n))
    val = np.zeros((nnzb, n * n))
    assert L.emu_flux(*args(col, res, val)) == nnzb
    # the same Jacobian by block columns (the traversal of k_flux_jacobian_cols): every block written exactly once, same values
    L.emu_flux_cols.restype = C.c_int64
    val_c = np.full((nnzb, n * n), np.nan)
    assert L.emu_flux_cols(C.byref(v), _p(st.active, C.c_uint8), C.c_int(n), C.c_int64(nconn), _p(conn['id_up'], C.c_int32),
                           _p(conn['id_dn'], C.c_int32), _p(conn.get('g2l'), C.c_int32), C.c_int64(nlocal), _p(conn['area'], C.c_double),
                           _p(conn['velocity'], C.c_double), _p(conn['disp'], C.c_double), _p(conn['fraction_upwind'], C.c_double),
                           C.c_int(int(use_upwinding)), _p(val_c, C.c_double)) == nnzb
    assert np.array_equal(val_c, val), 'column walk and row walk of the flux Jacobian differ'
    return row_ptr, col, res, val


def coupler(st, kind, id_dn, nlocal, ext_total, g2l=None, area=None, velocity=None, disp=None, use_upwinding=True, qsrc=None, ss_type=None,
            res=None, diag=None, want_flux=False):
    """Boundary (kind 0) / source-sink (kind 1) connections through the row view and per-row arithmetic of rxn_flux.h.
    res [nlocal, n] and diag [nlocal, n*n] are updated in place; returns flux_out [nconn, n] or None."""
    n = st.t.naqcomp
    nconn = len(id_dn)
    v = st.view()
    flux = np.zeros((nconn, n)) if want_flux else None
    rc = lib().emu_coupler(C.byref(v), _p(st.active, C.c_uint8), C.c_int(kind), C.c_int(n), C.c_int64(nconn), _p(id_dn, C.c_int32),
                           _p(g2l, C.c_int32), C.c_int64(nlocal), _p(area, C.c_double), _p(velocity, C.c_double), _p(disp, C.c_double),
                           C.c_int(int(use_upwinding)), _p(qsrc, C.c_double), _p(ss_type, C.c_int32), _p(ext_total, C.c_double),
                           _p(res, C.c_double), _p(flux, C.c_double), _p(diag, C.c_double))
    if rc != 0:
        raise ValueError('coupler set rejected')
    return flux
