"""Host-side mirror of the reference's reactive-transport cell loops over the C ABI.

The reference boundary is a set of Fortran module procedures called one cell at a time
from src/pflotran/reactive_transport.F90 (RTReact :1605, RTUpdateAuxVars :3704,
RTUpdateActivityCoefficients :3620, RTUpdateFixedAccumulation :726, the accumulation +
reaction loops of RTResidualNonFlux :2436 / RTJacobianNonFlux :3247, RTUpdateKineticState
:642).  `Realization` below offers the same operations, same names and argument meaning, as
batched calls into librxn_b200.so (include/rxn_b200.h).  It is plumbing: ctypes marshalling
only, no chemistry arithmetic, and NO fallback — if the CUDA library cannot be loaded or no
GPU is present the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# RXN_B200_LIB: an alternative build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get('RXN_B200_LIB') or os.path.join(_HERE, 'librxn_b200.so')
_LIB = None

c_i64 = C.c_int64
c_dp = abi.c_f64p
c_ip = abi.c_i32p


class RxnError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__('rxn_b200 status %d: %s' % (status, msg))
        self.status = status


def lib():
    """Load librxn_b200.so (built by __graft_entry__.build() / pflotran_b200/csrc/Makefile)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RxnError(abi.RXN_ERR_NO_DEVICE, 'CUDA library %s is missing: build it with '
                           '`python -c "import __graft_entry__ as g; g.build()"`; there is no CPU fallback' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.rxn_version.restype = C.c_char_p
        L.rxn_launch_count.restype = c_i64
        L.rxn_state_ncells.restype = c_i64
        L.rxn_last_kernel_ms.restype = C.c_float
        L.rxn_tables_create.argtypes = [C.POINTER(abi.RxnTablesDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.rxn_tables_destroy.argtypes = [C.c_void_p]
        L.rxn_state_create.argtypes = [C.c_void_p, c_i64, C.POINTER(C.c_void_p)]
        L.rxn_state_destroy.argtypes = [C.c_void_p]
        L.rxn_state_ncells.argtypes = [C.c_void_p]
        L.rxn_field_rows.argtypes = [C.c_void_p, C.c_int]
        L.rxn_state_materialize.argtypes = [C.c_void_p, C.c_int]
        L.rxn_set_react_kernel.argtypes = [C.c_void_p, C.c_int]
        L.rxn_react_kernel_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
        L.rxn_update_auxvars_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.rxn_residual_jacobian_blocks_batch_device.argtypes = [C.c_void_p, C.c_void_p, c_i64, C.c_double, C.c_void_p, C.c_void_p]
        L.rxn_connset_create.argtypes = [C.c_void_p, c_i64, c_ip, c_ip, c_ip, c_i64, C.POINTER(C.c_uint8), C.POINTER(C.c_void_p)]
        L.rxn_connset_destroy.argtypes = [C.c_void_p]
        L.rxn_connset_structure.argtypes = [C.c_void_p, C.POINTER(c_i64), c_ip, c_ip]
        L.rxn_connset_flux_coefs.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_int]
        L.rxn_flux_residual_batch.argtypes = [C.c_void_p, C.c_void_p, c_dp]
        L.rxn_flux_jacobian_batch.argtypes = [C.c_void_p, C.c_void_p, c_dp]
        L.rxn_flux_residual_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rxn_flux_jacobian_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rxn_couplerset_create.argtypes = [C.c_void_p, C.c_int, c_i64, c_ip, c_ip, c_i64, C.POINTER(C.c_uint8), C.POINTER(C.c_void_p)]
        L.rxn_couplerset_destroy.argtypes = [C.c_void_p]
        L.rxn_couplerset_bc_coefs.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, C.c_int]
        L.rxn_couplerset_ss_coefs.argtypes = [C.c_void_p, c_dp, c_ip]
        L.rxn_couplerset_set_totals.argtypes = [C.c_void_p, c_dp]
        L.rxn_couplerset_totals_from_state.argtypes = [C.c_void_p, C.c_void_p]
        L.rxn_coupler_residual_batch.argtypes = [C.c_void_p, C.c_void_p, c_dp, c_dp]
        L.rxn_coupler_jacobian_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_dp]
        L.rxn_coupler_residual_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rxn_coupler_jacobian_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rxn_equilibrate_constraint_batch.argtypes = [C.c_void_p, c_ip, c_dp, c_i64, c_ip, c_dp, C.c_int, C.c_int, c_ip, c_i64,
                                                       c_dp, c_ip, c_ip]
        L.rxn_state_upload.argtypes = [C.c_void_p, C.c_int, c_dp, c_i64, c_i64]
        L.rxn_state_download.argtypes = [C.c_void_p, C.c_int, c_dp, c_i64, c_i64]
        L.rxn_state_broadcast.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.rxn_set_cell_scalars.argtypes = [C.c_void_p] + [c_dp] * 7 + [C.POINTER(C.c_uint8)]
        L.rxn_react_batch.argtypes = [C.c_void_p, c_dp, c_ip, c_i64, C.c_double, C.c_int, c_ip, c_ip]
        L.rxn_react_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_double, C.c_int,
                                             C.c_void_p, C.c_void_p]
        L.rxn_update_auxvars_batch.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.rxn_fixed_accum_batch.argtypes = [C.c_void_p, c_dp, c_ip, c_i64, c_dp]
        L.rxn_residual_blocks_batch.argtypes = [C.c_void_p, c_ip, c_i64, C.c_double, c_dp]
        L.rxn_jacobian_blocks_batch.argtypes = [C.c_void_p, c_ip, c_i64, C.c_double, c_dp]
        L.rxn_residual_jacobian_blocks_batch.argtypes = [C.c_void_p, c_ip, c_i64, C.c_double, c_dp, c_dp]
        L.rxn_update_kinetic_state_batch.argtypes = [C.c_void_p, C.c_double]
        L.rxn_last_kernel_ms.argtypes = [C.c_void_p]
        L.rxn_state_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(c_i64)]
        L.rxn_device_alloc.argtypes = [C.c_void_p, c_i64, C.POINTER(C.c_void_p)]
        L.rxn_device_free.argtypes = [C.c_void_p, C.c_void_p]
        L.rxn_device_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_int]
        L.rxn_device_sync.argtypes = [C.c_void_p]
        L.rxn_host_alloc.argtypes = [c_i64, C.POINTER(C.c_void_p)]
        L.rxn_host_free.argtypes = [C.c_void_p]
        L.rxn_last_error.argtypes = [C.c_char_p, C.c_int32]
        L.rxn_timer_start.argtypes = [C.c_void_p]
        L.rxn_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.rxn_probe_fp64.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    lib().rxn_last_error(buf, 2048)
    return buf.value.decode(errors='replace')


def _ck(rc: int):
    if rc != abi.RXN_OK:
        raise RxnError(rc, last_error())


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


class Reaction:
    """The shared `reaction` object (reaction_type, reaction_aux.F90:142-335) on one GPU."""

    def __init__(self, tables, device: int = 0):
        self.tables = tables
        self.desc = tables if isinstance(tables, abi.RxnTablesDesc) else abi.make_desc(tables)
        self.h = C.c_void_p()
        _ck(lib().rxn_tables_create(C.byref(self.desc), device, C.byref(self.h)))
        self.device = device

    def field_rows(self, field: str) -> int:
        return lib().rxn_field_rows(self.h, abi.F[field])

    def close(self):
        if self.h:
            lib().rxn_tables_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConnectionSet:
    """Interior connections of a grid (grid%internal_connection_set_list flattened in loop order) bound to a Realization:
    the row view / block-CSR structure of the flux Jacobian and the TFluxCoef coefficients live on the GPU
    (include/rxn_b200.h "Flux side")."""

    def __init__(self, realization: 'Realization', id_up: np.ndarray, id_dn: np.ndarray, nlocal: int,
                 ghost_to_local: Optional[np.ndarray] = None, active: Optional[np.ndarray] = None):
        self.rz = realization
        self.nlocal = int(nlocal)
        self.nconn = len(id_up)
        id_up = np.ascontiguousarray(id_up, dtype=np.int32)
        id_dn = np.ascontiguousarray(id_dn, dtype=np.int32)
        g2l = None if ghost_to_local is None else np.ascontiguousarray(ghost_to_local, dtype=np.int32)
        act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        h = C.c_void_p()
        _ck(lib().rxn_connset_create(realization.h, self.nconn, _ip(id_up), _ip(id_dn), _ip(g2l), self.nlocal,
                                     act.ctypes.data_as(C.POINTER(C.c_uint8)) if act is not None else None, C.byref(h)))
        self.h = h
        nnzb = c_i64(0)
        _ck(lib().rxn_connset_structure(self.h, C.byref(nnzb), None, None))
        self.nnz_blocks = nnzb.value

    def structure(self):
        row_ptr = np.zeros(self.nlocal + 1, dtype=np.int32)
        col = np.zeros(self.nnz_blocks, dtype=np.int32)
        _ck(lib().rxn_connset_structure(self.h, None, _ip(row_ptr), _ip(col)))
        return row_ptr, col

    def TFluxCoef(self, area, velocity, disp_over_dist, fraction_upwind=None, use_upwinding: bool = True):
        a = np.ascontiguousarray(area, dtype=np.float64)
        q = np.ascontiguousarray(velocity, dtype=np.float64)
        d = np.ascontiguousarray(disp_over_dist, dtype=np.float64)
        assert a.shape == (self.nconn,) and q.shape == (self.nconn,) and d.shape == (self.nconn, self.rz.reaction.desc.naqcomp)
        f = None if fraction_upwind is None else np.ascontiguousarray(fraction_upwind, dtype=np.float64)
        _ck(lib().rxn_connset_flux_coefs(self.h, _dp(a), _dp(q), _dp(d), _dp(f), int(use_upwinding)))

    def close(self):
        if getattr(self, 'h', None):
            lib().rxn_connset_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CouplerSet:
    """Flattened patch%boundary_condition_list (kind abi.RXN_COUPLER_BOUNDARY) or patch%source_sink_list
    (abi.RXN_COUPLER_SRC_SINK) of a realization: one-sided connections between a local cell and an external total
    (reactive_transport.F90:2347-2430, 3176-3240; :2623-2672, 3394-3436)."""

    def __init__(self, realization: 'Realization', kind: int, id_dn: np.ndarray, nlocal: int,
                 ghost_to_local: Optional[np.ndarray] = None, active: Optional[np.ndarray] = None):
        self.rz, self.kind, self.nlocal, self.nconn = realization, kind, int(nlocal), len(id_dn)
        self.h = C.c_void_p()
        dn = np.ascontiguousarray(id_dn, dtype=np.int32)
        g2l = None if ghost_to_local is None else np.ascontiguousarray(ghost_to_local, dtype=np.int32)
        act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        _ck(lib().rxn_couplerset_create(realization.h, kind, self.nconn, _ip(dn), _ip(g2l), self.nlocal,
                                        act.ctypes.data_as(C.POINTER(C.c_uint8)) if act is not None else None, C.byref(self.h)))

    def TFluxCoefBC(self, area, velocity, disp_over_dist, use_upwinding: bool = True):
        """TFluxCoef with fraction_upwind = 0.5 (reactive_transport.F90:2369-2373)."""
        _ck(lib().rxn_couplerset_bc_coefs(self.h, _dp(np.ascontiguousarray(area)), _dp(np.ascontiguousarray(velocity)),
                                          _dp(np.ascontiguousarray(disp_over_dist)), int(use_upwinding)))

    def TSrcSinkCoef(self, qsrc, tran_src_sink_type):
        """transport.F90:901-954."""
        _ck(lib().rxn_couplerset_ss_coefs(self.h, _dp(np.ascontiguousarray(qsrc, dtype=np.float64)),
                                          _ip(np.ascontiguousarray(tran_src_sink_type, dtype=np.int32))))

    def set_totals(self, total: np.ndarray):
        assert total.shape == (self.nconn, self.rz.ncomp)
        _ck(lib().rxn_couplerset_set_totals(self.h, _dp(np.ascontiguousarray(total))))

    def totals_from_state(self, bc_realization: 'Realization'):
        """rt_auxvars_bc(:)%total of the boundary realization (one cell per connection)."""
        _ck(lib().rxn_couplerset_totals_from_state(self.h, bc_realization.h))

    def close(self):
        if self.h:
            lib().rxn_couplerset_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Realization:
    """Per-rank cell state (rt_auxvars + global/material auxvars of the ghosted cells,
    reactive_transport.F90:271-274) resident in HBM, with the reference's cell loops as methods."""

    def __init__(self, reaction: Reaction, ncells_ghosted: int):
        self.reaction = reaction
        self.ncells = int(ncells_ghosted)
        self.ncomp = reaction.desc.ncomp
        self.h = C.c_void_p()
        _ck(lib().rxn_state_create(reaction.h, self.ncells, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().rxn_state_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state access (PatchGetVariable / checkpoint / CondControlAssignTranInitCond) ----
    def upload(self, field: str, a: np.ndarray):
        """a: [rows, ncells] (SoA) float64."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        assert a.shape == (self.reaction.field_rows(field), self.ncells), (field, a.shape)
        _ck(lib().rxn_state_upload(self.h, abi.F[field], _dp(a), self.ncells, 1))

    def upload_aos(self, field: str, a: np.ndarray):
        """a: [ncells, rows] (the reference's Vec layout, dof fastest)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.ncells, self.reaction.field_rows(field))
        _ck(lib().rxn_state_upload(self.h, abi.F[field], _dp(a), 1, a.shape[1]))

    def download(self, field: str) -> np.ndarray:
        rows = self.reaction.field_rows(field)
        a = np.zeros((rows, self.ncells))
        if rows:
            _ck(lib().rxn_state_download(self.h, abi.F[field], _dp(a), self.ncells, 1))
        return a

    def download_aos(self, field: str) -> np.ndarray:
        rows = self.reaction.field_rows(field)
        a = np.zeros((self.ncells, rows))
        if rows:
            _ck(lib().rxn_state_download(self.h, abi.F[field], _dp(a), 1, rows))
        return a

    def broadcast(self, field: str, row_values: np.ndarray):
        """Same value in every cell: a uniform initial condition."""
        v = np.ascontiguousarray(row_values, dtype=np.float64).ravel()
        assert v.shape[0] == self.reaction.field_rows(field)
        if v.shape[0]:
            _ck(lib().rxn_state_broadcast(self.h, abi.F[field], _dp(v)))

    def materialize(self, field: str):
        _ck(lib().rxn_state_materialize(self.h, abi.F[field]))

    def upload_host_state(self, st: abi.HostState):
        for f in abi.FIELDS:
            if f in ('DTOTAL', 'DTOTAL_SORB_EQ'):
                continue
            if st[f].shape[0]:
                self.upload(f, st[f])
        self.set_cell_scalars(active=st.active)

    def download_host_state(self, st: abi.HostState):
        for f in abi.FIELDS:
            if f in ('DTOTAL', 'DTOTAL_SORB_EQ'):
                continue
            if st[f].shape[0]:
                st[f][:] = self.download(f)

    def set_cell_scalars(self, den_kg=None, sat=None, temp=None, pres=None, volume=None, porosity=None,
                         soil_particle_density=None, active=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (den_kg, sat, temp, pres, volume, porosity, soil_particle_density)]
        act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        _ck(lib().rxn_set_cell_scalars(self.h, *[_dp(a) for a in arrs],
                                       None if act is None else act.ctypes.data_as(C.POINTER(C.c_uint8))))

    def set_react_kernel(self, which: int):
        _ck(lib().rxn_set_react_kernel(self.h, which))

    def react_kernel_info(self) -> str:
        buf = C.create_string_buffer(512)
        _ck(lib().rxn_react_kernel_info(self.h, buf, 512))
        return buf.value.decode()

    # ---- the cell loops ----
    def RTReact(self, tran_xx: np.ndarray, dt: float, dt_mode: int = abi.RXN_DT_CONSISTENT,
                l2g: Optional[np.ndarray] = None, iters: Optional[np.ndarray] = None,
                flags: Optional[np.ndarray] = None):
        """reactive_transport.F90:1605 — tran_xx [nlocal, ncomp]: in totals, out free-ion (in place).
        Returns (num_iterations[nlocal], flags[nlocal])."""
        assert tran_xx.dtype == np.float64 and tran_xx.flags.c_contiguous and tran_xx.shape[1] == self.ncomp
        n = tran_xx.shape[0]
        if iters is None:
            iters = np.zeros(n, dtype=np.int32)
        if flags is None:
            flags = np.zeros(n, dtype=np.int32)
        _ck(lib().rxn_react_batch(self.h, _dp(tran_xx), _ip(l2g), n, dt, dt_mode, _ip(iters), _ip(flags)))
        return iters, flags

    def RTUpdateAuxVars(self, xx_loc: Optional[np.ndarray], update_activity_coefs: bool):
        """reactive_transport.F90:3704 (cells part) — xx_loc [nghosted, ncomp] free-ion molalities."""
        if xx_loc is not None:
            assert xx_loc.shape == (self.ncells, self.ncomp)
        _ck(lib().rxn_update_auxvars_batch(self.h, _dp(xx_loc), int(update_activity_coefs)))

    def RTUpdateFixedAccumulation(self, xx: Optional[np.ndarray], l2g: Optional[np.ndarray] = None) -> np.ndarray:
        """reactive_transport.F90:726 — returns accum [nlocal, ncomp] in mol."""
        n = self.ncells if l2g is None else len(l2g)
        out = np.zeros((n, self.ncomp))
        _ck(lib().rxn_fixed_accum_batch(self.h, _dp(xx), _ip(l2g), n, _dp(out)))
        return out

    def RTResidualJacobianNonFlux(self, dt: float, l2g: Optional[np.ndarray] = None, residual=True, jacobian=True,
                                  res: Optional[np.ndarray] = None, jac: Optional[np.ndarray] = None):
        """Accumulation + reaction parts of RTResidualNonFlux (:2436) and RTJacobianNonFlux (:3247):
        res [nlocal, ncomp], jac [nlocal, ncomp*ncomp] (column-major blocks); `res` / `jac`: caller-owned output buffers
        (e.g. pinned), else new arrays."""
        n = self.ncells if l2g is None else len(l2g)
        if res is None:
            res = np.zeros((n, self.ncomp)) if residual else None
        if jac is None:
            jac = np.zeros((n, self.ncomp * self.ncomp)) if jacobian else None
        assert res is None or res.shape == (n, self.ncomp)
        assert jac is None or jac.shape == (n, self.ncomp * self.ncomp)
        _ck(lib().rxn_residual_jacobian_blocks_batch(self.h, _ip(l2g), n, dt, _dp(res), _dp(jac)))
        return res, jac

    def ReactionEquilibrateConstraint(self, ctype, conc, cid, free_ion_guess=None, use_prev: bool = False,
                                      molal: bool = True, l2g: Optional[np.ndarray] = None):
        """reaction.F90:1308 for every cell (condition_control.F90:725-741): conc [naq] (one constraint) or
        [nlocal, naq] (per-cell concentrations).  Returns basis_molarity [nlocal, naq], iters, status (RXN_EQ_*)."""
        n = self.ncells if l2g is None else len(l2g)
        ctype = np.ascontiguousarray(ctype, dtype=np.int32)
        cid = np.ascontiguousarray(cid, dtype=np.int32)
        conc = np.ascontiguousarray(conc, dtype=np.float64)
        stride = 0 if conc.ndim == 1 else conc.shape[1]
        guess = None if free_ion_guess is None else np.ascontiguousarray(free_ion_guess, dtype=np.float64)
        basis = np.zeros((n, self.reaction.desc.naqcomp))
        iters = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        _ck(lib().rxn_equilibrate_constraint_batch(self.h, _ip(ctype), _dp(conc), stride, _ip(cid), _dp(guess), int(use_prev),
                                                   int(molal), _ip(l2g), n, _dp(basis), _ip(iters), _ip(status)))
        return basis, iters, status

    def RTUpdateKineticState(self, dt: float):
        """reactive_transport.F90:642."""
        _ck(lib().rxn_update_kinetic_state_batch(self.h, dt))

    # ---- device-resident variants and measurement helpers (bench.py) ----
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _ck(lib().rxn_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int):
        _ck(lib().rxn_device_free(self.h, C.c_void_p(ptr)))

    def device_copy(self, dst, src, nbytes: int, kind: int):
        """kind 0: host->device, 1: device->host, 2: device->device; dst/src are ints or numpy arrays."""
        d = dst.ctypes.data if isinstance(dst, np.ndarray) else dst
        s_ = src.ctypes.data if isinstance(src, np.ndarray) else src
        _ck(lib().rxn_device_copy(self.h, C.c_void_p(d), C.c_void_p(s_), nbytes, kind))

    def RTUpdateAuxVars_device(self, d_xx_loc: int, update_activity_coefs: bool):
        _ck(lib().rxn_update_auxvars_batch_device(self.h, C.c_void_p(d_xx_loc), int(update_activity_coefs)))

    def RTResidualJacobianNonFlux_device(self, nlocal: int, dt: float, d_res: int = 0, d_jac: int = 0, d_l2g: int = 0):
        _ck(lib().rxn_residual_jacobian_blocks_batch_device(self.h, C.c_void_p(d_l2g or None), nlocal, dt,
                                                            C.c_void_p(d_res or None), C.c_void_p(d_jac or None)))

    def RTResidualFlux(self, conn: 'ConnectionSet') -> np.ndarray:
        """Interior-flux part of RTResidualFlux: r_p [nlocal, ncomp]."""
        r = np.zeros((conn.nlocal, self.ncomp))
        _ck(lib().rxn_flux_residual_batch(self.h, conn.h, _dp(r)))
        return r

    def RTJacobianFlux(self, conn: 'ConnectionSet') -> np.ndarray:
        """Interior-flux part of RTJacobianFlux: block values [nnz_blocks, ncomp*ncomp] in conn.structure()."""
        n = self.ncomp
        val = np.zeros((conn.nnz_blocks, n * n))
        _ck(lib().rxn_flux_jacobian_batch(self.h, conn.h, _dp(val)))
        return val

    def RTResidualCoupler(self, cs: 'CouplerSet', res: np.ndarray, want_flux: bool = False):
        """Boundary part of RTResidualFlux (:2347-2430) / source-sink part of RTResidualNonFlux (:2623-2672): res [nlocal, ncomp]
        is updated in place; returns patch%boundary_tran_fluxes / patch%ss_tran_fluxes [nconn, ncomp] when asked."""
        assert res.shape == (cs.nlocal, self.ncomp) and res.flags.c_contiguous
        flux = np.zeros((cs.nconn, self.ncomp)) if want_flux else None
        _ck(lib().rxn_coupler_residual_batch(self.h, cs.h, _dp(res), _dp(flux)))
        return flux

    def RTJacobianCoupler(self, cs: 'CouplerSet', val: np.ndarray, conn: Optional['ConnectionSet'] = None):
        """Boundary part of RTJacobianFlux (:3176-3240) / source-sink part of RTJacobianNonFlux (:3394-3436), added into the
        diagonal blocks of `val`: the block-CSR values of `conn`, or [nlocal, ncomp*ncomp] diagonal blocks without it."""
        assert val.flags.c_contiguous and val.shape == ((conn.nnz_blocks if conn is not None else cs.nlocal), self.ncomp * self.ncomp)
        _ck(lib().rxn_coupler_jacobian_batch(self.h, conn.h if conn is not None else None, cs.h, _dp(val)))

    def boundary_free_ion(self, bc_type, basis_molarity, den_kg_bc, boundary_velocity, xx_loc_cells):
        """xxbc of every boundary connection as RTUpdateAuxVars builds it (reactive_transport.F90:3935-3990, liquid phase, no
        colloids): DIRICHLET (1) / CONCENTRATION_SS / NEUMANN -> basis_molarity / den_kg * 1000; ZERO_GRADIENT (4) -> the
        cell's free-ion molalities; DIRICHLET_ZERO_GRADIENT (3) -> Dirichlet where the boundary velocity is >= 0 (inflow), else
        zero gradient.  Pure data movement for the call that follows it: bc.RTUpdateAuxVars(xxbc, ...) on the boundary realization."""
        bc_type = np.asarray(bc_type)
        dirichlet = basis_molarity / np.asarray(den_kg_bc)[:, None] * 1000.0
        zero_grad = (bc_type == 4) | ((bc_type == 3) & ~(np.asarray(boundary_velocity) >= 0.0))
        return np.ascontiguousarray(np.where(zero_grad[:, None], xx_loc_cells, dirichlet))

    def RTResidualFlux_device(self, conn: 'ConnectionSet', d_res: int):
        _ck(lib().rxn_flux_residual_batch_device(self.h, conn.h, C.c_void_p(d_res)))

    def RTJacobianFlux_device(self, conn: 'ConnectionSet', d_val: int):
        _ck(lib().rxn_flux_jacobian_batch_device(self.h, conn.h, C.c_void_p(d_val)))

    def RTReact_device(self, d_xx: int, nlocal: int, dt: float, dt_mode: int = abi.RXN_DT_CONSISTENT,
                       d_l2g: int = 0, d_iters: int = 0, d_flags: int = 0):
        """RTReact with tran_xx / iters / flags already resident in this GPU's HBM."""
        _ck(lib().rxn_react_batch_device(self.h, C.c_void_p(d_xx), C.c_void_p(d_l2g or None), nlocal, dt, dt_mode,
                                         C.c_void_p(d_iters or None), C.c_void_p(d_flags or None)))

    def timer_start(self):
        _ck(lib().rxn_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _ck(lib().rxn_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def probe_fp64_tflops(self) -> float:
        v = C.c_double()
        _ck(lib().rxn_probe_fp64(self.h, C.byref(v)))
        return float(v.value)

    def last_kernel_ms(self) -> float:
        return float(lib().rxn_last_kernel_ms(self.h))


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array over page-locked host memory (cudaMallocHost); never freed (process lifetime)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _ck(lib().rxn_host_alloc(max(n, 8), C.byref(p)))
    buf = (C.c_char * max(n, 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def launch_count() -> int:
    return int(lib().rxn_launch_count())
