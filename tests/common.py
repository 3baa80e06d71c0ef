"""Shared helpers of the parity tests: run one entry point through two backends (oracle /
emulated device code / CUDA library) on the same seeded synthetic cells and compare."""
import numpy as np

from pflotran_b200 import abi, synth

STATE_FIELDS = ['PRI_MOLAL', 'TOTAL', 'SEC_MOLAL', 'PRI_ACT_COEF', 'SEC_ACT_COEF', 'LN_ACT_H2O', 'TOTAL_SORB_EQ',
                'FREE_SITE_CONC', 'EQSRFCPLX_CONC', 'KINMR_TOTAL_SORB', 'EQIONX_REF_CATION_SORBED_CONC', 'EQIONX_CONC',
                'MNRL_VOLFRAC', 'MNRL_RATE', 'KINSRFCPLX_CONC', 'KINSRFCPLX_CONC_KP1', 'KINSRFCPLX_FREE_SITE_CONC']

# north_star: relative 1e-10 on converged free-ion and mineral concentrations, identical flags
RTOL = 1.0e-10


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def total_magnitude(st, tables, cells=None):
    """m_i + sum_k |nu_ik| sec_molal_k, times den_kg*1e-3: the scale on which total_i (reaction.F90:4095-4124) carries
    a relative perturbation of its terms when the stoichiometries change sign (H+).  [naq, ncells]"""
    pick = (lambda a: a) if cells is None else (lambda a: a[:, cells])
    mag = np.abs(pick(st['PRI_MOLAL'])).copy()
    if tables.neqcplx:
        ids, stc = np.asarray(tables.eqcplxspecid), np.asarray(tables.eqcplxstoich)
        sm = pick(st['SEC_MOLAL'])
        for k in range(tables.neqcplx):
            for q in range(1, ids[k, 0] + 1):
                mag[ids[k, q] - 1] += abs(stc[k, q]) * np.abs(sm[k])
    return mag * pick(st['DEN_KG']) * 1.0e-3


def residual_scale(st, tables, r_o, a_o, dt):
    """Scale of the global-implicit residual block: res = accumulation/dt + sum_m nu_im Im_m V (reaction.F90:5072-5148,
    reaction_mineral.F90:795-830).  Near equilibrium accumulation/dt and the mineral terms cancel, and Im = -area k (1-QK)
    itself is a difference: a relative perturbation eps of the inputs moves the residual by
    eps * (|accumulation/dt| + sum_m |nu_im| area_m |k_m| V).  [ncells, ncomp]"""
    sc = np.maximum(np.abs(r_o), np.abs(a_o) / dt)
    if tables.nkinmnrl:
        ids, stc = np.asarray(tables.kinmnrlspecid), np.asarray(tables.kinmnrlstoich)
        kk = np.abs(np.asarray(tables.kinmnrl_rate_constant))
        mn = np.zeros_like(sc)
        for m in range(tables.nkinmnrl):
            for q in range(1, ids[m, 0] + 1):
                mn[:, ids[m, q] - 1] += abs(stc[m, q - 1 if stc.shape[1] < ids.shape[1] else q]) * st['MNRL_AREA'][m] * kk[m] * st['VOLUME'][0]
        sc = np.maximum(sc, mn)
    return np.maximum(sc, 1e-12 * np.abs(r_o).max(axis=1, keepdims=True))


def jacobian_scale(st, j_o, ncomp):
    """Scale of the entries of a Jacobian block (column-major, d/dm_j): the Newton solve works on the block with column j
    times m_j and every row divided by its maximum (RSolve, reaction.F90:4851-4870), so an entry matters relative to
    the largest |J_ij'| m_j' of its row: scale_ij = max(|J_ij|, max_j'(|J_ij'| m_j') / m_j).  Entries far below that are
    differences of the accumulation and kinetic derivative terms (they cancel near equilibrium).  [ncells, ncomp^2]"""
    n = j_o.shape[0]
    J = np.abs(j_o).reshape(n, ncomp, ncomp).transpose(0, 2, 1)          # [cell, i, j]
    m = np.abs(st['PRI_MOLAL']).T[:, None, :]                             # [cell, 1, j]
    rowmax = (J * m).max(axis=2, keepdims=True)
    sc = np.maximum(J, rowmax / np.maximum(m, 1e-300))
    return sc.transpose(0, 2, 1).reshape(n, ncomp * ncomp)


def assert_state_close(st_a, st_b, rtol=RTOL, fields=STATE_FIELDS, cells=None, what='', tables=None, kinetic_dt=None):
    for f in fields:
        a, b = st_a[f], st_b[f]
        if not a.size:
            continue
        if cells is not None:
            a, b = a[:, cells], b[:, cells]
        if f in ('PRI_MOLAL', 'MNRL_VOLFRAC'):
            # the north-star quantities: pure relative comparison, no floor
            scale = np.abs(b)
        else:
            # derived values far below the row's scale are differences of O(1) sums: compare them on that scale
            scale = np.maximum(np.abs(b), 1e-13 * np.max(np.abs(b), axis=1, keepdims=True))
        if f == 'TOTAL' and tables is not None and tables.neqcplx:
            # total_i = m_i + sum_k nu_ik sec_molal_k (reaction.F90:4095-4124) cancels when nu changes sign (H+):
            # a relative perturbation eps of the terms moves it by eps * (m_i + sum_k |nu_ik| sec_molal_k)
            ids, st = np.asarray(tables.eqcplxspecid), np.asarray(tables.eqcplxstoich)
            sm = st_b['SEC_MOLAL'] if cells is None else st_b['SEC_MOLAL'][:, cells]
            mag = np.abs(st_b['PRI_MOLAL'] if cells is None else st_b['PRI_MOLAL'][:, cells]).copy()
            for k in range(tables.neqcplx):
                for q in range(1, ids[k, 0] + 1):
                    mag[ids[k, q] - 1] += abs(st[k, q]) * np.abs(sm[k])
            den = (st_b['DEN_KG'] if cells is None else st_b['DEN_KG'][:, cells]) * 1.0e-3
            scale = np.maximum(scale, mag * den)
        if f == 'MNRL_RATE' and tables is not None:
            # rate = -area*k*(1 - QK) (reaction_mineral.F90:795-816): near equilibrium 1 - QK cancels, so a
            # relative perturbation eps of the molalities moves the rate by ~eps*area*k*QK*sum|nu|, not eps*|rate|
            area = st_b['MNRL_AREA'] if cells is None else st_b['MNRL_AREA'][:, cells]
            scale = np.maximum(scale, area * np.abs(np.asarray(tables.kinmnrl_rate_constant))[:, None])
        if f == 'MNRL_VOLFRAC' and tables is not None and kinetic_dt is not None:
            # after RUpdateKineticState: volfrac += rate*molar_vol*dt (reaction.F90:5354-5364) with rate on the scale area*k
            area = st_b['MNRL_AREA'] if cells is None else st_b['MNRL_AREA'][:, cells]
            scale = np.maximum(scale, area * (np.abs(np.asarray(tables.kinmnrl_rate_constant)) *
                                              np.abs(np.asarray(tables.kinmnrl_molar_vol)))[:, None] * kinetic_dt)
        err = np.abs(a - b) / np.maximum(scale, 1e-300)
        bad = ~(err <= rtol) & ~((a == b) | (np.isnan(a) & np.isnan(b)))
        assert not bad.any(), '%s field %s: max rel err %.3e at %s' % (what, f, np.nanmax(err), np.argwhere(bad)[:3])


def workload_cells(name, n, start=0, **kw):
    w = synth.Workload(name)
    cells = synth.make_cells(w, start, n, **kw)
    return w, cells
