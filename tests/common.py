"""Shared helpers of the parity tests: run one entry point through two backends (oracle /
emulated device code / CUDA library) on the same seeded synthetic cells and compare."""
import numpy as np

from pflotran_b200 import abi, synth

STATE_FIELDS = ['PRI_MOLAL', 'TOTAL', 'SEC_MOLAL', 'PRI_ACT_COEF', 'SEC_ACT_COEF', 'LN_ACT_H2O', 'TOTAL_SORB_EQ',
                'FREE_SITE_CONC', 'EQSRFCPLX_CONC', 'KINMR_TOTAL_SORB', 'EQIONX_REF_CATION_SORBED_CONC', 'EQIONX_CONC',
                'MNRL_VOLFRAC', 'MNRL_RATE', 'KINSRFCPLX_CONC', 'KINSRFCPLX_CONC_KP1', 'KINSRFCPLX_FREE_SITE_CONC', 'IMMOBILE']

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


def accumulation_scale(st, tables, a_o):
    """Scale of the fixed accumulation phi*s*1000*V*total (+ sorbed*V), reaction.F90:5072-5148: total's own terms for the
    aqueous dofs, the value itself for the immobile dofs (immobile*V, no cancellation).  [ncells, ncomp]"""
    aq = (st['POROSITY'] * st['SAT'] * 1000.0 * st['VOLUME'] * total_magnitude(st, tables)).T
    sc = np.abs(a_o).copy()
    sc[:, :aq.shape[1]] = np.maximum(sc[:, :aq.shape[1]], aq)
    return sc


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
    m = np.abs(np.concatenate([st['PRI_MOLAL'], st['IMMOBILE']], axis=0)).T[:, None, :]   # [cell, 1, j]: aqueous, then immobile dofs
    rowmax = (J * m).max(axis=2, keepdims=True)
    sc = np.maximum(J, rowmax / np.maximum(m, 1e-300))
    return sc.transpose(0, 2, 1).reshape(n, ncomp * ncomp)


class PerturbedOracle:
    """The oracle's RReact on the same cells with tran_xx changed in the last bit (relative 2.2e-16, random sign), run once
    and only when a comparison needs it: how much the REFERENCE algorithm itself moves - in values, iteration counts and
    flags - under a rounding-level change of its input.  That measured response is the yardstick for the few cells where a
    chemistry is too ill conditioned for "1e-10 and identical iteration counts" to be a property of the algorithm."""

    def __init__(self, w, cells, dt, mode, nthreads=8):
        self.args = (w, cells, dt, mode, nthreads)
        self.res = None

    def run(self):
        if self.res is None:
            from oracle.pyoracle import Oracle
            w, cells, dt, mode, nthreads = self.args
            st_p = synth.host_state(w, cells)
            sign = np.where(np.random.default_rng(11).random(cells['tran_xx'].shape) < 0.5, -1.0, 1.0)
            xp = cells['tran_xx'] * (1.0 + 2.2e-16 * sign)
            it_p, fl_p = Oracle(w.tables).react(st_p, xp, dt, mode, maxit=10000, nthreads=nthreads)
            self.res = (xp, it_p, fl_p)
        return self.res


def iteration_parity(it_a, fl_a, it_o, fl_o, perturbed, max_fraction=0.01, factor=3):
    """north star: identical Newton iteration counts and convergence flags.  Returns the boolean mask of the cells where
    they are identical.  Any mismatch is an error UNLESS the reference algorithm itself is shown not to have a stable count
    there: the oracle re-run on last-bit-perturbed inputs (PerturbedOracle) must change its own counts / flags in a comparable
    number of cells (mismatches <= factor x that number, and <= max_fraction of the batch).  Measured: 300A, calcite, hpt,
    ion exchange, surface complexation, prefactor chemistries: 0 cells, i.e. strict equality; the 22-primary ascem redox
    chemistry (cells that take up to 2688 damped iterations): the oracle flips 17 of 3000 cells, FMA contraction 16."""
    # a cell whose arithmetic left the finite range (RXN_FLAG_NONFINITE in both; the reference itself would spin or abort there)
    # is compared on the iteration count and on that flag: whether the closing RTotalSorb then also runs its free-site loop
    # into the 100000-iteration guard (RXN_FLAG_CAPPED) depends on which garbage value the singular Newton step produced
    # (measured: cell 2415 of hanford300a_stoich, |update| differs 10x between summation orders at iteration 3)
    nf = ((fl_a & abi.RXN_FLAG_NONFINITE) != 0) & ((fl_o & abi.RXN_FLAG_NONFINITE) != 0)
    same = (it_a == it_o) & ((fl_a == fl_o) | (nf & (((fl_a ^ fl_o) & ~abi.RXN_FLAG_CAPPED) == 0)))
    if same.all():
        return same
    _, it_p, fl_p = perturbed.run()
    n_ref = int(((it_p != it_o) | (fl_p != fl_o)).sum())
    n_bad = int((~same).sum())
    assert n_bad <= factor * n_ref and n_bad <= max_fraction * len(same), (
        'iteration counts / flags differ in %d cells; the oracle itself changes %d under a last-bit input change' % (n_bad, n_ref))
    return same


def free_ion_parity(xa, xo, ok, perturbed, rtol=RTOL, max_fraction=0.01, amplification=10.0):
    """The north-star comparison: converged free-ion concentrations of backend `xa` against the oracle's `xo`, PURE relative
    `rtol`, over the cells `ok` (boolean).  Returns the indices of the cells that meet it.

    A cell may miss it only if the problem itself is ill conditioned there, which is MEASURED, not assumed: `perturbed`
    (PerturbedOracle) runs the oracle again on inputs changed in the last bit and the response of the oracle's own answer
    is the yardstick - a deviation is accepted when it is below `amplification` x the largest such response in the batch
    (measured: the oracle moves by 3.6e-9 there, the FMA-contracted device arithmetic deviates by 4.4e-9), such cells are at
    most `max_fraction` of the batch, and never above 1e-6.  In the fixtures this concerns only the uraninite / O2(aq) redox couple of
    regression_tests/default/column/mineral_prefactor.in (O2(aq) ~ 1e-66 molal, UO2++ ~ 1e-21: about 10 of 5000 cells)."""
    idx = np.where(ok)[0]
    err = rel_err(xa[idx], xo[idx])
    good = (err <= rtol).all(axis=1)
    if good.all():
        return idx
    xp = perturbed.run()[0]
    sens = rel_err(xp[idx], xo[idx])
    bad = ~good
    assert bad.mean() <= max_fraction, 'free-ion parity: %.2f %% of the cells miss %.0e' % (100 * bad.mean(), rtol)
    lim = min(max(rtol, amplification * float(sens.max())), 1.0e-6)
    assert (err[bad] <= lim).all(), 'free-ion parity: max rel err %.3e not explained by the conditioning (oracle response %.3e)' % (
        err[bad].max(), sens[bad].max())
    return idx[good]


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
                # ... and sec_molal_k itself carries eps * sum_j |nu_kj| (see SEC_MOLAL below)
                amp_k = max(1.0, sum(abs(st[k, q]) for q in range(1, ids[k, 0] + 1)))
                for q in range(1, ids[k, 0] + 1):
                    mag[ids[k, q] - 1] += abs(st[k, q]) * np.abs(sm[k]) * amp_k
            den = (st_b['DEN_KG'] if cells is None else st_b['DEN_KG'][:, cells]) * 1.0e-3
            scale = np.maximum(scale, mag * den)
        if f == 'SEC_MOLAL' and tables is not None and tables.neqcplx:
            # sec_molal_k = exp(sum_j nu_kj ln a_j - ln K) (reaction.F90:4104-4122): a relative deviation eps of the free-ion
            # concentrations (the quantity held to 1e-10) becomes eps * sum_j |nu_kj| in the complex (U++++: |nu| sums to 5.5)
            ids, stc = np.asarray(tables.eqcplxspecid), np.asarray(tables.eqcplxstoich)
            amp = np.array([max(1.0, sum(abs(stc[k, q]) for q in range(1, ids[k, 0] + 1))) for k in range(tables.neqcplx)])
            scale = scale * amp[:, None]
        if f == 'MNRL_RATE' and tables is not None:
            # rate = -area*k*(1 - QK) (reaction_mineral.F90:795-816): near equilibrium 1 - QK cancels, so a
            # relative perturbation eps of the molalities moves the rate by ~eps*area*k*QK*sum|nu|, not eps*|rate|;
            # with prefactors k is sum_p k_p prod a^alpha (:743-782), bounded here by the largest k_p
            area = st_b['MNRL_AREA'] if cells is None else st_b['MNRL_AREA'][:, cells]
            keff = np.abs(np.asarray(tables.kinmnrl_rate_constant))
            if getattr(tables, 'max_num_prefactors', 0) > 0:
                keff = np.maximum(keff, np.abs(np.asarray(tables.kinmnrl_pref_rate)).max(axis=1))
            scale = np.maximum(scale, area * keff[:, None])
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
