"""Replays the reference's GLOBAL_IMPLICIT transport time step for batch (flux-free) decks so that the
per-cell entry points can be pinned to the reference's time-stepped .regression.gold files
(regression_tests/ascem/batch/calcite-kinetics*.gold, regression_tests/default/batch/solute_KD_*.gold).

Test infrastructure.  Sequence restated (all citations /root/reference/src/pflotran):
  per time step (timestepper_BE StepDT -> pm_rt.F90):
    PMRTInitializeTimestep :361   RTUpdateFixedAccumulation (reactive_transport.F90:726), then, for
                                  ACTIVITY_COEFFICIENTS TIMESTEP, RTUpdateActivityCoefficients (:3620)
    PMRTPreSolve :419             x = ln(tran_xx) for LOG_FORMULATION
    SNESSolve (newtonls, basic line search: pmc_subsurface.F90:365)
      F(x): RTResidual :2057      xx = exp(x); RTUpdateAuxVars (:3704; + RActivityCoefficients for
                                  NEWTON_ITERATION); r = -accum/dt + RTAccumulation/dt + RReaction (:2545-2586, 2735-2758)
      J(x): RTJacobian :2903      accumulation + reaction derivative blocks (:3342-3389, 3445-3465), columns
                                  times xx for LOG_FORMULATION (MatDiagonalScaleLocal :2980)
      step                        Y = J^-1 F; PMRTCheckUpdatePre :734 (|d ln C| <= max_dlnC, or the
                                  min-ratio scaling of the linear formulation); x <- x - Y
      convergence                 SNESConvergedDefault (convergence.F90:172) with PETSc's defaults
                                  atol 1e-50, rtol 1e-8, stol 1e-8, max 50 iterations; at least
                                  newton_min_iterations = 1 iteration (solver.F90:144, convergence.F90:246,316)
    PMRTUpdateSolution :991       RTUpdateEquilibriumState (:552: RTUpdateAuxVars, no activity update),
                                  RTUpdateKineticState (:642)
The cells are independent (no flux), so a batch of identical or different cells steps together; the
SNES norms are taken PER CELL (each cell is its own 1-cell simulation as in the reference decks).
The linear solve of the block system is a dense LU (numpy): the reference's BCGS+ILU(0) on a single dense
block is exact after one iteration ("Solver Iterations" = "Newton Iterations" in the gold files).
"""
import numpy as np

from pflotran_b200 import abi


class OracleGI:
    """global-implicit entry points of the CPU oracle on a HostState"""

    def __init__(self, t, st):
        from oracle.pyoracle import Oracle
        self.orc = Oracle(t)
        self.st = st

    def fixed_accum(self, xx):
        return self.orc.fixed_accum(self.st, xx)

    def update_auxvars(self, xx, act):
        self.orc.update_auxvars(self.st, xx, act)

    def residual_jacobian(self, dt):
        return self.orc.residual_jacobian(self.st, dt)

    def update_kinetic_state(self, dt):
        self.orc.update_kinetic_state(self.st, dt)

    def state(self):
        return self.st


class DeviceGI:
    """the same calls through the C ABI (pflotran_b200.reactive_transport.Realization)"""

    def __init__(self, rz, st):
        self.rz = rz
        self.st = st
        rz.upload_host_state(st)

    def fixed_accum(self, xx):
        return self.rz.RTUpdateFixedAccumulation(xx)

    def update_auxvars(self, xx, act):
        self.rz.RTUpdateAuxVars(xx, act)

    def residual_jacobian(self, dt):
        return self.rz.RTResidualJacobianNonFlux(dt)

    def update_kinetic_state(self, dt):
        self.rz.RTUpdateKineticState(dt)

    def state(self):
        self.rz.download_host_state(self.st)
        return self.st


TFAC = (2.0, 2.0, 2.0, 2.0, 2.0, 1.8, 1.6, 1.4, 1.2, 1.0, 1.0, 1.0, 1.0)    # timestepper_BE.F90:125-133


def run_deck(t, be, xx, final_time, dt0, dt_max, iaccel=5, tolerance=0.1, **kw):
    """Time loop of the reference for one simulation: target time with the final time as the only waypoint
    (timestepper_base.F90:400-452, time_step_tolerance 0.1 :153) and the step-size controller of
    PMRTUpdateTimestep (pm_rt.F90:585-652, "original implementation": volfrac_change_governor = 1).
    Returns (number of time steps, Newton iterations summed over the steps) - the two counters the
    reference prints in the `-- SOLUTION: Transport --` block of a .regression.gold file."""
    time, dt, steps, newton = 0.0, dt0, 0, 0
    while time < final_time:
        dt = min(dt, dt_max)
        target = time + dt
        if target + tolerance * dt >= final_time:
            d = final_time - time
            if d > dt_max and abs(d - dt_max) > 1.0:
                target = time + dt_max
                d = dt_max
            else:
                target = final_time
            dt = d
        its = run(t, be, xx, [dt], **kw)
        assert (its == its[0]).all(), 'run_deck drives one simulation: all cells must take the same path'
        n = int(its[0])
        time = target
        steps += 1
        newton += n
        if iaccel != 0:                        # TimestepperBEUpdateDT (timestepper_BE.F90:238): TS_ACCELERATION 0 keeps dt
            if n <= iaccel:
                dtt = TFAC[n - 1] * dt if n <= len(TFAC) else 0.5 * dt
            else:
                dtt = 0.5 * dt
            dtt = min(dtt, 2.0 * dt, dt_max)
            dt = dtt
    return steps, newton


def run(t, be, xx, dts, atol=1.0e-50, rtol=1.0e-8, stol=1.0e-8, maxit=50):
    """Advances every cell of the backend through the time steps `dts`.  xx [ncells, ncomp]: free-ion
    molalities (the solution vector), updated in place.  Returns the Newton iteration count per cell."""
    ncells, n = xx.shape
    use_log = bool(t.use_log_formulation)
    act_ts = t.act_coef_update_frequency == 1   # RXN_ACT_COEF_FREQUENCY_TIMESTEP (include/rxn_b200.h:60)
    act_ni = t.act_coef_update_frequency == 2   # RXN_ACT_COEF_FREQUENCY_NEWTON_ITER
    newton = np.zeros(ncells, dtype=np.int64)

    def residual(dt, fixed, first):
        be.update_auxvars(xx, act_ni or (first and act_ts))
        res, jac = be.residual_jacobian(dt)
        return res - fixed / dt, jac.reshape(ncells, n, n).transpose(0, 2, 1)   # blocks are column-major

    for dt in dts:
        fixed = be.fixed_accum(xx)
        # RTUpdateActivityCoefficients + the RTUpdateAuxVars of the first residual see the same pri_molal:
        # one update_auxvars(xx, True) is both
        F, J = residual(dt, fixed, True)
        x = np.log(xx) if use_log else xx.copy()
        fnorm = np.linalg.norm(F, axis=1)
        ttol = rtol * fnorm
        active = np.ones(ncells, dtype=bool)      # iteration 0 never converges: newton_min_iterations = 1
        for it in range(maxit):
            if not active.any():
                break
            if use_log:
                J = J * xx[:, None, :]
            Y = np.zeros_like(F)
            for c in np.nonzero(active)[0]:
                Y[c] = np.linalg.solve(J[c], F[c])
            if use_log:
                Y = np.sign(Y) * np.minimum(np.abs(Y), t.max_dlnC)
            else:
                for c in np.nonzero(active)[0]:
                    m = (x[c] <= Y[c])
                    if m.any():
                        ratio = np.abs(x[c][m] / Y[c][m]).min()
                        if ratio < 1.0:
                            Y[c] *= ratio * 0.99
            x = np.where(active[:, None], x - Y, x)
            xx[:] = np.exp(x) if use_log else x
            newton += active
            F, J = residual(dt, fixed, False)
            fnorm = np.linalg.norm(F, axis=1)
            xnorm = np.linalg.norm(x, axis=1)
            ynorm = np.linalg.norm(Y, axis=1)
            conv = (fnorm < atol) | (fnorm <= ttol) | (ynorm < stol * xnorm)
            active = active & ~conv
        assert not active.any(), 'SNES did not converge within %d iterations' % maxit
        be.update_auxvars(xx, False)           # RTUpdateEquilibriumState
        be.update_kinetic_state(dt)            # RTUpdateKineticState
    return newton


TIME_STEPPED_GOLD = ['calcite_kinetics', 'calcite_kinetics_vf', 'kd_w_mineral', 'kd_wo_mineral', 'general_reaction',
                     'abcd_microbial', 'abcd_microbial_act_high', 'abcd_microbial_act_low']


def check_time_stepped_gold(w, be, t, xx, tol=1.0e-12):
    """Runs the deck of fixture `w` on backend `be` and asserts every value of the reference's
    .regression.gold file: printed variables within `tol` (the reference's own criterion, batch.cfg:12-21:
    1e-12 absolute; relative above 1), time-step and Newton-iteration counters equal, solution 2-norm."""
    import kat
    tm = w.meta['time']
    kw = {k.lower(): tm[k] for k in ('ATOL', 'RTOL', 'STOL') if k in tm}          # NEWTON_SOLVER card of the deck
    steps, newton = run_deck(t, be, xx, tm['FINAL_TIME'], tm['INITIAL_TIMESTEP_SIZE'], tm['MAXIMUM_TIMESTEP_SIZE'], tm['iaccel'], **kw)
    out = kat.outputs(t, be.state())
    gold = w.gold
    assert steps == gold['Transport']['Time Steps'] and newton == gold['Transport']['Newton Iterations'], (steps, newton)
    g2 = gold['Transport']['Solution 2-Norm']
    assert abs(np.linalg.norm(xx[0]) - g2) <= 1.0e-12 * g2          # cell 0 is the reference's single cell
    checked = 0
    for var, vals in gold.items():
        if var in ('Transport', 'Material ID'):
            continue
        g = vals['1']
        assert abs(out[var] - g) <= tol * max(1.0, abs(g)), '%s %s: %.14e gold %.14e' % (w.name, var, out[var], g)
        checked += 1
    return checked


def check_decay_closed_form(w, be, t, xx, rtol=1.0e-12):
    """Radioactive decay A -> B stepped by backward Euler has the closed form A_n = A_0 / (1 + k dt)^n, B_n = B_0 + A_0 - A_n
    (linear problem: one Newton iteration per step).  Fixture `decay_ab`: the first-order reaction of the reference's
    general-reaction.in written as RADIOACTIVE_DECAY_REACTION with the same rate (half life ln 2 / k)."""
    import kat
    tm = w.meta['time']
    kw = {k.lower(): tm[k] for k in ('ATOL', 'RTOL', 'STOL') if k in tm}
    a0, b0 = float(xx[0, 0]), float(xx[0, 1])
    steps, newton = run_deck(t, be, xx, tm['FINAL_TIME'], tm['INITIAL_TIMESTEP_SIZE'], tm['MAXIMUM_TIMESTEP_SIZE'], tm['iaccel'], **kw)
    assert steps == 500 and newton == 500
    k = float(t.radiodecay_kf[0])
    an = a0 / (1.0 + k * tm['INITIAL_TIMESTEP_SIZE']) ** steps
    st = be.state()
    assert abs(st['PRI_MOLAL'][0, 0] - an) <= rtol * an, (st['PRI_MOLAL'][0, 0], an)
    assert abs(st['PRI_MOLAL'][1, 0] - (b0 + a0 - an)) <= rtol * (b0 + a0), (st['PRI_MOLAL'][1, 0], b0 + a0 - an)
