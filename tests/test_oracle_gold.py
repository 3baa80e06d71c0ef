"""Pins the CPU oracle (oracle/rxn_oracle.cpp) to the reference's own 14-digit regression
gold files.  Tolerance is the reference's own: 1e-12 absolute
(regression_tests/ascem/batch/batch.cfg:12-21), relaxed to 1e-12 relative for values > 1.

The decks with `MAX_STEPS -1` only equilibrate the initial constraint, so every printed value
(pH, totals, activity coefficients, sorbed concentrations, free sites) is reproduced by
ReactionEquilibrateConstraint + RTUpdateAuxVars.  The fixtures (tests/golden/*.json) carry the
tables, the constraint and the gold values; no file of /root/reference is read here."""
import numpy as np
import pytest

import kat
from pflotran_b200 import synth

EQUILIBRATE_ONLY = ['carbonate_unit', 'carbonate_dh', 'ca_carbonate_unit', 'ca_carbonate_dh', 'ion_exchange',
                    'surface_complexation']


def _check(name):
    w = synth.Workload(name)
    t, orc, st, xx, nit, cst = kat.initial_cell_from_fixture(w)
    out = kat.outputs(t, st)
    gold = w.gold
    assert gold['Transport']['Time Steps'] == 0.0
    checked = 0
    for var, vals in gold.items():
        if var in ('Transport', 'Material ID') or var.endswith('Site Density'):
            continue
        assert var in out, (var, sorted(out))
        g = vals['1']
        tol = 1.0e-12 * max(1.0, abs(g))
        assert abs(out[var] - g) <= tol, '%s %s: oracle %.14e gold %.14e' % (name, var, out[var], g)
        checked += 1
    assert checked >= 4
    return out


@pytest.mark.parametrize('name', EQUILIBRATE_ONLY)
def test_gold_initial_speciation(name):
    _check(name)


@pytest.mark.parametrize('name', ['calcite_kinetics', 'calcite_kinetics_vf', 'kd_w_mineral', 'kd_wo_mineral', 'general_reaction',
                                  'abcd_microbial', 'abcd_microbial_act_high', 'abcd_microbial_act_low'])
def test_gold_time_stepped(name):
    """Kinetic side of the oracle (RTAccumulation, RKineticMineral, RTotalSorbKD, RUpdateKineticState and the
    accumulation/reaction Jacobian blocks) pinned to the reference's time-stepped gold files through the 1-cell
    global-implicit loop (tests/gi_driver.py): every printed value at the reference's own 1e-12, and the same number
    of time steps and Newton iterations as the reference's SNES took (500/1000, 62/164, 2/2, 2/2).
    abcd_microbial*: RMicrobial (Monod, inverse-Monod inhibition, biomass, Arrhenius factor), the immobile dof of the
    accumulation and RImmobileDecay - regression_tests/default/batch/ABCD_microbial*.regression.gold, 111 steps / 222
    Newton iterations under the reference's step-size controller."""
    import gi_driver
    w = synth.Workload(name)
    t, orc, st, xx, nit, cst = kat.initial_cell_from_fixture(w)
    assert gi_driver.check_time_stepped_gold(w, gi_driver.OracleGI(t, st), t, xx) >= 1


def test_radioactive_decay_closed_form():
    """RRadioactiveDecay through the 1-cell global-implicit loop: 500 backward-Euler steps hit A_0 / (1 + k dt)^500 at 1e-12."""
    import gi_driver
    w = synth.Workload('decay_ab')
    t, orc, st, xx, nit, cst = kat.initial_cell_from_fixture(w)
    gi_driver.check_decay_closed_form(w, gi_driver.OracleGI(t, st), t, xx)


def test_ascem_22_primaries_164_complexes_kat():
    """BASELINE config 1: the 22-primary / 164-complex chemistry of example_problems/ascem_chemistry equilibrated by the
    oracle reproduces the reference's own printed speciation (pflotran.out:5811-5870): `iterations: 179` exactly and all
    201 printed molalities to the 5 figures printed."""
    w = synth.Workload('ascem')
    t, orc, st, xx, nit, cst = kat.initial_cell_from_fixture(w)
    assert (t.naqcomp, t.neqcplx) == (22, 164)
    assert kat.check_speciation_kat(w, t, cst, nit) == 2 * 22 + 157


def test_gold_values_spot():
    """The three numbers SURVEY.md 8c quotes explicitly."""
    out = _check('carbonate_dh')
    assert abs(out['pH'] - 4.6763534253004) < 1e-12
    assert abs(out['Gamma H+'] - 0.99466955675345) < 1e-12
    assert abs(out['Gamma HCO3-'] - 0.99462956032298) < 1e-12
    out = _check('surface_complexation')
    assert abs(out['Free >FeOH_w'] - 5.0116078057798e+02) < 1e-10
    assert abs(out['>FeOHZn+_w'] - 9.6959833998680e-03) < 1e-14


def test_fixture_base_state_is_oracle_output():
    """The base state stored in the fixtures is what the oracle computes today."""
    for name in ['calcite', 'hanford300a_eq', 'hanford300a_mr', 'hpt_calcite']:
        w = synth.Workload(name)
        t, orc, st, xx, nit, cst = kat.initial_cell_from_fixture(w)
        assert nit == w.meta['equilibrate_iterations']
        for f, v in w.base.items():
            if st[f].shape[0]:
                np.testing.assert_allclose(st[f][:, 0], v, rtol=1e-13, atol=0, err_msg='%s %s' % (name, f))
