"""Harness that replays the reference's start-up sequence for 1-cell batch decks so the
oracle can be compared with the reference's .regression.gold files.

Sequence restated (test infrastructure, drives oracle functions only):
  1. PatchInitCouplerConstraints (src/pflotran/patch.F90:3346-3461): equilibrate the
     constraint on the coupler's own auxvar with den_kg = reference water density,
     porosity = option%reference_porosity = 0.25 (option.F90:479), sat = 1.
  2. CondControlAssignTranInitCond (condition_control.F90:498-949): cell auxvar is fresh
     (act coefs 1, free site 1e-9); xx = basis_molarity / den_kg * 1000; mineral volume
     fractions / areas from the constraint; multirate sorbed + free sites copied.
  3. RTUpdateAuxVars once without, then twice with activity-coefficient updates
     (condition_control.F90:940-946).
"""
import math
import os
import re

import numpy as np

from pflotran_b200 import abi
from pflotran_b200.chem import read_deck, build_tables
from pflotran_b200.chem.setup import constraint_arrays, mineral_arrays
from oracle.pyoracle import Oracle


def read_gold(path):
    out = {}
    name = None
    with open(path) as f:
        for line in f:
            m = re.match(r'^-- (\w+): (.*) --', line)
            if m:
                name = m.group(2).strip()
                out[name] = {}
                continue
            m = re.match(r'^\s*(\w[\w ()-]*):\s+(\S+)\s*$', line)
            if m and name is not None:
                try:
                    out[name][m.group(1).strip()] = float(m.group(2))
                except ValueError:
                    pass
    return out


def fill_scalars(st, t, porosity, volume=1.0, rock_density=None):
    st['DEN_KG'][:] = t.reference_water_density
    st['SAT'][:] = 1.0
    st['TEMP'][:] = t.reference_temperature
    st['PRES'][:] = t.reference_pressure
    st['VOLUME'][:] = volume
    st['POROSITY'][:] = porosity
    st['SOIL_PARTICLE_DENSITY'][:] = -999.0 if rock_density is None else rock_density   # UNINITIALIZED_DOUBLE unless ROCK_DENSITY


def initial_cell(deck_path, constraint='initial', porosity=None, volume=1.0, isothermal=True, database_path=None):
    deck = read_deck(deck_path)
    t = build_tables(deck, database_path=database_path, isothermal=isothermal)
    orc = Oracle(t)
    c = deck.constraints[constraint]
    ctype, conc, cid, guess = constraint_arrays(t, c)
    vf, area = mineral_arrays(t, c)
    # 1. coupler auxvar
    cst = abi.HostState(t, 1)
    fill_scalars(cst, t, 0.25, volume, deck.rock_density)
    cst['MNRL_VOLFRAC'][:, 0] = vf
    cst['MNRL_AREA'][:, 0] = area
    basis_molarity, nit = orc.equilibrate(cst, 0, ctype, conc, cid, guess, use_prev=False)
    # 2. the grid cell
    st = abi.HostState(t, 1)
    fill_scalars(st, t, deck.porosity if porosity is None else porosity, volume, deck.rock_density)
    st['MNRL_VOLFRAC'][:, 0] = vf
    st['MNRL_AREA'][:, 0] = area
    if t.nkinmrsrfcplxrxn > 0:
        st['KINMR_TOTAL_SORB'][:] = cst['KINMR_TOTAL_SORB']
        st['FREE_SITE_CONC'][:] = cst['FREE_SITE_CONC']
    xx = (basis_molarity / t.reference_water_density * 1000.0).reshape(1, -1).copy()
    xx = with_immobile(t, xx, immobile_array(t, c))
    # 3.
    orc.update_auxvars(st, xx, False)
    if t.act_coef_update_frequency != 0:
        orc.update_auxvars(st, xx, True)
        orc.update_auxvars(st, xx, True)
    return deck, t, orc, st, xx, nit, cst


def immobile_array(t, c):
    """immobile concentrations of a deck constraint in immobile-species order (ImmobileProcessConstraint, reaction_immobile.F90:163-236)"""
    return np.array([c.immobile.get(n, 0.0) for n in getattr(t, 'immobile_names', [])], dtype=np.float64)


def with_immobile(t, xx, im):
    """solution vector [aqueous free-ion molalities, immobile concentrations] (condition_control.F90:842-856)"""
    if getattr(t, 'nimmobile', 0) == 0:
        return xx
    return np.concatenate([xx, np.tile(np.asarray(im, dtype=np.float64).reshape(1, -1), (xx.shape[0], 1))], axis=1)


class OracleBackend:
    """equilibrate / update_auxvars on a 1-cell HostState through the oracle."""

    def __init__(self, t):
        self.orc = Oracle(t)

    def equilibrate(self, st, ctype, conc, cid, guess):
        return self.orc.equilibrate(st, 0, ctype, conc, cid, guess, use_prev=False)

    def update_auxvars(self, st, xx, act):
        self.orc.update_auxvars(st, xx, act)


def fixture_constraint(w):
    ca = w.meta['constraint_arrays']
    ctype = np.array(ca['ctype'], dtype=np.int32)
    conc = np.array([float(x) for x in ca['conc']])
    cid = np.array(ca['cid'], dtype=np.int32)
    guess = None if ca['guess'] is None else np.array([float(x) for x in ca['guess']])
    vf = np.array([float(x) for x in ca['volfrac']])
    area = np.array([float(x) for x in ca['area']])
    return ctype, conc, cid, guess, vf, area


def fixture_immobile(w):
    return np.array([float(x) for x in w.meta['constraint_arrays'].get('immobile', [])], dtype=np.float64)


def initial_cell_from_fixture(w, porosity=None, volume=1.0, backend=None):
    """Same start-up sequence as initial_cell(), driven only by a committed fixture
    (tests/golden/<name>.json): no deck, no database, no /root/reference.  `backend` (default: the oracle)
    supplies equilibrate() and update_auxvars() - the device routines under test plug in here."""
    t = w.tables
    orc = backend or OracleBackend(t)
    ctype, conc, cid, guess, vf, area = fixture_constraint(w)
    cst = abi.HostState(t, 1)
    fill_scalars(cst, t, 0.25, volume, w.meta.get('rock_density'))
    cst['MNRL_VOLFRAC'][:, 0] = vf
    cst['MNRL_AREA'][:, 0] = area
    basis_molarity, nit = orc.equilibrate(cst, ctype, conc, cid, guess)
    st = abi.HostState(t, 1)
    fill_scalars(st, t, w.meta['porosity'] if porosity is None else porosity, volume, w.meta.get('rock_density'))
    st['MNRL_VOLFRAC'][:, 0] = vf
    st['MNRL_AREA'][:, 0] = area
    if t.nkinmrsrfcplxrxn > 0:
        st['KINMR_TOTAL_SORB'][:] = cst['KINMR_TOTAL_SORB']
        st['FREE_SITE_CONC'][:] = cst['FREE_SITE_CONC']
    xx = (basis_molarity / t.reference_water_density * 1000.0).reshape(1, -1).copy()
    xx = with_immobile(t, xx, fixture_immobile(w))
    orc.update_auxvars(st, xx, False)
    if t.act_coef_update_frequency != 0:
        orc.update_auxvars(st, xx, True)
        orc.update_auxvars(st, xx, True)
    return t, (orc.orc if isinstance(orc, OracleBackend) else orc), st, xx, nit, cst


def outputs(t, st, cell=0):
    """Variables as the reference prints them (patch.F90 PatchGetVariable)."""
    o = {}
    if t.h_ion_id > 0:
        h = t.h_ion_id - 1
        o['pH'] = -math.log10(st['PRI_ACT_COEF'][h, cell] * st['PRI_MOLAL'][h, cell])
    den = st['DEN_KG'][0, cell]
    for i, n in enumerate(t.primary_species_names):
        tot = st['TOTAL'][i, cell]
        o['Total ' + n] = tot / den * 1000.0 if t.initialize_with_molality else tot
        o['Free ' + n] = st['PRI_MOLAL'][i, cell] if t.initialize_with_molality else \
            st['PRI_MOLAL'][i, cell] * den / 1000.0
        o['Gamma ' + n] = st['PRI_ACT_COEF'][i, cell]
        o['Total Sorbed ' + n] = st['TOTAL_SORB_EQ'][i, cell]
    for i, n in enumerate(t.kinmnrl_names):
        o[n + ' VF'] = st['MNRL_VOLFRAC'][i, cell]
        o[n + ' Rate'] = st['MNRL_RATE'][i, cell]
    for i, n in enumerate(t.srfcplx_names):
        o[n] = st['EQSRFCPLX_CONC'][i, cell]
    for i, n in enumerate(t.srfcplxrxn_site_names):
        o['Free ' + n] = st['FREE_SITE_CONC'][i, cell]
    for i, n in enumerate(getattr(t, 'immobile_names', [])):
        o[n] = st['IMMOBILE'][i, cell]
    return o


def check_speciation_kat(w, t, cst, nit, rtol=6.0e-5):
    """Constraint speciation as the reference printed it (fixture key `kat`, from the deck's own pflotran.out: iteration
    count, free and total molality of every primary species, molality of every listed complex; 5 significant figures,
    hence rtol 6e-5).  cst: the equilibrated constraint auxvar.  Returns the number of values checked."""
    k = w.meta['kat']
    assert nit == k['iterations'], (nit, k['iterations'])
    den = cst['DEN_KG'][0, 0]
    checked = 0
    for i, n in enumerate(t.primary_species_names):
        free, tot = k['primary'][n]
        assert abs(cst['PRI_MOLAL'][i, 0] - free) <= rtol * abs(free), (n, cst['PRI_MOLAL'][i, 0], free)
        assert abs(cst['TOTAL'][i, 0] / den * 1000.0 - tot) <= rtol * abs(tot), (n, cst['TOTAL'][i, 0] / den * 1000.0, tot)
        checked += 2
    names = list(t.secondary_species_names)
    for n, v in k['complex'].items():
        j = names.index(n)
        assert abs(cst['SEC_MOLAL'][j, 0] - v) <= rtol * abs(v), (n, cst['SEC_MOLAL'][j, 0], v)
        checked += 1
    return checked
