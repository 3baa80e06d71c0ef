"""Generates the committed fixtures under tests/golden/ from the reference's own decks and
thermodynamic databases.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_fixtures.py

For every workload it writes tests/golden/<name>.json holding
  * the flat chemistry tables (ReactionTables) built from the deck + .dat database,
  * the equilibrated 1-cell base state of the deck's initial constraint (oracle run of
    ReactionEquilibrateConstraint + RTUpdateAuxVars, the reference's start-up sequence),
  * the gold values of the reference's .regression.gold file when the deck has one.
The GPU box has no /root/reference: tests, smoke() and bench.py read only these files.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

import kat  # noqa: E402
from pflotran_b200 import abi  # noqa: E402

REF = '/root/reference'

WORKLOADS = {
    # name: (deck, constraint, gold or None)
    'calcite': ('example_problems/100_100_100/calcite/pflotran.in', 'initial', None),
    'hanford300a_eq': ('regression_tests/default/543/543_hanford_srfcplx_base.in', 'groundwater', None),
    'hanford300a_mr': ('regression_tests/default/543/543_hanford_srfcplx_mr.in', 'groundwater', None),
    'hpt_calcite': ('regression_tests/geothermal_hpt/1D_Calcite/calcite_tran_only_hpt.in', 'initial', None),
    'carbonate_unit': ('regression_tests/ascem/batch/carbonate-unit-activity.in', 'initial',
                       'regression_tests/ascem/batch/carbonate-unit-activity.regression.gold'),
    'carbonate_dh': ('regression_tests/ascem/batch/carbonate-debye-huckel-activity.in', 'initial',
                     'regression_tests/ascem/batch/carbonate-debye-huckel-activity.regression.gold'),
    'ca_carbonate_unit': ('regression_tests/ascem/batch/ca-carbonate-unit-activity.in', 'initial',
                          'regression_tests/ascem/batch/ca-carbonate-unit-activity.regression.gold'),
    'ca_carbonate_dh': ('regression_tests/ascem/batch/ca-carbonate-debye-huckel-activity.in', 'initial',
                        'regression_tests/ascem/batch/ca-carbonate-debye-huckel-activity.regression.gold'),
    'calcite_kinetics': ('regression_tests/ascem/batch/calcite-kinetics.in', 'initial',
                         'regression_tests/ascem/batch/calcite-kinetics.regression.gold'),
    'calcite_kinetics_vf': ('regression_tests/ascem/batch/calcite-kinetics-volume-fractions.in', 'initial',
                            'regression_tests/ascem/batch/calcite-kinetics-volume-fractions.regression.gold'),
    'ion_exchange': ('regression_tests/ascem/batch/ion-exchange-valocchi.in', 'initial',
                     'regression_tests/ascem/batch/ion-exchange-valocchi.regression.gold'),
    'surface_complexation': ('regression_tests/ascem/batch/surface-complexation-1.in', 'initial',
                             'regression_tests/ascem/batch/surface-complexation-1.regression.gold'),
    'kd_w_mineral': ('regression_tests/default/batch/solute_KD_w_mineral.in', 'initial',
                     'regression_tests/default/batch/solute_KD_w_mineral.regression.gold'),
    'kd_wo_mineral': ('regression_tests/default/batch/solute_KD_wo_mineral.in', 'initial',
                      'regression_tests/default/batch/solute_KD_wo_mineral.regression.gold'),
    # mineral prefactors (reaction_mineral.F90:743-782) on primary species, 9 primaries / 57 complexes / 6 kinetic minerals
    'mineral_prefactor': ('regression_tests/default/column/mineral_prefactor.in', 'initial_ore', None),
    # RGeneral (reaction.F90:4694-4831), linear formulation, 500 steps of 0.1 d with the deck's own Newton tolerances
    'general_reaction': ('regression_tests/ascem/batch/general-reaction.in', 'Initial',
                         'regression_tests/ascem/batch/general-reaction.regression.gold'),
    # non-isothermal run: 5-term logK fit evaluated per cell (reaction_aux.F90:1336-1408, 1461-1488) + Arrhenius factor
    'calcite_fit5': ('regression_tests/default/anisothermal/thc_1d.in', 'initial_constraint', None),
    # RMicrobial (reaction_microbial.F90:236-450) with Monod / inverse-Monod terms, biomass as an immobile species and
    # RImmobileDecay (reaction_immobile.F90:240-293): log formulation, 111 steps with the reference's step-size controller
    'abcd_microbial': ('regression_tests/default/batch/ABCD_microbial.in', 'initial',
                       'regression_tests/default/batch/ABCD_microbial.regression.gold'),
    # the same with an Arrhenius factor at 35 C / 15 C (activation energy in J/mol resp. kJ/mol)
    'abcd_microbial_act_high': ('regression_tests/default/batch/ABCD_microbial_activation_high.in', 'initial',
                                'regression_tests/default/batch/ABCD_microbial_activation_high.regression.gold'),
    'abcd_microbial_act_low': ('regression_tests/default/batch/ABCD_microbial_activation_low.in', 'initial',
                               'regression_tests/default/batch/ABCD_microbial_activation_low.regression.gold'),
    # microbial reaction without biomass (no immobile dof), linear formulation.  Parity fixture only: the deck's gold run drives A(aq)
    # to 1e-63 with Newton tolerances of 1e-50 and ends each solve on SNES criteria outside this path, so it is not asserted
    'ab_microbial_linear': ('regression_tests/default/batch/AB_microbial_linear_scaling.in', 'initial', None),
    # BASELINE config 1: 22 primaries / 164 complexes (example_problems/ascem_chemistry, savannah_river.dat)
    'ascem': ('example_problems/ascem_chemistry/pflotran.in', 'initial', None),
}
NON_ISOTHERMAL = {'calcite_fit5'}


# Decks the reference does not ship: a reference deck with a documented text patch, so that every coded branch of the path is
# reached by a fixture (NEWTON activity algorithm, ACTIVITY_WATER, free-site inner Newton, Langmuir / Freundlich
# isotherms, the optional mineral rate-law parameters).  (base deck, constraint, [(old text, new text), ...])
_MR_BLOCK_START = '    SURFACE_COMPLEXATION_RXN'
VARIANTS = {
    # RActivityCoefficients NEWTON algorithm at every Newton iteration (reaction.F90:3846-3990) + activity of water (:4043-4050)
    'hanford300a_act_newton': ('regression_tests/default/543/543_hanford_srfcplx_base.in', 'groundwater',
                               [('ACTIVITY_COEFFICIENTS NEWTON_ITERATION', 'ACTIVITY_COEFFICIENTS NEWTON NEWTON_ITERATION\n  ACTIVITY_WATER')]),
    # surface complexes with two free sites per complex: srfcplxrxn_stoich_flag, inner Newton on the free-site
    # concentration (reaction_surf_complex.F90:793-818); database complexes >S2---, >SO2UO2, >SO2UO2CO3-- on site >S(OH)2
    'hanford300a_stoich': ('regression_tests/default/543/543_hanford_srfcplx_base.in', 'groundwater',
                           [('      SITE >SOH 152.64d0', '      SITE >S(OH)2 152.64d0'),
                            ('        >SOUO2OH\n        >SOHUO2CO3\n', '        >S2---\n        >SO2UO2\n        >SO2UO2CO3--\n')]),
    # RTotalSorbKD Langmuir and Freundlich isotherms (reaction.F90:4220-4301).  The reader resets the type to LINEAR on
    # every keyword of the block (reaction.F90:535), so the keyword that sets the type has to come last.
    'kd_langmuir': ('regression_tests/default/batch/solute_KD_w_mineral.in', 'initial',
                    [('KD_MINERAL_NAME A(s)', 'KD_MINERAL_NAME A(s)\n        LANGMUIR_B 2.5d5')]),
    'kd_freundlich': ('regression_tests/default/batch/solute_KD_w_mineral.in', 'initial',
                      [('KD_MINERAL_NAME A(s)', 'KD_MINERAL_NAME A(s)\n        FREUNDLICH_N 0.8d0')]),
    # RKineticMineral optional rate-law parameters (reaction_mineral.F90:699-870): Temkin constant, mineral scale factor,
    # affinity power and threshold, rate limiter, Arrhenius activation energy
    # RRadioactiveDecay (reaction.F90:4607-4690): the first-order A -> B reaction of general-reaction.in as a decay reaction
    'decay_ab': ('regression_tests/ascem/batch/general-reaction.in', 'Initial',
                 [('  GENERAL_REACTION\n    REACTION A(aq) <-> B(aq)\n    FORWARD_RATE 1.15741d-6 ! 0.1 1/d\n    BACKWARD_RATE 0.d0\n  /',
                   '  RADIOACTIVE_DECAY_REACTION\n    REACTION A(aq) <-> B(aq)\n    HALF_LIFE 6.9314718056 d\n  /')]),
    # RKineticSurfCplx (reaction_surf_complex.F90:938-1137): the 300A surface complexation reaction (reaction 1, on kinetic
    # mineral 1 = Calcite) with forward / backward rates instead of equilibrium
    'hanford300a_kinsrf': ('regression_tests/default/543/543_hanford_srfcplx_base.in', 'groundwater',
                           [('      MINERAL Calcite\n', '      KINETIC\n      COMPLEX_KINETICS\n        >SOUO2OH\n          FORWARD_RATE_CONSTANT 2.d-3\n'
                             '          BACKWARD_RATE_CONSTANT 1.d-4\n        /\n        >SOHUO2CO3\n          FORWARD_RATE_CONSTANT 5.d-2\n'
                             '          BACKWARD_RATE_CONSTANT 2.d-4\n        /\n      /\n      MINERAL Calcite\n')]),
    # BASELINE config 4's "CO2 style high-ionic-strength chemistry": the aqueous chemistry of the reference's MPHASE CO2 deck
    # (8 primaries with CO2(aq) swapped into the basis, 12 complexes, kinetic quartz + calcite, 1 molal NaCl brine) without its
    # ACTIVE_GAS_SPECIES card - the supercritical phase (RTotalCO2, reaction_gas.F90:172-296) is outside the path
    'scco2_brine': ('regression_tests/default/scco2/mphase/mphase_chem.in', 'initial',
                    [('  ACTIVE_GAS_SPECIES\n    CO2(g)\n    O2(g)\n  /\n', '')]),
    # RMicrobial's other inhibition branches (reaction_microbial.F90:322-339, 393-412): THRESHOLD (atan) and MONOD next to the deck's
    # INVERSE_MONOD term
    'abcd_microbial_inhibition': ('regression_tests/default/batch/ABCD_microbial.in', 'initial',
                                  [('    BIOMASS\n      SPECIES_NAME D(im)', '    INHIBITION\n      SPECIES_NAME B(aq)\n      TYPE THRESHOLD 1.d5\n'
                                    '      INHIBITION_CONSTANT 9.d-4\n    /\n    INHIBITION\n      SPECIES_NAME A(aq)\n      TYPE MONOD\n'
                                    '      INHIBITION_CONSTANT 1.d-3\n    /\n    BIOMASS\n      SPECIES_NAME D(im)')]),
    'calcite_rate_laws': ('regression_tests/ascem/batch/calcite-kinetics.in', 'initial',
                          [('      RATE_CONSTANT 1.d-13 mol/cm^2-sec\n',
                            '      RATE_CONSTANT 1.d-13 mol/cm^2-sec\n      ACTIVATION_ENERGY 40.d0\n      AFFINITY_THRESHOLD 1.d-3\n'
                            '      AFFINITY_POWER 1.5d0\n      TEMKIN_CONSTANT 2.d0\n      MINERAL_SCALE_FACTOR 1.2d0\n      RATE_LIMITER 1.d-9\n')]),
}

for _v in VARIANTS:
    WORKLOADS[_v] = (None, VARIANTS[_v][1], None)


def variant_deck(name):
    base, constraint, patches = VARIANTS[name]
    src = os.path.join(REF, base)
    text = open(src).read()
    for old, new in patches:
        assert text.count(old) >= 1, (name, old)
        text = text.replace(old, new)
    # the variant lives in a scratch directory: make the database path absolute
    import re
    m = re.search(r'^\s*DATABASE\s+(\S+)', text, re.M)
    dbase = os.path.normpath(os.path.join(os.path.dirname(src), m.group(1)))
    text = text.replace(m.group(0), '  DATABASE ' + dbase)
    out = os.path.join('/tmp', 'rxn_b200_variant_' + name + '.in')
    with open(out, 'w') as f:
        f.write(text)
    return out, constraint, 'variant of %s: %s' % (base, '; '.join('%r -> %r' % (o.strip(), n.strip()) for o, n in patches))


def ascem_kat(out_path, constraint):
    """Speciation of `constraint` as the reference printed it (ReactionPrintConstraint, 5 significant figures):
    example_problems/ascem_chemistry/pflotran.out:5811-5870 - iteration count, free / total molality of every primary
    species, molality of the listed complexes."""
    import re
    lines = open(out_path).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.strip() == 'Constraint: ' + constraint)
    kat = {'iterations': None, 'primary': {}, 'complex': {}, 'source': 'example_problems/ascem_chemistry/pflotran.out:%d' % (i0 + 1)}
    mode = None
    for l in lines[i0:i0 + 400]:
        m = re.match(r'\s*iterations:\s+(\d+)', l)
        if m:
            kat['iterations'] = int(m.group(1))
        if l.strip().startswith('species') and 'molal' in l:
            mode = 'primary'; continue
        if l.strip().startswith('complex') and 'molality' in l and 'logK' in l:
            mode = 'complex'; continue
        if l.strip().startswith('primary species:'):
            break
        f = l.split()
        if mode == 'primary' and len(f) >= 4 and re.match(r'^-?\d\.\d+E[+-]\d+$', f[1]):
            kat['primary'][f[0]] = [float(f[1]), float(f[2])]
        elif mode == 'complex' and len(f) >= 4 and re.match(r'^-?\d\.\d+E[+-]\d+$', f[1]):
            kat['complex'][f[0]] = float(f[1])
    return kat


def time_block(path):
    """TIME card (FINAL_TIME / INITIAL_TIMESTEP_SIZE / MAXIMUM_TIMESTEP_SIZE, converted to seconds as
    units.F90 does) and TS_ACCELERATION of the TIMESTEPPER card: what tests/gi_driver.run_deck needs."""
    import re
    from pflotran_b200.chem.units import units_convert_to_internal
    out = {'iaccel': 5}
    for line in open(path):
        line = line.split('!')[0].split('#')[0]
        m = re.match(r'\s*(FINAL_TIME|INITIAL_TIMESTEP_SIZE|MAXIMUM_TIMESTEP_SIZE)\s+(\S+)\s+(\S+)', line)
        if m:
            out[m.group(1)] = float(m.group(2).lower().replace('d', 'e')) * units_convert_to_internal(m.group(3), 's')
        m = re.match(r'\s*TS_ACCELERATION\s+(\d+)', line)
        if m:
            out['iaccel'] = int(m.group(1))
        m = re.match(r'\s*MAX_STEPS\s+(-?\d+)', line)
        if m:
            out['MAX_STEPS'] = int(m.group(1))
    # NEWTON_SOLVER TRANSPORT card: RTOL / ATOL / STOL (solver.F90:842-851)
    text = open(path).read()
    m = re.search(r'^NEWTON_SOLVER\s+TRANSPORT(.*?)^/', text, re.M | re.S)
    if m:
        for key in ('RTOL', 'ATOL', 'STOL'):
            mm = re.search(r'^\s*%s\s+(\S+)' % key, m.group(1), re.M)
            if mm:
                out[key] = float(mm.group(1).lower().replace('d', 'e'))
    return out


def main():
    only = set(sys.argv[1:])
    for name, (deck, constraint, gold) in WORKLOADS.items():
        if only and name not in only:
            continue
        if deck is None:
            path, constraint, deck = variant_deck(name)
        else:
            path = os.path.join(REF, deck)
        d, t, orc, st, xx, nit, cst = kat.initial_cell(path, constraint=constraint, isothermal=name not in NON_ISOTHERMAL)
        base = {f: [repr(float(x)) for x in st[f][:, 0]] for f in abi.FIELDS
                if f not in ('DTOTAL', 'DTOTAL_SORB_EQ')}
        if name == 'hanford300a_kinsrf':
            # a state with sorbed kinetic complexes (the deck starts from none): S^k = 2 % and 5 % of the site density
            base['KINSRFCPLX_CONC'] = [repr(0.02 * float(t.srfcplxrxn_site_density[0])), repr(0.05 * float(t.srfcplxrxn_site_density[0]))]
        from pflotran_b200.chem.setup import constraint_arrays, mineral_arrays
        ctype, conc, cid, guess = constraint_arrays(t, d.constraints[constraint])
        vf, area = mineral_arrays(t, d.constraints[constraint])
        cons = {'ctype': [int(x) for x in ctype], 'conc': [repr(float(x)) for x in conc], 'cid': [int(x) for x in cid],
                'guess': None if guess is None else [repr(float(x)) for x in guess],
                'volfrac': [repr(float(x)) for x in vf], 'area': [repr(float(x)) for x in area],
                'immobile': [repr(float(x)) for x in kat.immobile_array(t, d.constraints[constraint])]}
        out = {
            'constraint_arrays': cons,
            'name': name, 'deck': deck, 'constraint': constraint, 'equilibrate_iterations': int(nit),
            'porosity': d.porosity, 'rock_density': d.rock_density, 'tables': t.to_dict(), 'base': base, 'time': time_block(path),
        }
        if name == 'ascem':
            out['kat'] = ascem_kat(os.path.join(REF, 'example_problems/ascem_chemistry/pflotran.out'), constraint)
        if gold:
            g = kat.read_gold(os.path.join(REF, gold))
            out['gold_file'] = gold
            out['gold'] = {k: v for k, v in g.items() if isinstance(v, dict) and v}
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(out, f, separators=(',', ':'))
        print(name, t.work_counts(), 'equilibrate its', nit, os.path.getsize(os.path.join(HERE, name + '.json')), 'bytes')


if __name__ == '__main__':
    main()
