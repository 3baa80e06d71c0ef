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
}


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
    return out


def main():
    only = set(sys.argv[1:])
    for name, (deck, constraint, gold) in WORKLOADS.items():
        if only and name not in only:
            continue
        path = os.path.join(REF, deck)
        d, t, orc, st, xx, nit, cst = kat.initial_cell(path, constraint=constraint)
        base = {f: [repr(float(x)) for x in st[f][:, 0]] for f in abi.FIELDS
                if f not in ('DTOTAL', 'DTOTAL_SORB_EQ')}
        from pflotran_b200.chem.setup import constraint_arrays, mineral_arrays
        ctype, conc, cid, guess = constraint_arrays(t, d.constraints[constraint])
        vf, area = mineral_arrays(t, d.constraints[constraint])
        cons = {'ctype': [int(x) for x in ctype], 'conc': [repr(float(x)) for x in conc], 'cid': [int(x) for x in cid],
                'guess': None if guess is None else [repr(float(x)) for x in guess],
                'volfrac': [repr(float(x)) for x in vf], 'area': [repr(float(x)) for x in area]}
        out = {
            'constraint_arrays': cons,
            'name': name, 'deck': deck, 'constraint': constraint, 'equilibrate_iterations': int(nit),
            'porosity': d.porosity, 'tables': t.to_dict(), 'base': base, 'time': time_block(path),
        }
        if gold:
            g = kat.read_gold(os.path.join(REF, gold))
            out['gold_file'] = gold
            out['gold'] = {k: v for k, v in g.items() if isinstance(v, dict) and v}
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(out, f, separators=(',', ':'))
        print(name, t.work_counts(), 'equilibrate its', nit, os.path.getsize(os.path.join(HERE, name + '.json')), 'bytes')


if __name__ == '__main__':
    main()
