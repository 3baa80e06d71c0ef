"""Host-side helpers that turn deck CONSTRAINT cards into the flat arrays the
reaction path takes (ordering by primary species, name -> id linking).

Follows reference src/pflotran/reaction.F90:1104-1304 (ReactionProcessConstraint):
constraint lines are re-ordered to the primary-species order and mineral / gas
constraint names are linked to ids in the mineral / passive-gas lists.
No chemistry arithmetic lives here.
"""
from __future__ import annotations

import numpy as np

from . import deck as dk


def constraint_arrays(t, c: dk.Constraint):
    naq = t.naqcomp
    ctype = np.zeros(naq, dtype=np.int32)
    conc = np.zeros(naq, dtype=np.float64)
    cid = np.zeros(naq, dtype=np.int32)
    seen = set()
    for name, val, ty, aux in zip(c.names, c.conc, c.ctype, c.aux):
        if name not in t.primary_species_names:
            raise RuntimeError('Species %s from CONSTRAINT %s not found among primary species.'
                               % (name, c.name))
        j = t.primary_species_names.index(name)
        seen.add(j)
        ctype[j] = ty
        conc[j] = val
        if ty == dk.CONSTRAINT_MINERAL:
            if aux not in t.mineral_names:
                raise RuntimeError('Constraint mineral: %s not found.' % aux)
            cid[j] = t.mineral_names.index(aux) + 1
        elif ty in (dk.CONSTRAINT_GAS, dk.CONSTRAINT_SUPERCRIT_CO2):
            if aux not in t.passive_gas_names:
                raise RuntimeError('Constraint gas: %s not found.' % aux)
            cid[j] = t.passive_gas_names.index(aux) + 1
    if len(seen) != naq:
        raise RuntimeError('Number of concentration constraints is less than number of primary '
                           'species in aqueous constraint.')
    guess = None
    if c.free_ion_guess is not None:
        guess = np.zeros(naq)
        for name, val in c.free_ion_guess.items():
            guess[t.primary_species_names.index(name)] = val
    return ctype, conc, cid, guess


def mineral_arrays(t, c: dk.Constraint):
    """kinetic-mineral volume fractions and specific areas [m^2/m^3] in kinetic order."""
    vf = np.zeros(t.nkinmnrl)
    area = np.zeros(t.nkinmnrl)
    for i, n in enumerate(t.kinmnrl_names):
        if n in c.minerals:
            vf[i], area[i] = c.minerals[n]
    return vf, area
