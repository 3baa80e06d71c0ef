"""Flat chemistry tables: the read-only `reaction_type` data the hot path uses.

Field names and array shapes follow the reference's compressed tables
(src/pflotran/reaction_aux.F90:142-335, reaction_mineral_aux.F90:77-128,
reaction_surf_complex_aux.F90:68-128).  Arrays are stored exactly in the
memory order the Fortran arrays have, so the same bytes can be handed to the
C ABI (`RxnTablesDesc`, include/rxn_b200.h) by a Fortran caller with `c_loc`:

  Fortran  specid(0:m, n)   <->  numpy int32  [n, m+1]   (row 0 of Fortran = count)
  Fortran  stoich(0:m, n)   <->  numpy float64[n, m+1]   (offset 0)
  Fortran  stoich(m, n)     <->  numpy float64[n, m]     (offset 1: species i at [i-1])

Species ids are 1-based, as in the reference.
"""
from __future__ import annotations

import json
from typing import Any, Dict

import numpy as np

_INT_FIELDS = {
    'eqcplxspecid', 'eqcplxh2oid', 'kinmnrlspecid', 'kinmnrlh2oid', 'mnrlspecid', 'mnrlh2oid',
    'kinmnrl_num_prefactors', 'kinmnrl_prefactor_id',
    'srfcplxspecid', 'srfcplxh2oid', 'srfcplxrxn_to_surf', 'srfcplxrxn_surf_type',
    'srfcplxrxn_to_complex', 'srfcplxrxn_stoich_flag', 'eqsrfcplxrxn_to_srfcplxrxn',
    'kinmrsrfcplxrxn_to_srfcplxrxn', 'kinmr_nrate',
    'eqionx_rxn_cationid', 'eqionx_rxn_Z_flag', 'eqionx_rxn_to_surf',
    'eqkdspecid', 'eqkdtype', 'eqkdmineral',
    'paseqspecid', 'paseqh2oid',
    'generalspecid', 'generalforwardspecid', 'generalbackwardspecid', 'radiodecayspecid', 'radiodecayforwardspecid',
    'kinsrfcplxrxn_to_srfcplxrxn',
    'immobile_decayspecid', 'microbial_specid', 'microbial_biomassid', 'microbial_monodid', 'microbial_inhibitionid',
    'microbial_monod_specid', 'microbial_inhibition_type', 'microbial_inhibition_specid',
}


class ReactionTables:
    """Attribute bag; every attribute is a python scalar, list of str, or numpy array."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to_dict(self) -> Dict[str, Any]:
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                # repr() round-trips doubles exactly
                out[k] = {'shape': list(v.shape), 'dtype': str(v.dtype),
                          'data': [repr(float(x)) if v.dtype.kind == 'f' else int(x)
                                   for x in v.ravel()]}
            elif isinstance(v, float):
                out[k] = {'f': repr(v)}
            else:
                out[k] = v
        return out

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> 'ReactionTables':
        kw = {}
        for k, v in d.items():
            if isinstance(v, dict) and 'shape' in v:
                dt = np.dtype(v['dtype'])
                if dt.kind == 'f':
                    arr = np.array([float(x) for x in v['data']], dtype=dt)
                else:
                    arr = np.array(v['data'], dtype=dt)
                kw[k] = arr.reshape(v['shape'])
            elif isinstance(v, dict) and 'f' in v:
                kw[k] = float(v['f'])
            else:
                kw[k] = v
        return ReactionTables(**kw)

    def save(self, path: str):
        with open(path, 'w') as f:
            json.dump(self.to_dict(), f, indent=0, separators=(',', ':'))

    @staticmethod
    def load(path: str) -> 'ReactionTables':
        with open(path) as f:
            return ReactionTables.from_dict(json.load(f))

    # ---- derived sizes used by the roofline accounting ----
    def work_counts(self) -> Dict[str, int]:
        ncplx = self.neqcplx
        S = int(sum(self.eqcplxspecid[k, 0] for k in range(ncplx)))
        S2 = int(sum(int(self.eqcplxspecid[k, 0]) ** 2 for k in range(ncplx)))
        return {'naq': self.naqcomp, 'ncomp': self.ncomp, 'ncplx': ncplx, 'S': S, 'S2': S2,
                'nkin': self.nkinmnrl, 'nsrfcplx': self.nsrfcplx}


def idarray(rows, ld):
    """rows: list of lists of 1-based ids -> int32 [n, ld] with the count in column 0."""
    a = np.zeros((len(rows), ld), dtype=np.int32)
    for i, r in enumerate(rows):
        a[i, 0] = len(r)
        a[i, 1:1 + len(r)] = r
    return a


def starray0(rows, ld):
    """stoich(0:m, n) layout: float64 [n, ld], entry i (1-based) at column i."""
    a = np.zeros((len(rows), ld), dtype=np.float64)
    for i, r in enumerate(rows):
        a[i, 1:1 + len(r)] = r
    return a


def starray1(rows, ld):
    """stoich(m, n) layout: float64 [n, ld], entry i (1-based) at column i-1."""
    a = np.zeros((len(rows), max(ld, 1)), dtype=np.float64)
    for i, r in enumerate(rows):
        a[i, 0:len(r)] = r
    return a
