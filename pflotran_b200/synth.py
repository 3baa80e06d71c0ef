"""Deterministic synthetic reaction workloads.

A workload = committed chemistry tables + the equilibrated 1-cell base state of the deck's
initial constraint (tests/golden/<name>.json, made by tests/golden/make_fixtures.py from the
reference's decks and databases) + per-cell perturbations drawn from a counter-based RNG:

  tran_xx   = T0 * exp(sigma * z),  z ~ N(0,1), sigma = 0.05; on a `front_fraction` of the
              cells ("reaction front") sigma = front_sigma = 0.5
  pri_molal = m0 (initial guess), activity coefficients / sec_molal / free sites = base state
  porosity ~ U(0.2, 0.4), sat = 1, den_kg / temp / pres = reference values of the deck
  (logK_mode != FIXED: T ~ U(25,150) C, P ~ U(1e5,3e7) Pa)
  mnrl_volfrac ~ U(0, 0.2) with 5 % exact zeros, mnrl_area = base state
  kinmr_total_sorb(:, r) = base state (f_r * S_eq(m0))

Cells are generated in blocks of BLOCK cells, block b using Philox(key=(seed, b)), so the
values of cell c do not depend on how cells are partitioned across ranks / GPUs.
No chemistry arithmetic lives here and nothing is read from /root/reference.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np

from . import abi
from .chem.tables import ReactionTables

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
SEED = 20261017
BLOCK = 4096


class Workload:
    def __init__(self, name: str):
        path = os.path.join(GOLDEN, name + '.json')
        with open(path) as f:
            d = json.load(f)
        self.name = name
        self.meta = d
        self.tables = ReactionTables.from_dict(d['tables'])
        self.base: Dict[str, np.ndarray] = {k: np.array([float(x) for x in v], dtype=np.float64)
                                            for k, v in d['base'].items()}
        self.gold = d.get('gold')

    @property
    def ncomp(self) -> int:
        return self.tables.ncomp

    def base_totals(self) -> np.ndarray:
        return self.base['TOTAL'].copy()

    def base_solution(self) -> np.ndarray:
        """the solution vector of the base state: free-ion molalities, then the immobile concentrations [ncomp]"""
        return np.concatenate([self.base['PRI_MOLAL'], self.base.get('IMMOBILE', np.zeros(0))])


def _block(w: Workload, b: int, sigma: float, front_fraction: float, front_sigma: float, seed: int, variant: int = 0):
    t = w.tables
    n, nk = t.naqcomp, t.nkinmnrl
    rng = np.random.Generator(np.random.Philox(key=[seed, b]))
    z = rng.standard_normal((BLOCK, n))
    if variant:          # another noise realisation for the SAME cells (same front membership, porosity, minerals): "the next time step"
        z = np.random.Generator(np.random.Philox(key=[seed + 7919 * variant, b + (1 << 40)])).standard_normal((BLOCK, n))
    u_front = rng.random(BLOCK)
    por = 0.2 + 0.2 * rng.random(BLOCK)
    vf = 0.2 * rng.random((BLOCK, max(nk, 1)))
    zero = rng.random((BLOCK, max(nk, 1))) < 0.05
    uT = rng.random(BLOCK)
    uP = rng.random(BLOCK)
    sig = np.where(u_front < front_fraction, front_sigma, sigma)[:, None]
    xx = w.base['TOTAL'][None, :] * np.exp(sig * z)
    nim = getattr(t, 'nimmobile', 0)
    if nim > 0:          # immobile dofs follow the aqueous ones; drawn last so that the other streams do not depend on nim
        xx = np.concatenate([xx, w.base['IMMOBILE'][None, :] * np.exp(sig * rng.standard_normal((BLOCK, nim)))], axis=1)
    vf = np.where(zero, 0.0, vf)[:, :nk]
    if t.logK_mode != 0:
        temp = 25.0 + 125.0 * uT
        pres = 1.0e5 + (3.0e7 - 1.0e5) * uP
    else:
        temp = np.full(BLOCK, w.base['TEMP'][0])
        pres = np.full(BLOCK, w.base['PRES'][0])
    return xx, por, vf, temp, pres


def make_cells(w: Workload, start: int, ncells: int, sigma: float = 0.05, front_fraction: float = 0.1,
               front_sigma: float = 0.5, seed: int = SEED, variant: int = 0) -> Dict[str, np.ndarray]:
    """Per-cell inputs for cells [start, start+ncells): tran_xx [ncells, ncomp] (AoS, C order),
    porosity [ncells], volfrac [nkin, ncells], temp, pres [ncells]."""
    t = w.tables
    n, nk = t.ncomp, t.nkinmnrl
    xx = np.empty((ncells, n))
    por = np.empty(ncells)
    vf = np.empty((nk, ncells))
    temp = np.empty(ncells)
    pres = np.empty(ncells)
    b0, b1 = start // BLOCK, (start + ncells - 1) // BLOCK
    for b in range(b0, b1 + 1):
        bx, bp, bv, bt, bpr = _block(w, b, sigma, front_fraction, front_sigma, seed, variant)
        lo = max(start, b * BLOCK)
        hi = min(start + ncells, (b + 1) * BLOCK)
        s = slice(lo - b * BLOCK, hi - b * BLOCK)
        d = slice(lo - start, hi - start)
        xx[d] = bx[s]
        por[d] = bp[s]
        vf[:, d] = bv[s].T
        temp[d] = bt[s]
        pres[d] = bpr[s]
    return {'tran_xx': xx, 'porosity': por, 'volfrac': vf, 'temp': temp, 'pres': pres}


def host_state(w: Workload, cells: Dict[str, np.ndarray]) -> abi.HostState:
    """Host SoA image (for the oracle / emulation harness and for uploads in tests)."""
    ncells = cells['porosity'].shape[0]
    st = abi.HostState(w.tables, ncells)
    for f, v in w.base.items():
        if st[f].shape[0]:
            st[f][:] = v[:, None]
    st['POROSITY'][0] = cells['porosity']
    st['TEMP'][0] = cells['temp']
    st['PRES'][0] = cells['pres']
    if w.tables.nkinmnrl:
        st['MNRL_VOLFRAC'][:] = cells['volfrac']
    return st
