"""Thermodynamic database reader (hanford.dat / geothermal-hpt.dat format).

Host-side setup.  Restates reference src/pflotran/reaction_database.F90:25-424
(DatabaseRead) and reaction_mineral.F90:416-477 (MineralReadFromDatabase):
line 1 gives the temperature points (or 'Number of Parameters' 17 for hpt);
sections are separated by lines whose first quoted word is 'null':
primaries, aqueous complexes, gases, minerals, surface complexes.
Only species named in the deck are kept.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional


def _tokens(line: str) -> List[str]:
    """Split on blanks, keeping 'quoted words' (which may hold blanks)."""
    out = []
    i = 0
    n = len(line)
    while i < n:
        c = line[i]
        if c in ' \t\r\n,':
            i += 1
        elif c == "'":
            j = line.index("'", i + 1)
            out.append(line[i + 1:j])
            i = j + 1
        else:
            j = i
            while j < n and line[j] not in ' \t\r\n,':
                j += 1
            out.append(line[i:j])
            i = j
    return out


def _f(tok: str) -> float:
    return float(tok.replace('d', 'e').replace('D', 'e'))


@dataclass
class DbRxn:
    spec_name: List[str]
    stoich: List[float]
    logK: List[float]


@dataclass
class AqSpecies:
    name: str
    a0: float = 0.0
    Z: float = 0.0
    molar_weight: float = 0.0
    dbaserxn: Optional[DbRxn] = None
    found: bool = False
    is_redox: bool = False


@dataclass
class GasSpecies:
    name: str
    molar_volume: float = 0.0
    molar_weight: float = 0.0
    dbaserxn: Optional[DbRxn] = None
    found: bool = False


@dataclass
class MineralSpecies:
    name: str
    molar_volume: float = 0.0
    molar_weight: float = 0.0
    dbaserxn: Optional[DbRxn] = None
    found: bool = False


@dataclass
class SrfCplxSpecies:
    name: str
    free_site_name: str = ''
    free_site_stoich: float = 0.0
    Z: float = 0.0
    dbaserxn: Optional[DbRxn] = None
    found: bool = False


@dataclass
class DatabaseContent:
    temperatures: List[float]
    num_logKs: int
    primary: Dict[str, AqSpecies] = field(default_factory=dict)
    secondary: Dict[str, AqSpecies] = field(default_factory=dict)
    gases: Dict[str, GasSpecies] = field(default_factory=dict)
    minerals: Dict[str, MineralSpecies] = field(default_factory=dict)
    srfcplx: Dict[str, SrfCplxSpecies] = field(default_factory=dict)


def read_database(path: str, primary: List[str], secondary: List[str], gases: List[str],
                  minerals: List[str], srfcplx: List[str], hpt: bool = False) -> DatabaseContent:
    with open(path) as f:
        lines = [l for l in f.read().splitlines()]
    it = iter(lines)

    def next_line():
        for l in it:
            s = l.strip()
            if not s or s[0] in '#!':
                continue
            return s
        return None

    head = _tokens(next_line())
    num = int(head[1])
    temps = [] if hpt else [_f(t) for t in head[2:2 + num]]
    db = DatabaseContent(temps, num)
    db.primary = {n: AqSpecies(n) for n in primary}
    db.secondary = {n: AqSpecies(n) for n in secondary}
    db.gases = {n: GasSpecies(n) for n in gases}
    db.minerals = {n: MineralSpecies(n) for n in minerals}
    db.srfcplx = {n: SrfCplxSpecies(n) for n in srfcplx}

    num_nulls = 0
    max_nulls = 4 if hpt else 5
    while True:
        line = next_line()
        if line is None:
            break
        t = _tokens(line)
        name = t[0]
        if name == 'null':
            num_nulls += 1
            if num_nulls >= max_nulls:
                break
            continue
        if num_nulls in (0, 1):
            sp = None
            if name in db.primary:
                sp = db.primary[name]
            elif name in db.secondary:
                sp = db.secondary[name]
            if sp is None:
                continue
            sp.found = True
            p = 1
            if num_nulls > 0:
                nspec = int(t[p]); p += 1
                st, nm = [], []
                for _ in range(nspec):
                    st.append(_f(t[p])); nm.append(t[p + 1]); p += 2
                logK = [_f(x) for x in t[p:p + num]]; p += num
                sp.dbaserxn = DbRxn(nm, st, logK)
            sp.a0 = _f(t[p]); sp.Z = _f(t[p + 1]); sp.molar_weight = _f(t[p + 2])
        elif num_nulls == 2:
            if name not in db.gases:
                continue
            g = db.gases[name]
            g.found = True
            p = 1
            g.molar_volume = _f(t[p]) * 1.0e-6; p += 1
            nspec = int(t[p]); p += 1
            st, nm = [], []
            for _ in range(nspec):
                st.append(_f(t[p])); nm.append(t[p + 1]); p += 2
            logK = [_f(x) for x in t[p:p + num]]; p += num
            g.dbaserxn = DbRxn(nm, st, logK)
            g.molar_weight = _f(t[p])
        elif num_nulls == 3:
            if name not in db.minerals:
                continue
            m = db.minerals[name]
            m.found = True
            p = 1
            m.molar_volume = _f(t[p]) * 1.0e-6; p += 1
            nspec = int(t[p]); p += 1
            st, nm = [], []
            for _ in range(nspec):
                st.append(_f(t[p])); nm.append(t[p + 1]); p += 2
            logK = [_f(x) for x in t[p:p + num]]; p += num
            m.dbaserxn = DbRxn(nm, st, logK)
            m.molar_weight = _f(t[p])
        elif num_nulls == 4:
            if name not in db.srfcplx:
                continue
            s = db.srfcplx[name]
            s.found = True
            p = 1
            nspec = int(t[p]); p += 1
            st, nm = [], []
            for _ in range(nspec):
                stoich = _f(t[p]); nme = t[p + 1]; p += 2
                if nme.startswith('>'):
                    s.free_site_name = nme
                    s.free_site_stoich = stoich
                else:
                    st.append(stoich); nm.append(nme)
            logK = [_f(x) for x in t[p:p + num]]; p += num
            s.dbaserxn = DbRxn(nm, st, logK)
            s.Z = _f(t[p])

    missing = [n for grp in (db.primary, db.secondary, db.gases, db.minerals, db.srfcplx)
               for n, s in grp.items() if not s.found]
    if missing:
        raise RuntimeError('species not found in database %s: %s' % (path, missing))
    return db
