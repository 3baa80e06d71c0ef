"""Deck + database -> flat chemistry tables (`ReactionTables`).

Host-side setup.  Restates the parts of reference
src/pflotran/reaction_database.F90:772-3590 (BasisInit) that the reaction hot
path needs:

  * Debye-Hueckel A/B/Bdot at the reference temperature (:897-990), NO_BDOT;
  * bracketing database temperatures for isothermal logK (:992-1016) and the
    linear `Interpolate` (utility.F90:677-700);
  * basis swap through the LU inverse of the secondary block (:1100-1276),
    with the `> 1.d-40` species filter (:1304-1329);
  * substitution of secondary/gas species into kinetic-mineral and surface
    complex reactions (:1392-1518; reaction_database_aux.F90:355-545, including
    the `1.d-10` cancellation filter) and re-alignment to basis order (:278-351);
  * H2O split-out when packing (:1683-1702, :2123-2139, :2436-2452);
  * 5-term logK least-squares fit for non-isothermal runs
    (reaction_aux.F90:1336-1408) and the 17-coefficient hpt form (:1529-1571);
  * mineral TST parameter arrays incl. the "allocated only if some mineral sets
    it" flags (:1967-2020), prefactors (:2200-2255);
  * surface-complexation reaction tables (:2342-2774), ion exchange (:2818-2914),
    KD isotherms (:3374-3461), species_idx h+/h2o (:3466-3510).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List

import numpy as np

from . import deck as dk
from .database import DbRxn, read_database
from .nr import ludcmp, lubksb
from .tables import ReactionTables, idarray, starray0, starray1
from .eos import water_density_ifc67

H2O_NAME = 'H2O'


def interpolate(x_high, x_low, x, y_high, y_low):
    """utility.F90:677-700"""
    x_diff = x_high - x_low
    if abs(x_diff) < 1.0e-10:
        return y_low
    weight = (x - x_low) / x_diff
    return y_low + weight * (y_high - y_low)


_DEBYE = [  # (t_low, t_high, A_high, A_low, B_high, B_low, Bdot_high, Bdot_low)
    (0.0, 25.0, 0.5114, 0.4939, 0.3288, 0.3253, 0.0410, 0.0374),
    (25.0, 60.0, 0.5465, 0.5114, 0.3346, 0.3288, 0.0440, 0.0410),
    (60.0, 100.0, 0.5995, 0.5465, 0.3421, 0.3346, 0.0460, 0.0440),
    (100.0, 150.0, 0.6855, 0.5995, 0.3525, 0.3421, 0.0470, 0.0460),
    (150.0, 200.0, 0.7994, 0.6855, 0.3639, 0.3525, 0.0470, 0.0470),
    (200.0, 250.0, 0.9593, 0.7994, 0.3766, 0.3639, 0.0340, 0.0470),
    (250.0, 300.0, 1.2180, 0.9593, 0.3925, 0.3766, 0.0000, 0.0340),
    (300.0, 350.0, 1.2180, 1.2180, 0.3925, 0.3925, 0.0000, 0.0000),
]


def debye_huckel(tref: float, use_bdot: bool):
    if tref <= 0.01:
        A, B, Bdot = 0.4939, 0.3253, 0.0374
    elif tref > 350.0:
        A, B, Bdot = 1.2180, 0.3925, 0.0
    else:
        for (tl, th, Ah, Al, Bh, Bl, Dh, Dl) in _DEBYE:
            if tl < tref <= th:
                A = interpolate(th, tl, tref, Ah, Al)
                B = interpolate(th, tl, tref, Bh, Bl)
                Bdot = interpolate(th, tl, tref, Dh, Dl)
                break
    if not use_bdot:
        Bdot = 0.0
    return A, B, Bdot


def fit_logK_coef(logK: List[float], temperatures: List[float]) -> List[float]:
    """reaction_aux.F90:1336-1408 (ReactionFitLogKCoef)."""
    n = len(temperatures)
    vec = [[0.0] * n for _ in range(5)]
    for i in range(n):
        tk = temperatures[i] + 273.15
        vec[0][i] = math.log(tk)
        vec[1][i] = 1.0
        vec[2][i] = tk
        vec[3][i] = 1.0 / tk
        vec[4][i] = 1.0 / (tk * tk)
    coefs = [0.0] * 5
    use = [1] * n
    for j in range(5):
        coefs[j] = 0.0
        for i in range(n):
            if abs(logK[i] - 500.0) < 1.0e-10:
                use[i] = 0
            else:
                coefs[j] = coefs[j] + vec[j][i] * logK[i]
                use[i] = 1
    a = [[0.0] * 5 for _ in range(5)]
    for j in range(5):
        for k in range(j, 5):
            a[j][k] = 0.0
            for i in range(n):
                if use[i] == 1:
                    a[j][k] = a[j][k] + vec[j][i] * vec[k][i]
            if j != k:
                a[k][j] = a[j][k]
    indx = ludcmp(a, 5)
    lubksb(a, 5, indx, coefs)
    return coefs


def interpolate_logK(coefs, temp):
    """reaction_aux.F90:1461-1488"""
    tk = temp + 273.15
    return (coefs[0] * math.log(tk) + coefs[1] + coefs[2] * tk + coefs[3] / tk
            + coefs[4] / (tk * tk))


def interpolate_logK_hpt(c, temp, pres):
    """reaction_aux.F90:1529-1571"""
    tk = temp + 273.15
    tr = tk / 273.15
    pr = pres / 1.0e7
    logtr = math.log(tr) / math.log(10.0)
    return (c[0] + c[1] * tr + c[2] / tr + c[3] * logtr + c[4] * tr * tr + c[5] / tr / tr
            + c[6] * math.sqrt(tr) + c[7] * pr + c[8] * pr * tr + c[9] * pr / tr
            + c[10] * pr * logtr + c[11] / pr + c[12] / pr * tr + c[13] / pr / tr
            + c[14] * pr * pr + c[15] * pr * pr * tr + c[16] * pr * pr / tr)


def _sub_species(name: str, sec: DbRxn, tgt: DbRxn, mineral_variant: bool):
    """BasisSubSpeciesInMineralRxn / BasisSubSpeciesInGasOrSecRxn
    (reaction_database_aux.F90:355-545): replace species `name` in tgt by sec's
    reaction."""
    tempnames = [''] * 20
    tempstoich = [0.0] * 20
    scale = 1.0
    tempcount = 0
    nspec0 = len(tgt.spec_name)
    for i in range(nspec0):
        if tgt.spec_name[i] != name:
            tempnames[tempcount] = tgt.spec_name[i]
            tempstoich[tempcount] = tgt.stoich[i]
            tempcount += 1
        else:
            scale = tgt.stoich[i]
    for j in range(len(sec.spec_name)):
        found = False
        # the mineral variant searches the first nspec0 slots, the other one tempcount slots
        nsearch = nspec0 if mineral_variant else tempcount
        for i in range(nsearch):
            if tempnames[i] == sec.spec_name[j]:
                tempstoich[i] = tempstoich[i] + scale * sec.stoich[j]
                found = True
                break
        if not found:
            tempnames[tempcount] = sec.spec_name[j]
            tempstoich[tempcount] = scale * sec.stoich[j]
            tempcount += 1
    names, st = [], []
    for i in range(tempcount):
        if abs(tempstoich[i]) > 1.0e-10:
            names.append(tempnames[i])
            st.append(tempstoich[i])
    tgt.spec_name = names
    tgt.stoich = st
    tgt.logK = [tgt.logK[i] + scale * sec.logK[i] for i in range(len(tgt.logK))]


def _align(basis_names: List[str], rxn: DbRxn, who: str):
    """BasisAlignSpeciesInRxn (reaction_database_aux.F90:278-351) -> spec_ids (1-based)."""
    stoich_new = [0.0] * len(basis_names)
    for n, s in zip(rxn.spec_name, rxn.stoich):
        if n not in basis_names:
            raise RuntimeError('%s not found in basis (BasisAlignSpeciesInRxn) for species %s'
                               % (n, who))
        stoich_new[basis_names.index(n)] = s
    names, st, ids = [], [], []
    for i, s in enumerate(stoich_new):
        if abs(s) > 1.0e-40:
            names.append(basis_names[i]); st.append(s); ids.append(i + 1)
    if len(names) != len(rxn.spec_name):
        raise RuntimeError('Number of reaction species does not match original: %s' % who)
    rxn.spec_name, rxn.stoich = names, st
    return ids


def _pack(rxn: DbRxn, ids: List[int]):
    """split H2O (basis id 1) out of a reaction; shift ids (:1683-1702)."""
    sid, sst = [], []
    h2oid, h2ost = 0, 0.0
    for i, s in zip(ids, rxn.stoich):
        if i != 1:
            sid.append(i - 1); sst.append(s)
        else:
            h2oid, h2ost = 1, s
    return sid, sst, h2oid, h2ost


def build_tables(deck: dk.Deck, database_path: str = None, isothermal: bool = True) -> ReactionTables:
    chem = deck.chemistry
    if database_path is None:
        database_path = os.path.normpath(os.path.join(os.path.dirname(deck.path), chem.database))
    tref = deck.reference_temperature
    pref = deck.reference_pressure
    hpt = chem.use_geothermal_hpt

    # master list of surface complexes in order of first appearance
    srf_names: List[str] = []
    for rxn in chem.srfcplx_rxns:
        for c in rxn.complexes:
            if c not in srf_names:
                srf_names.append(c)

    db = read_database(database_path, chem.primary_species, chem.secondary_species,
                       chem.gas_species, chem.minerals, srf_names, hpt=hpt)
    num_logKs = db.num_logKs

    debyeA, debyeB, debyeBdot = debye_huckel(tref, chem.act_coef_use_bdot)

    itl = ith = 0
    tl = th = 0.0
    if not hpt:
        T = db.temperatures
        if tref <= T[0]:
            itl = ith = 0
        elif tref > T[-1]:
            itl = ith = len(T) - 1
        else:
            for it in range(len(T) - 1):
                itl, ith = it, it + 1
                if T[itl] < tref <= T[ith]:
                    break
        tl, th = T[itl], T[ith]

    def logK_at_ref(logKs):
        if hpt:
            return interpolate_logK_hpt(logKs, tref, pref), list(logKs)
        if isothermal:
            return interpolate(th, tl, tref, logKs[ith], logKs[itl]), list(logKs)
        coefs = fit_logK_coef(logKs, db.temperatures)
        return interpolate_logK(coefs, tref), coefs

    naq = len(chem.primary_species)
    ncplx = len(chem.secondary_species)
    ngas = len(chem.gas_species)
    ncomp_h2o = naq + 1
    pri_names = [H2O_NAME] + list(chem.primary_species)
    sec_names = list(chem.secondary_species)
    gas_names = list(chem.gas_species)

    for n in chem.redox_species:
        if n not in db.primary:
            raise RuntimeError('Redox species "%s" not found among primary species.' % n)
        db.primary[n].dbaserxn = None

    nsec = ncplx + ngas
    rxn_owners = ([db.primary[n] for n in chem.primary_species if db.primary[n].dbaserxn] +
                  [db.secondary[n] for n in sec_names if db.secondary[n].dbaserxn] +
                  [db.gases[n] for n in gas_names if db.gases[n].dbaserxn])
    if len(rxn_owners) != nsec:
        raise RuntimeError('Too %s reactions read from database for number of secondary species '
                           'defined.' % ('few' if len(rxn_owners) < nsec else 'many'))

    def basis_id(name):
        if name in pri_names:
            return pri_names.index(name) + 1
        if name in sec_names:
            return -(sec_names.index(name) + 1)
        if name in gas_names:
            return -(ncplx + gas_names.index(name) + 1)
        raise RuntimeError('Species %s not found among primary, secondary, or gas species.' % name)

    sec_stoich: Dict[str, DbRxn] = {}
    if nsec > 0:
        pri_matrix = [[0.0] * ncomp_h2o for _ in range(nsec)]
        sec_matrix = [[0.0] * nsec for _ in range(nsec)]
        logKvector = [[0.0] * nsec for _ in range(num_logKs)]
        for icount, sp in enumerate(rxn_owners):
            for k in range(num_logKs):
                logKvector[k][icount] = sp.dbaserxn.logK[k]
            i = basis_id(sp.name)
            if i > 0:
                pri_matrix[icount][i - 1] = -1.0
            else:
                sec_matrix[icount][-i - 1] = -1.0
            for st, nm in zip(sp.dbaserxn.stoich, sp.dbaserxn.spec_name):
                i = basis_id(nm)
                if i > 0:
                    pri_matrix[icount][i - 1] = st
                else:
                    sec_matrix[icount][-i - 1] = st
        indices = ludcmp(sec_matrix, nsec)
        inv = [[0.0] * nsec for _ in range(nsec)]
        for ispec in range(nsec):
            unit = [0.0] * nsec
            unit[ispec] = 1.0
            lubksb(sec_matrix, nsec, indices, unit)
            for r in range(nsec):
                inv[r][ispec] = unit[r]
        stoich_matrix = [[0.0] * ncomp_h2o for _ in range(nsec)]
        for j in range(ncomp_h2o):
            for i in range(nsec):
                for ispec in range(nsec):
                    stoich_matrix[i][j] = stoich_matrix[i][j] + inv[i][ispec] * pri_matrix[ispec][j]
        for i in range(nsec):
            for j in range(ncomp_h2o):
                stoich_matrix[i][j] = -1.0 * stoich_matrix[i][j]
        logK_sw = [[0.0] * nsec for _ in range(num_logKs)]
        for j in range(nsec):
            for i in range(num_logKs):
                dot = 0.0
                for k in range(nsec):
                    dot = dot + inv[j][k] * logKvector[i][k]
                logK_sw[i][j] = logK_sw[i][j] - dot
        for n in chem.primary_species:
            db.primary[n].dbaserxn = None
        for icount, n in enumerate(sec_names + gas_names):
            names, st, ids = [], [], []
            for icol in range(ncomp_h2o):
                if abs(stoich_matrix[icount][icol]) > 1.0e-40:
                    names.append(pri_names[icol]); st.append(stoich_matrix[icount][icol])
                    ids.append(icol + 1)
            rx = DbRxn(names, st, [logK_sw[k][icount] for k in range(num_logKs)])
            rx.spec_ids = ids
            sec_stoich[n] = rx
            (db.secondary if icount < ncplx else db.gases)[n].dbaserxn = rx

    kinetic = list(chem.mineral_kinetics.keys())
    # substitute gases then secondary aqueous species (:1392-1518)
    for sub in gas_names + sec_names:
        subrxn = sec_stoich[sub]
        for mn in chem.minerals:
            # The snapshot substitutes only into minerals with a kinetic rate law (`associated(cur_mineral%tstrxn)`,
            # reaction_database.F90:1401,1467) and would then stop in BasisAlignSpeciesInRxn for a non-kinetic mineral written
            # in a secondary species (Goethite / Fe+++ in example_problems/ascem_chemistry).  That deck's own pflotran.out
            # ("Final Basis", :4886) shows the substitution applied to every mineral: followed here (identical for kinetic ones).
            m = db.minerals[mn]
            while sub in m.dbaserxn.spec_name:
                _sub_species(sub, subrxn, m.dbaserxn, True)
        for sn in srf_names:
            s = db.srfcplx[sn]
            while sub in s.dbaserxn.spec_name:
                _sub_species(sub, subrxn, s.dbaserxn, False)

    for mn in chem.minerals:
        m = db.minerals[mn]
        m.dbaserxn.spec_ids = _align(pri_names, m.dbaserxn, mn)
    for sn in srf_names:
        s = db.srfcplx[sn]
        s.dbaserxn.spec_ids = _align(pri_names, s.dbaserxn, sn)

    t = ReactionTables()
    t.source_deck = os.path.basename(deck.path)
    t.database = os.path.basename(database_path)
    t.reference_temperature = float(tref)
    t.reference_pressure = float(pref)
    t.reference_water_density = water_density_ifc67(tref, pref) if deck.reference_density is None else float(deck.reference_density)
    t.naqcomp = naq
    t.ncomp = naq
    t.primary_species_names = list(chem.primary_species)
    t.primary_spec_Z = np.array([db.primary[n].Z for n in chem.primary_species])
    t.primary_spec_a0 = np.array([db.primary[n].a0 for n in chem.primary_species])
    t.primary_spec_molar_wt = np.array([db.primary[n].molar_weight for n in chem.primary_species])
    t.debyeA, t.debyeB, t.debyeBdot = debyeA, debyeB, debyeBdot
    t.use_log_formulation = int(chem.use_log_formulation)
    t.act_coef_update_frequency = chem.act_coef_update_frequency
    t.act_coef_update_algorithm = chem.act_coef_update_algorithm
    t.use_activity_h2o = int(chem.use_activity_h2o)
    t.initialize_with_molality = int(chem.initialize_with_molality)
    t.max_dlnC = chem.max_dlnC
    t.max_relative_change_tolerance = chem.max_relative_change_tolerance
    t.max_residual_tolerance = chem.max_residual_tolerance
    t.unsupported = list(chem.unsupported)
    # 0: logK fixed at reference T (isothermal); 1: 5-term fit per cell T; 2: hpt 17-term per cell T,P
    t.logK_mode = 2 if hpt else (0 if isothermal else 1)
    t.num_logK_coef = num_logKs if (hpt or isothermal) else 5
    t.dbase_temperatures = np.array(db.temperatures, dtype=np.float64)

    def pack_group(names, getrxn):
        ids_rows, st_rows, h2oid, h2ost, logK, coef = [], [], [], [], [], []
        for n in names:
            rx = getrxn(n)
            sid, sst, hid, hst = _pack(rx, rx.spec_ids)
            ids_rows.append(sid); st_rows.append(sst); h2oid.append(hid); h2ost.append(hst)
            lk, cf = logK_at_ref(rx.logK)
            logK.append(lk); coef.append(cf)
        return ids_rows, st_rows, h2oid, h2ost, logK, coef

    # --- aqueous complexes
    t.neqcplx = ncplx
    t.secondary_species_names = sec_names
    ids_rows, st_rows, h2oid, h2ost, logK, coef = pack_group(sec_names, lambda n: sec_stoich[n])
    mx = max([len(r) for r in ids_rows], default=0)
    # reference: max_aq_species counts H2O too (dbaserxn%nspec); keep that leading dimension
    mx_ld = max([len(sec_stoich[n].spec_name) for n in sec_names], default=0)
    t.eqcplxspecid = idarray(ids_rows, mx_ld + 1)
    t.eqcplxstoich = starray0(st_rows, mx_ld + 1)
    t.eqcplxh2oid = np.array(h2oid, dtype=np.int32)
    t.eqcplxh2ostoich = np.array(h2ost, dtype=np.float64)
    t.eqcplx_logK = np.array(logK, dtype=np.float64)
    t.eqcplx_logKcoef = np.array(coef, dtype=np.float64).reshape(ncplx, t.num_logK_coef)
    t.eqcplx_Z = np.array([db.secondary[n].Z for n in sec_names], dtype=np.float64)
    t.eqcplx_a0 = np.array([db.secondary[n].a0 for n in sec_names], dtype=np.float64)
    t.eqcplx_molar_wt = np.array([db.secondary[n].molar_weight for n in sec_names])

    # --- passive gases (constraints only)
    t.npassive_gas = ngas
    t.passive_gas_names = gas_names
    ids_rows, st_rows, h2oid, h2ost, logK, coef = pack_group(gas_names, lambda n: sec_stoich[n])
    mx_ld = max([len(sec_stoich[n].spec_name) for n in gas_names], default=0)
    t.paseqspecid = idarray(ids_rows, mx_ld + 1)
    t.paseqstoich = starray0(st_rows, mx_ld + 1)
    t.paseqh2oid = np.array(h2oid, dtype=np.int32)
    t.paseqh2ostoich = np.array(h2ost, dtype=np.float64)
    t.paseqlogK = np.array(logK, dtype=np.float64)

    # --- minerals (all: constraint equilibrium; kinetic: rates)
    t.nmnrl = len(chem.minerals)
    t.mineral_names = list(chem.minerals)
    ids_rows, st_rows, h2oid, h2ost, logK, coef = pack_group(
        chem.minerals, lambda n: db.minerals[n].dbaserxn)
    mx_ld = max([len(db.minerals[n].dbaserxn.spec_name) for n in chem.minerals], default=0)
    t.mnrlspecid = idarray(ids_rows, mx_ld + 1)
    t.mnrlstoich = starray1(st_rows, mx_ld)
    t.mnrlh2oid = np.array(h2oid, dtype=np.int32)
    t.mnrlh2ostoich = np.array(h2ost, dtype=np.float64)
    t.mnrl_logK = np.array(logK, dtype=np.float64)

    kin = [n for n in chem.minerals if n in chem.mineral_kinetics]
    nkin = len(kin)
    t.nkinmnrl = nkin
    t.kinmnrl_names = kin
    ids_rows, st_rows, h2oid, h2ost, logK, coef = pack_group(kin, lambda n: db.minerals[n].dbaserxn)
    mx_ld = max([len(db.minerals[n].dbaserxn.spec_name) for n in kin], default=0)
    t.kinmnrlspecid = idarray(ids_rows, mx_ld + 1)
    t.kinmnrlstoich = starray1(st_rows, mx_ld)
    t.kinmnrlh2oid = np.array(h2oid, dtype=np.int32)
    t.kinmnrlh2ostoich = np.array(h2ost, dtype=np.float64)
    t.kinmnrl_logK = np.array(logK, dtype=np.float64)
    t.kinmnrl_logKcoef = np.array(coef, dtype=np.float64).reshape(nkin, t.num_logK_coef)
    tst = [chem.mineral_kinetics[n] for n in kin]
    t.kinmnrl_molar_vol = np.array([db.minerals[n].molar_volume for n in kin], dtype=np.float64)
    t.kinmnrl_molar_wt = np.array([db.minerals[n].molar_weight for n in kin], dtype=np.float64)
    t.kinmnrl_affinity_threshold = np.array([x.affinity_threshold for x in tst], dtype=np.float64)
    t.kinmnrl_rate_limiter = np.array([x.rate_limiter for x in tst], dtype=np.float64)
    nprefs = [len(x.prefactors) for x in tst]
    t.kinmnrl_num_prefactors = np.array(nprefs, dtype=np.int32)
    t.kinmnrl_rate_constant = np.array(
        [x.rate if len(x.prefactors) == 0 else 0.0 for x in tst], dtype=np.float64)
    t.kinmnrl_activation_energy = np.array(
        [x.activation_energy if len(x.prefactors) == 0 else 0.0 for x in tst], dtype=np.float64)
    U = dk.UNINITIALIZED_DOUBLE
    t.has_min_scale_factor = int(any(x.min_scale_factor != U for x in tst))
    t.has_Temkin_const = int(any(x.affinity_factor_sigma != U for x in tst))
    t.has_affinity_power = int(any(x.affinity_factor_beta != U for x in tst))
    t.kinmnrl_min_scale_factor = np.array(
        [x.min_scale_factor if x.min_scale_factor != U else 1.0 for x in tst], dtype=np.float64)
    t.kinmnrl_Temkin_const = np.array(
        [x.affinity_factor_sigma if x.affinity_factor_sigma != U else 1.0 for x in tst],
        dtype=np.float64)
    t.kinmnrl_affinity_power = np.array(
        [x.affinity_factor_beta if x.affinity_factor_beta != U else 1.0 for x in tst],
        dtype=np.float64)
    maxpref = max(nprefs, default=0)
    maxprefspec = max([len(p.species) for x in tst for p in x.prefactors], default=0)
    t.max_num_prefactors = maxpref
    t.max_num_prefactor_species = maxprefspec
    # Fortran (0:maxspec, maxpref, nkin) <-> numpy [nkin, maxpref, maxspec+1]
    t.kinmnrl_pref_rate = np.zeros((nkin, max(maxpref, 1)), dtype=np.float64)
    t.kinmnrl_pref_activation_energy = np.zeros((nkin, max(maxpref, 1)), dtype=np.float64)
    t.kinmnrl_prefactor_id = np.zeros((nkin, max(maxpref, 1), maxprefspec + 1), dtype=np.int32)
    t.kinmnrl_pref_alpha = np.zeros((nkin, max(maxpref, 1), max(maxprefspec, 1)), dtype=np.float64)
    t.kinmnrl_pref_beta = np.zeros_like(t.kinmnrl_pref_alpha)
    t.kinmnrl_pref_atten_coef = np.zeros_like(t.kinmnrl_pref_alpha)
    lower_pri = [n.lower() for n in chem.primary_species]
    lower_sec = [n.lower() for n in sec_names]
    for im, x in enumerate(tst):
        for ip, p in enumerate(x.prefactors):
            t.kinmnrl_pref_rate[im, ip] = p.rate
            t.kinmnrl_pref_activation_energy[im, ip] = p.activation_energy
            t.kinmnrl_prefactor_id[im, ip, 0] = len(p.species)
            for js, ps in enumerate(p.species):
                nm = ps.name.lower()
                if nm in lower_pri:
                    pid = lower_pri.index(nm) + 1
                elif nm in lower_sec:
                    pid = -(lower_sec.index(nm) + 1)
                else:
                    raise RuntimeError('Kinetic mineral prefactor species "%s" not found' % ps.name)
                t.kinmnrl_prefactor_id[im, ip, js + 1] = pid
                t.kinmnrl_pref_alpha[im, ip, js] = ps.alpha
                t.kinmnrl_pref_beta[im, ip, js] = ps.beta
                t.kinmnrl_pref_atten_coef[im, ip, js] = ps.attenuation_coef

    # --- surface complexation
    nsrf = len(srf_names)
    t.nsrfcplx = nsrf
    t.srfcplx_names = srf_names
    ids_rows, st_rows, h2oid, h2ost, logK, coef = pack_group(
        srf_names, lambda n: db.srfcplx[n].dbaserxn)
    mx_ld = max([len(db.srfcplx[n].dbaserxn.spec_name) for n in srf_names], default=0)
    t.srfcplxspecid = idarray(ids_rows, mx_ld + 1)
    t.srfcplxstoich = starray1(st_rows, mx_ld)
    t.srfcplxh2oid = np.array(h2oid, dtype=np.int32)
    t.srfcplxh2ostoich = np.array(h2ost, dtype=np.float64)
    t.srfcplx_free_site_stoich = np.array([db.srfcplx[n].free_site_stoich for n in srf_names],
                                          dtype=np.float64)
    t.srfcplx_logK = np.array(logK, dtype=np.float64)
    t.srfcplx_logKcoef = np.array(coef, dtype=np.float64).reshape(nsrf, t.num_logK_coef)
    t.srfcplx_Z = np.array([db.srfcplx[n].Z for n in srf_names], dtype=np.float64)

    rxns = chem.srfcplx_rxns
    nrxn = len(rxns)
    t.nsrfcplxrxn = nrxn
    mxc = max([len(r.complexes) for r in rxns], default=0)
    t.srfcplxrxn_site_names = [r.free_site_name for r in rxns]
    t.srfcplxrxn_surf_type = np.array([r.surface_itype for r in rxns], dtype=np.int32)
    to_surf = []
    for r in rxns:
        if r.surface_itype == dk.MINERAL_SURFACE:
            if r.surface_name not in kin:
                raise RuntimeError('Mineral %s listed in surface complexation reaction not found '
                                   'in kinetic mineral list' % r.surface_name)
            to_surf.append(kin.index(r.surface_name) + 1)
        else:
            to_surf.append(0)
    t.srfcplxrxn_to_surf = np.array(to_surf, dtype=np.int32)
    t.srfcplxrxn_site_density = np.array([r.site_density for r in rxns], dtype=np.float64)
    t.srfcplxrxn_to_complex = idarray([[srf_names.index(c) + 1 for c in r.complexes] for r in rxns],
                                      mxc + 1)
    t.srfcplxrxn_stoich_flag = np.array(
        [int(any(db.srfcplx[c].free_site_stoich > 1.0 for c in r.complexes)) for r in rxns],
        dtype=np.int32)
    eq = [i + 1 for i, r in enumerate(rxns)
          if r.itype in (dk.SRFCMPLX_RXN_NULL, dk.SRFCMPLX_RXN_EQUILIBRIUM)]
    mr = [i + 1 for i, r in enumerate(rxns) if r.itype == dk.SRFCMPLX_RXN_MULTIRATE_KINETIC]
    kn = [i + 1 for i, r in enumerate(rxns) if r.itype == dk.SRFCMPLX_RXN_KINETIC]
    t.neqsrfcplxrxn = len(eq)
    t.eqsrfcplxrxn_to_srfcplxrxn = np.array(eq, dtype=np.int32)
    t.nkinmrsrfcplxrxn = len(mr)
    t.kinmrsrfcplxrxn_to_srfcplxrxn = np.array(mr, dtype=np.int32)
    t.nkinsrfcplxrxn = len(kn)
    mxr = max([len(rxns[i - 1].rates) for i in mr], default=0)
    t.kinmr_max_nrate = mxr
    t.kinmr_nrate = np.array([mxr] + [len(rxns[i - 1].rates) for i in mr], dtype=np.int32)
    t.kinmr_rate = starray1([rxns[i - 1].rates for i in mr], mxr)
    t.kinmr_frac = starray1([rxns[i - 1].site_fractions for i in mr], mxr)
    # kinetic surface complexation (reaction_database.F90:2568-2655): rates stored per (complex position in its reaction,
    # kinetic reaction); the backward rate is the deck's (the `Uninitialized` test at :2633 looks at the freshly zeroed table,
    # so the Kb = Kf * Keq branch is never taken)
    t.nkinsrfcplx = sum(len(rxns[i - 1].complexes) for i in kn)
    t.kinsrfcplxrxn_to_srfcplxrxn = np.array(kn, dtype=np.int32)
    mxk = max([len(rxns[i - 1].complexes) for i in kn], default=0)
    t.kinsrfcplx_forward_rate = np.zeros((len(kn), max(mxk, 1)), dtype=np.float64)
    t.kinsrfcplx_backward_rate = np.zeros((len(kn), max(mxk, 1)), dtype=np.float64)
    for ik, i in enumerate(kn):
        for ic, cname in enumerate(rxns[i - 1].complexes):
            ck = rxns[i - 1].complex_kinetics.get(cname, {})
            t.kinsrfcplx_forward_rate[ik, ic] = ck.get('forward', 0.0)
            t.kinsrfcplx_backward_rate[ik, ic] = ck.get('backward', -999.0)     # UNINITIALIZED_DOUBLE when the deck omits it

    # --- ion exchange
    nix = len(chem.ionx_rxns)
    t.neqionxrxn = nix
    mxcat = max([len(r.cations) for r in chem.ionx_rxns], default=0)
    cat_ids = []
    for r in chem.ionx_rxns:
        row = []
        for (n, k) in r.cations:
            if n not in chem.primary_species:
                raise RuntimeError('Cation %s in ion exchange reaction not found in swapped basis.' % n)
            row.append(chem.primary_species.index(n) + 1)
        cat_ids.append(row)
    t.eqionx_rxn_cationid = idarray(cat_ids, mxcat + 1)
    t.eqionx_rxn_k = starray1([[k for (_, k) in r.cations] for r in chem.ionx_rxns], mxcat)
    t.eqionx_rxn_CEC = np.array([r.CEC for r in chem.ionx_rxns], dtype=np.float64)
    zflag, ixsurf = [], []
    for r, ids in zip(chem.ionx_rxns, cat_ids):
        found = False
        for i in ids:
            for j in ids:
                if abs(t.primary_spec_Z[i - 1] - t.primary_spec_Z[j - 1]) > 0.1:
                    found = True
        zflag.append(int(found))
        ixsurf.append(kin.index(r.mineral_name) + 1 if len(r.mineral_name) > 1 else 0)
    t.eqionx_rxn_Z_flag = np.array(zflag, dtype=np.int32)
    t.eqionx_rxn_to_surf = np.array(ixsurf, dtype=np.int32)

    # --- KD isotherms
    nkd = len(chem.kd_rxns)
    t.neqkdrxn = nkd
    if nkd > 0 and ncplx > 0:
        raise RuntimeError('Isotherm reactions currently calculated as a function of free-ion, '
                           'not totals.')
    t.eqkdspecid = np.array([chem.primary_species.index(r.species_name) + 1 for r in chem.kd_rxns],
                            dtype=np.int32)
    t.eqkdtype = np.array([r.itype for r in chem.kd_rxns], dtype=np.int32)
    t.eqkddistcoef = np.array([r.Kd for r in chem.kd_rxns], dtype=np.float64)
    t.eqkdlangmuirb = np.array([r.Langmuir_b for r in chem.kd_rxns], dtype=np.float64)
    t.eqkdfreundlichn = np.array([r.Freundlich_n for r in chem.kd_rxns], dtype=np.float64)

    # --- radioactive decay and general reactions (reaction_database.F90:2915-3135): the REACTION string gives the species in
    # the order written, reactants negative (database_rxn from DatabaseRxnCreateFromRxnString, reaction_database_aux.F90:59-274)
    def rxn_from_string(text):
        names, st = [], []
        negative, value, right = False, None, False
        for w in text.split():
            if w == '+':
                continue
            if w == '-':
                negative = not negative
                continue
            if w in ('=', '<=>', '<->'):
                right = True
                continue
            if not w[0].isalpha():
                value = fnum_py(w)
                continue
            if w.upper() == 'H2O':
                value, negative = None, False
                continue
            v = 1.0 if value is None else value
            if negative:
                v = -v
            if not right:
                v = -v
            if w not in chem.primary_species and w not in chem.immobile_species:
                raise RuntimeError('Species %s in reaction "%s" not found among primary species' % (w, text))
            names.append(w); st.append(v)
            value, negative = None, False
        # immobile species are dofs naqcomp + i (offset_immobile = naqcomp, reaction.F90:105-106)
        return [chem.primary_species.index(n) + 1 if n in chem.primary_species else naq + chem.immobile_species.index(n) + 1
                for n in names], st

    def fnum_py(tok):
        return float(tok.replace('d', 'e').replace('D', 'e'))

    g_ids, g_st, gf_ids, gf_st, gb_ids, gb_st = [], [], [], [], [], []
    for r in chem.general_rxns:
        ids, st = rxn_from_string(r.reaction)
        g_ids.append(ids); g_st.append(st)
        gf_ids.append([i for i, v in zip(ids, st) if v < 0.0]); gf_st.append([abs(v) for v in st if v < 0.0])
        gb_ids.append([i for i, v in zip(ids, st) if v > 0.0]); gb_st.append([v for v in st if v > 0.0])
    t.ngeneral_rxn = len(chem.general_rxns)
    mg = max([len(x) for x in g_ids], default=0)
    t.generalspecid = idarray(g_ids, mg + 1); t.generalstoich = starray1(g_st, mg)
    t.generalforwardspecid = idarray(gf_ids, mg + 1); t.generalforwardstoich = starray1(gf_st, mg)
    t.generalbackwardspecid = idarray(gb_ids, mg + 1); t.generalbackwardstoich = starray1(gb_st, mg)
    t.general_kf = np.array([r.forward_rate for r in chem.general_rxns], dtype=np.float64)
    t.general_kr = np.array([r.backward_rate for r in chem.general_rxns], dtype=np.float64)
    d_ids, d_st, d_fwd = [], [], []
    for r in chem.radiodecay_rxns:
        ids, st = rxn_from_string(r.reaction)
        if sum(1 for v in st if v < 0.0) > 1:
            raise RuntimeError('Cannot have more than one reactant in radioactive decay reaction: (%s).' % r.reaction)
        d_ids.append(ids); d_st.append(st)
        d_fwd.append(([i for i, v in zip(ids, st) if v < 0.0] or [0])[-1])
    t.nradiodecay_rxn = len(chem.radiodecay_rxns)
    md = max([len(x) for x in d_ids], default=0)
    t.radiodecayspecid = idarray(d_ids, md + 1); t.radiodecaystoich = starray1(d_st, md)
    t.radiodecayforwardspecid = np.array(d_fwd, dtype=np.int32)
    t.radiodecay_kf = np.array([r.rate_constant for r in chem.radiodecay_rxns], dtype=np.float64)
    t.eqkdmineral = np.array(
        [kin.index(r.kd_mineral_name) + 1 if len(r.kd_mineral_name) > 1 else 0
         for r in chem.kd_rxns], dtype=np.int32)

    # --- immobile species, immobile decay (reaction_database.F90:3336-3370) and microbial reactions (:3126-3333)
    t.nimmobile = len(chem.immobile_species)
    t.immobile_names = list(chem.immobile_species)
    t.ncomp = naq + t.nimmobile
    t.nimmobile_decay_rxn = len(chem.immobile_decay_rxns)
    for name, _ in chem.immobile_decay_rxns:
        if name not in chem.immobile_species:
            raise RuntimeError('Species "%s" in immobile decay reaction not found among immobile species.' % name)
    t.immobile_decayspecid = np.array([chem.immobile_species.index(n) + 1 for n, _ in chem.immobile_decay_rxns], dtype=np.int32)
    t.immobile_decay_rate_constant = np.array([k for _, k in chem.immobile_decay_rxns], dtype=np.float64)
    t.nmicrobial_rxn = len(chem.microbial_rxns)
    m_ids, m_st, m_mon, m_inh, bio, yld = [], [], [], [], [], []
    mon_spec, mon_K, mon_Cth, inh_type, inh_spec, inh_C, inh_C2 = [], [], [], [], [], [], []
    for r in chem.microbial_rxns:
        ids, st = rxn_from_string(r.reaction)
        names = [(chem.primary_species + chem.immobile_species)[i - 1] for i in ids]
        m_ids.append(ids); m_st.append(st)
        if r.biomass is not None:
            if r.biomass[0] not in chem.immobile_species:
                raise RuntimeError('Biomass species "%s" not found among immobile species.' % r.biomass[0])
            if r.biomass[0] in names:
                raise RuntimeError('Biomass species "%s" should not be included in microbial reaction.' % r.biomass[0])
            bio.append(chem.immobile_species.index(r.biomass[0]) + 1); yld.append(r.biomass[1])
        else:
            bio.append(0); yld.append(0.0)
        row = []
        for (name, K, Cth) in r.monod:
            if name not in names:
                raise RuntimeError('Monod species "%s" not found in microbial reaction.' % name)
            if st[names.index(name)] > 0.0:
                raise RuntimeError('Monod species "%s" must be a reactant and not a product in microbial reaction.' % name)
            mon_spec.append(chem.primary_species.index(name) + 1); mon_K.append(K); mon_Cth.append(Cth)
            row.append(len(mon_spec))
        m_mon.append(row)
        row = []
        for (name, itype, Cc, C2) in r.inhibition:
            inh_spec.append(chem.primary_species.index(name) + 1); inh_type.append(itype); inh_C.append(Cc); inh_C2.append(C2)
            row.append(len(inh_spec))
        m_inh.append(row)
    mm = max([len(x) for x in m_ids], default=0)
    t.microbial_specid = idarray(m_ids, mm + 1); t.microbial_stoich = starray1(m_st, mm)
    t.microbial_rate_constant = np.array([r.rate_constant for r in chem.microbial_rxns], dtype=np.float64)
    # allocated only when some reaction sets a positive activation energy (reaction_database.F90:3146-3148, 3168-3171)
    t.has_microbial_activation_energy = int(any(r.activation_energy > 0.0 for r in chem.microbial_rxns))
    t.microbial_activation_energy = np.array([r.activation_energy for r in chem.microbial_rxns], dtype=np.float64)
    t.microbial_biomassid = np.array(bio, dtype=np.int32); t.microbial_biomass_yield = np.array(yld, dtype=np.float64)
    t.microbial_monodid = idarray(m_mon, max([len(x) for x in m_mon], default=0) + 1)
    t.microbial_inhibitionid = idarray(m_inh, max([len(x) for x in m_inh], default=0) + 1)
    t.microbial_monod_specid = np.array(mon_spec, dtype=np.int32)
    t.microbial_monod_K = np.array(mon_K, dtype=np.float64); t.microbial_monod_Cth = np.array(mon_Cth, dtype=np.float64)
    t.microbial_inhibition_type = np.array(inh_type, dtype=np.int32); t.microbial_inhibition_specid = np.array(inh_spec, dtype=np.int32)
    t.microbial_inhibition_C = np.array(inh_C, dtype=np.float64); t.microbial_inhibition_C2 = np.array(inh_C2, dtype=np.float64)

    t.neqsorb = nix + nkd + t.neqsrfcplxrxn
    t.nsorb = t.neqsorb + t.nkinmrsrfcplxrxn + t.nkinsrfcplxrxn

    # --- species_idx (:3466-3557)
    def find_ci(word, names):
        for i, n in enumerate(names):
            if n.lower() == word.lower():
                return i + 1
        return 0
    h = find_ci('H+', chem.primary_species)
    if h == 0:
        h = -find_ci('H+', sec_names)
    t.h_ion_id = h
    t.h2o_aq_id = find_ci('H2O', chem.primary_species)
    return t
