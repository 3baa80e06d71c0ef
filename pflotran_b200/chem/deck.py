"""PFLOTRAN input-deck reader: the CHEMISTRY and CONSTRAINT cards only.

Host-side setup code.  It restates the subset of the reference's deck grammar
that the reaction hot path needs so tests and benchmarks can build the flat
chemistry tables from the reference's own decks:

  * line conventions: reference src/pflotran/input_aux.F90:642-732
    (comment lines '#'/'!', SKIP/NOSKIP) and :1236-1273 (block end '/', 'END',
    'END_*').
  * CHEMISTRY keywords: reference src/pflotran/reaction.F90:113-926.
  * MINERAL_KINETICS: reference src/pflotran/reaction_mineral.F90:78-412.
  * SURFACE_COMPLEXATION_RXN: reference src/pflotran/reaction_surf_complex.F90:30-370.
  * CONSTRAINT: reference src/pflotran/transport_constraint.F90:217-500.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from .units import units_convert_to_internal

UNINITIALIZED_DOUBLE = -999.0

# reaction_aux.F90:23-27
ACT_COEF_FREQUENCY_OFF = 0
ACT_COEF_FREQUENCY_TIMESTEP = 1
ACT_COEF_FREQUENCY_NEWTON_ITER = 2
ACT_COEF_ALGORITHM_LAG = 3
ACT_COEF_ALGORITHM_NEWTON = 4

# reaction_surf_complex_aux.F90:19-27
NULL_SURFACE, COLLOID_SURFACE, MINERAL_SURFACE, ROCK_SURFACE = 0, 1, 2, 3
SRFCMPLX_RXN_NULL, SRFCMPLX_RXN_EQUILIBRIUM, SRFCMPLX_RXN_MULTIRATE_KINETIC, SRFCMPLX_RXN_KINETIC = 0, 1, 2, 3

# pflotran_constants.F90:94-96
SORPTION_LINEAR, SORPTION_LANGMUIR, SORPTION_FREUNDLICH = 1, 2, 3

# transport_constraint.F90:20-29
CONSTRAINT_NULL = 0
CONSTRAINT_FREE = 1
CONSTRAINT_TOTAL = 2
CONSTRAINT_LOG = 3
CONSTRAINT_PH = 4
CONSTRAINT_MINERAL = 5
CONSTRAINT_GAS = 6
CONSTRAINT_CHARGE_BAL = 7
CONSTRAINT_TOTAL_SORB = 9
CONSTRAINT_SUPERCRIT_CO2 = 10


def fnum(tok: str) -> float:
    """Fortran list-directed real: 1.d-5, 1.D0, 1.e-5."""
    return float(tok.replace('d', 'e').replace('D', 'e'))


class DeckError(RuntimeError):
    pass


class LineReader:
    """InputReadPflotranString equivalent over a list of lines."""

    def __init__(self, text: str):
        self.lines = text.splitlines()
        self.pos = 0

    def next(self) -> Optional[List[str]]:
        """Next non-comment line as tokens, or None at EOF."""
        while self.pos < len(self.lines):
            raw = self.lines[self.pos].strip()
            self.pos += 1
            if not raw or raw[0] in '#!':
                continue
            toks = raw.replace('\t', ' ').split()
            first = toks[0].upper()
            if first.startswith('SKIP'):
                depth = 1
                while self.pos < len(self.lines):
                    t = self.lines[self.pos].split()
                    self.pos += 1
                    w = t[0].upper() if t else ''
                    if w.startswith('SKIP'):
                        depth += 1
                    if w.startswith('NOSK'):
                        depth -= 1
                        if depth == 0:
                            break
                continue
            if first.startswith('NOSK'):
                continue
            return toks
        return None

    @staticmethod
    def is_exit(toks: List[str]) -> bool:
        t = toks[0]
        return t[0] == '/' or (len(toks) == 1 and t == 'END') or t.startswith('END_') or \
            (t == 'END')

    def block(self):
        """Iterate token lines until the block terminator."""
        while True:
            toks = self.next()
            if toks is None or self.is_exit(toks):
                return
            yield toks

    def skip_block(self):
        for _ in self.block():
            pass

    def read_array(self, toks: List[str]) -> List[float]:
        """UtilityReadArray for an inline list with '\\' continuation."""
        vals: List[float] = []
        cur = toks
        while True:
            cont = False
            for t in cur:
                if t == '\\':
                    cont = True
                    break
                if t[0] in '!#':
                    break
                vals.append(fnum(t))
            if not cont:
                break
            cur = self.next()
            if cur is None:
                break
        return vals


@dataclass
class PrefactorSpecies:
    name: str
    alpha: float = 0.0
    beta: float = 0.0
    attenuation_coef: float = 0.0


@dataclass
class Prefactor:
    rate: float = UNINITIALIZED_DOUBLE
    activation_energy: float = UNINITIALIZED_DOUBLE
    species: List[PrefactorSpecies] = field(default_factory=list)


@dataclass
class TSTRxn:
    """transition_state_rxn_type, reaction_mineral_aux.F90"""
    rate: float = UNINITIALIZED_DOUBLE
    activation_energy: float = 0.0
    affinity_threshold: float = 0.0
    affinity_factor_beta: float = UNINITIALIZED_DOUBLE
    affinity_factor_sigma: float = UNINITIALIZED_DOUBLE
    min_scale_factor: float = UNINITIALIZED_DOUBLE
    rate_limiter: float = 0.0
    prefactors: List[Prefactor] = field(default_factory=list)


@dataclass
class SrfCplxRxn:
    itype: int = SRFCMPLX_RXN_EQUILIBRIUM
    surface_itype: int = NULL_SURFACE
    surface_name: str = ''
    free_site_name: str = ''
    site_density: float = 0.0
    complexes: List[str] = field(default_factory=list)
    rates: Optional[List[float]] = None
    site_fractions: Optional[List[float]] = None
    kinmr_scale_factor: float = 1.0
    complex_kinetics: Dict[str, Dict[str, float]] = field(default_factory=dict)


@dataclass
class IonxRxn:
    mineral_name: str = ''
    CEC: float = 0.0
    cations: List[tuple] = field(default_factory=list)  # (name, k), reference first


@dataclass
class KDRxn:
    species_name: str = ''
    itype: int = SORPTION_LINEAR
    Kd: float = 0.0
    Langmuir_b: float = 0.0
    Freundlich_n: float = 0.0
    kd_mineral_name: str = ''


@dataclass
class RateRxn:
    """GENERAL_REACTION (forward / backward rate) or RADIOACTIVE_DECAY_REACTION (rate_constant) block"""
    reaction: str = ''
    forward_rate: float = 0.0
    backward_rate: float = 0.0
    rate_constant: Optional[float] = None


@dataclass
class MicrobialRxn:
    """MICROBIAL_REACTION block (reaction_microbial.F90:30-233).  monod: (species, K, Cth); inhibition: (species, type, C, C2)"""
    reaction: str = ''
    rate_constant: float = 0.0
    activation_energy: float = 0.0
    monod: List[tuple] = field(default_factory=list)
    inhibition: List[tuple] = field(default_factory=list)
    biomass: Optional[tuple] = None            # (species name, yield)


INHIBITION_THRESHOLD, INHIBITION_THERMODYNAMIC, INHIBITION_MONOD, INHIBITION_INVERSE_MONOD = 1, 2, 3, 4   # reaction_microbial_aux.F90:13-16


@dataclass
class Chemistry:
    primary_species: List[str] = field(default_factory=list)
    secondary_species: List[str] = field(default_factory=list)
    gas_species: List[str] = field(default_factory=list)        # passive (constraint) gases
    active_gas_species: List[str] = field(default_factory=list)
    minerals: List[str] = field(default_factory=list)
    redox_species: List[str] = field(default_factory=list)
    mineral_kinetics: Dict[str, TSTRxn] = field(default_factory=dict)
    srfcplx_rxns: List[SrfCplxRxn] = field(default_factory=list)
    ionx_rxns: List[IonxRxn] = field(default_factory=list)
    kd_rxns: List[KDRxn] = field(default_factory=list)
    general_rxns: List[RateRxn] = field(default_factory=list)
    radiodecay_rxns: List[RateRxn] = field(default_factory=list)
    immobile_species: List[str] = field(default_factory=list)
    immobile_decay_rxns: List[tuple] = field(default_factory=list)     # (species name, rate constant 1/s)
    microbial_rxns: List[MicrobialRxn] = field(default_factory=list)
    database: str = ''
    use_log_formulation: bool = False
    use_geothermal_hpt: bool = False
    act_coef_update_frequency: int = ACT_COEF_FREQUENCY_OFF
    act_coef_update_algorithm: int = ACT_COEF_ALGORITHM_LAG
    act_coef_use_bdot: bool = True
    use_activity_h2o: bool = False
    initialize_with_molality: bool = False
    max_dlnC: float = 5.0
    max_relative_change_tolerance: float = 1.0e-6
    max_residual_tolerance: float = 1.0e-12
    unsupported: List[str] = field(default_factory=list)


@dataclass
class Constraint:
    name: str
    # per line, in deck order
    names: List[str] = field(default_factory=list)
    conc: List[float] = field(default_factory=list)
    ctype: List[int] = field(default_factory=list)
    aux: List[str] = field(default_factory=list)
    free_ion_guess: Optional[Dict[str, float]] = None
    minerals: Dict[str, tuple] = field(default_factory=dict)  # name -> (volfrac, area m^2/m^3)
    immobile: Dict[str, float] = field(default_factory=dict)  # name -> concentration [mol/m^3 bulk]


@dataclass
class Deck:
    chemistry: Chemistry
    constraints: Dict[str, Constraint]
    porosity: Optional[float] = None
    reference_temperature: float = 25.0
    reference_pressure: float = 101325.0
    path: str = ''
    rock_density: Optional[float] = None         # ROCK_DENSITY of the first MATERIAL_PROPERTY -> material_auxvar%soil_particle_density
    reference_density: Optional[float] = None    # REFERENCE_DENSITY (factory_subsurface.F90:1734); None: IFC-67 at the reference T, P (:995)


def _read_names(rd: LineReader) -> List[str]:
    return [t[0] for t in rd.block()]


def _read_rate(toks, what) -> float:
    """RATE_CONSTANT x [units]; negative x means 10**x; default internal unit
    mol/m^2-sec (reaction_mineral.F90:151-167)."""
    rate = fnum(toks[1])
    if rate < 0.0:
        rate = 10.0 ** rate
    if len(toks) > 2 and toks[2][0] not in '!#':
        rate = rate * units_convert_to_internal(toks[2], 'mol/m^2-sec')
    return rate


def _read_mineral_kinetics(rd: LineReader, chem: Chemistry):
    for toks in rd.block():
        name = toks[0]
        if name not in chem.minerals:
            raise DeckError('Mineral "%s" specified under MINERAL_KINETICS not found' % name)
        tst = TSTRxn()
        for t in rd.block():
            kw = t[0]
            if kw == 'RATE_CONSTANT':
                tst.rate = _read_rate(t, name)
            elif kw == 'ACTIVATION_ENERGY':
                tst.activation_energy = fnum(t[1])
                if len(t) > 2 and t[2][0] not in '!#':
                    tst.activation_energy *= units_convert_to_internal(t[2], 'J/mol')
            elif kw == 'AFFINITY_THRESHOLD':
                tst.affinity_threshold = fnum(t[1])
            elif kw == 'AFFINITY_POWER':
                tst.affinity_factor_beta = fnum(t[1])
            elif kw == 'MINERAL_SCALE_FACTOR':
                tst.min_scale_factor = fnum(t[1])
            elif kw == 'TEMKIN_CONSTANT':
                tst.affinity_factor_sigma = fnum(t[1])
            elif kw == 'RATE_LIMITER':
                tst.rate_limiter = fnum(t[1])
            elif kw == 'PREFACTOR':
                pf = Prefactor()
                for p in rd.block():
                    if p[0] == 'RATE_CONSTANT':
                        pf.rate = _read_rate(p, name)
                    elif p[0] == 'ACTIVATION_ENERGY':
                        pf.activation_energy = fnum(p[1])
                        if len(p) > 2 and p[2][0] not in '!#':
                            pf.activation_energy *= units_convert_to_internal(p[2], 'J/mol')
                    elif p[0] == 'PREFACTOR_SPECIES':
                        ps = PrefactorSpecies(p[1])
                        for s in rd.block():
                            if s[0] == 'ALPHA':
                                ps.alpha = fnum(s[1])
                            elif s[0] == 'BETA':
                                ps.beta = fnum(s[1])
                            elif s[0] == 'ATTENUATION_COEF':
                                ps.attenuation_coef = fnum(s[1])
                            else:
                                raise DeckError('PREFACTOR_SPECIES keyword ' + s[0])
                        pf.species.append(ps)
                    else:
                        raise DeckError('PREFACTOR keyword ' + p[0])
                tst.prefactors.append(pf)
            elif kw in ('SURFACE_AREA_POROSITY_POWER', 'SURFACE_AREA_VOL_FRAC_POWER',
                        'ARMOR_MINERAL', 'ARMOR_PWR', 'ARMOR_CRIT_VOL_FRAC'):
                chem.unsupported.append('MINERAL_KINETICS,' + kw)
            else:
                raise DeckError('MINERAL_KINETICS keyword ' + kw)
        # reaction_mineral.F90:352-369: inner prefactor defaults to the outer values
        for pf in tst.prefactors:
            if pf.rate == UNINITIALIZED_DOUBLE:
                pf.rate = tst.rate
                if pf.rate == UNINITIALIZED_DOUBLE:
                    raise DeckError('prefactor rate constants uninitialized for ' + name)
            if pf.activation_energy == UNINITIALIZED_DOUBLE:
                pf.activation_energy = tst.activation_energy
        chem.mineral_kinetics[name] = tst


def _read_srfcplx_rxn(rd: LineReader, chem: Chemistry):
    rxn = SrfCplxRxn()
    for toks in rd.block():
        kw = toks[0].upper()
        if kw == 'EQUILIBRIUM':
            rxn.itype = SRFCMPLX_RXN_EQUILIBRIUM
        elif kw == 'MULTIRATE_KINETIC':
            rxn.itype = SRFCMPLX_RXN_MULTIRATE_KINETIC
        elif kw == 'KINETIC':
            rxn.itype = SRFCMPLX_RXN_KINETIC
        elif kw == 'COMPLEX_KINETICS':
            for c in rd.block():
                d = {}
                for r in rd.block():
                    if r[0] == 'FORWARD_RATE_CONSTANT':
                        d['forward'] = fnum(r[1])
                    elif r[0] == 'BACKWARD_RATE_CONSTANT':
                        d['backward'] = fnum(r[1])
                    else:
                        raise DeckError('COMPLEX_KINETICS keyword ' + r[0])
                rxn.complex_kinetics[c[0]] = d
        elif kw in ('RATE', 'RATES'):
            rxn.itype = SRFCMPLX_RXN_MULTIRATE_KINETIC
            rxn.rates = rd.read_array(toks[1:])
        elif kw == 'SITE_FRACTION':
            rxn.site_fractions = rd.read_array(toks[1:])
        elif kw == 'MULTIRATE_SCALE_FACTOR':
            rxn.kinmr_scale_factor = fnum(toks[1])
        elif kw == 'MINERAL':
            rxn.surface_itype = MINERAL_SURFACE
            rxn.surface_name = toks[1]
        elif kw == 'ROCK_DENSITY':
            rxn.surface_itype = ROCK_SURFACE
        elif kw == 'COLLOID':
            rxn.surface_itype = COLLOID_SURFACE
            rxn.surface_name = toks[1]
            chem.unsupported.append('COLLOID surface')
        elif kw == 'SITE':
            rxn.free_site_name = toks[1]
            rxn.site_density = fnum(toks[2])
        elif kw == 'COMPLEXES':
            rxn.complexes = _read_names(rd)
        else:
            raise DeckError('SURFACE_COMPLEXATION_RXN keyword ' + kw)
    if rxn.itype == SRFCMPLX_RXN_MULTIRATE_KINETIC:
        # reaction_surf_complex.F90:273-301
        if rxn.site_fractions is None and rxn.rates is not None:
            rxn.site_fractions = [1.0 / float(len(rxn.rates))] * len(rxn.rates)
        if len(rxn.rates) != len(rxn.site_fractions):
            raise DeckError('number of kinetic rates does not match site fractions')
        s = 0.0
        for i in range(len(rxn.site_fractions)):
            s = s + rxn.site_fractions[i]
            rxn.rates[i] = rxn.rates[i] * rxn.kinmr_scale_factor
        if abs(1.0 - s) > 1.0e-6:
            raise DeckError('site fractions do not add up to 1')
    chem.srfcplx_rxns.append(rxn)


def _read_sorption(rd: LineReader, chem: Chemistry):
    for toks in rd.block():
        kw = toks[0].upper()
        if kw == 'ISOTHERM_REACTIONS':
            for s in rd.block():
                kd = KDRxn(species_name=s[0])
                kd_units = ''
                for t in rd.block():
                    k = t[0].upper()
                    kd.itype = SORPTION_LINEAR  # reaction.F90:529 (reset on every keyword)
                    if k == 'TYPE':
                        kd.itype = {'LINEAR': SORPTION_LINEAR, 'LANGMUIR': SORPTION_LANGMUIR,
                                    'FREUNDLICH': SORPTION_FREUNDLICH}[t[1]]
                    elif k in ('DISTRIBUTION_COEFFICIENT', 'KD'):
                        kd.Kd = fnum(t[1])
                        if len(t) > 2 and t[2][0] not in '!#':
                            kd_units = t[2]
                    elif k == 'LANGMUIR_B':
                        kd.Langmuir_b = fnum(t[1])
                        kd.itype = SORPTION_LANGMUIR
                    elif k == 'FREUNDLICH_N':
                        kd.Freundlich_n = fnum(t[1])
                        kd.itype = SORPTION_FREUNDLICH
                    elif k == 'KD_MINERAL_NAME':
                        kd.kd_mineral_name = t[1]
                    else:
                        raise DeckError('ISOTHERM_REACTIONS keyword ' + k)
                if kd_units:
                    internal = 'L/kg' if kd.kd_mineral_name else 'kg/m^3'
                    kd.Kd = kd.Kd * units_convert_to_internal(kd_units, internal)
                chem.kd_rxns.append(kd)
        elif kw == 'SURFACE_COMPLEXATION_RXN':
            _read_srfcplx_rxn(rd, chem)
        elif kw == 'ION_EXCHANGE_RXN':
            ix = IonxRxn()
            for t in rd.block():
                k = t[0].upper()
                if k == 'MINERAL':
                    ix.mineral_name = t[1]
                elif k == 'CEC':
                    ix.CEC = fnum(t[1])
                elif k == 'CATIONS':
                    ref = ''
                    for c in rd.block():
                        ix.cations.append((c[0], fnum(c[1])))
                        if len(c) > 2 and c[2].upper() == 'REFERENCE':
                            ref = c[0]
                    if not ref:
                        raise DeckError('Reference cation missing in Ion Exchange reaction.')
                    # reference cation is moved to the head of the list (reaction.F90:700-722)
                    idx = [n for n, _ in ix.cations].index(ref)
                    if abs(ix.cations[idx][1] - 1.0) > 1e-40:
                        raise DeckError('Reference cation must have k = 1.d0.')
                    ix.cations.insert(0, ix.cations.pop(idx))
                else:
                    raise DeckError('ION_EXCHANGE_RXN keyword ' + k)
            chem.ionx_rxns.append(ix)
        elif kw in ('JUMPSTART_KINETIC_SORPTION', 'NO_CHECKPOINT_KINETIC_SORPTION',
                    'NO_RESTART_KINETIC_SORPTION'):
            pass
        else:
            raise DeckError('SORPTION keyword ' + kw)


def _read_chemistry(rd: LineReader) -> Chemistry:
    chem = Chemistry()
    for toks in rd.block():
        kw = toks[0].upper()
        if kw == 'PRIMARY_SPECIES':
            chem.primary_species = _read_names(rd)
        elif kw == 'SECONDARY_SPECIES':
            chem.secondary_species = _read_names(rd)
        elif kw in ('GAS_SPECIES', 'PASSIVE_GAS_SPECIES'):
            for n in _read_names(rd):
                if n not in chem.gas_species:
                    chem.gas_species.append(n)
        elif kw == 'ACTIVE_GAS_SPECIES':
            chem.active_gas_species = _read_names(rd)
            chem.unsupported.append('ACTIVE_GAS_SPECIES')
        elif kw == 'MINERALS':
            chem.minerals = _read_names(rd)
        elif kw == 'REDOX_SPECIES':
            chem.redox_species = _read_names(rd)
        elif kw == 'MINERAL_KINETICS':
            _read_mineral_kinetics(rd, chem)
        elif kw == 'SORPTION':
            _read_sorption(rd, chem)
        elif kw == 'DATABASE':
            chem.database = toks[1]
        elif kw == 'LOG_FORMULATION':
            chem.use_log_formulation = True
        elif kw == 'GEOTHERMAL_HPT':
            chem.use_geothermal_hpt = True
        elif kw == 'ACTIVITY_COEFFICIENTS':
            chem.act_coef_update_algorithm = ACT_COEF_ALGORITHM_LAG
            chem.act_coef_update_frequency = ACT_COEF_FREQUENCY_TIMESTEP
            for w in toks[1:]:
                if w[0] in '!#':
                    break
                if w == 'OFF':
                    chem.act_coef_update_frequency = ACT_COEF_FREQUENCY_OFF
                elif w == 'LAG':
                    chem.act_coef_update_algorithm = ACT_COEF_ALGORITHM_LAG
                elif w == 'NEWTON':
                    chem.act_coef_update_algorithm = ACT_COEF_ALGORITHM_NEWTON
                elif w == 'TIMESTEP':
                    chem.act_coef_update_frequency = ACT_COEF_FREQUENCY_TIMESTEP
                elif w == 'NEWTON_ITERATION':
                    chem.act_coef_update_frequency = ACT_COEF_FREQUENCY_NEWTON_ITER
                else:
                    raise DeckError('ACTIVITY_COEFFICIENTS keyword ' + w)
        elif kw == 'NO_BDOT':
            chem.act_coef_use_bdot = False
        elif kw in ('MOLAL', 'MOLALITY'):
            chem.initialize_with_molality = True
        elif kw in ('ACTIVITY_H2O', 'ACTIVITY_WATER'):
            chem.use_activity_h2o = True
        elif kw == 'MAX_DLNC':
            chem.max_dlnC = fnum(toks[1])
        elif kw in ('MAX_RELATIVE_CHANGE_TOLERANCE', 'REACTION_TOLERANCE'):
            chem.max_relative_change_tolerance = fnum(toks[1])
        elif kw == 'MAX_RESIDUAL_TOLERANCE':
            chem.max_residual_tolerance = fnum(toks[1])
        elif kw == 'OUTPUT':
            rd.skip_block()
        elif kw == 'GENERAL_REACTION':                      # reaction.F90:315-358
            r = RateRxn()
            for t in rd.block():
                k = t[0].upper()
                if k == 'REACTION':
                    r.reaction = ' '.join(t[1:])
                elif k == 'FORWARD_RATE':
                    r.forward_rate = fnum(t[1])
                elif k == 'BACKWARD_RATE':
                    r.backward_rate = fnum(t[1])
                else:
                    raise DeckError('GENERAL_REACTION keyword ' + k)
            chem.general_rxns.append(r)
        elif kw == 'RADIOACTIVE_DECAY_REACTION':            # reaction.F90:254-314
            r = RateRxn()
            for t in rd.block():
                k = t[0].upper()
                if k == 'REACTION':
                    r.reaction = ' '.join(t[1:])
                elif k == 'RATE_CONSTANT':
                    r.rate_constant = fnum(t[1])
                    if len(t) > 2 and t[2][0] not in '!#':
                        r.rate_constant = r.rate_constant * units_convert_to_internal(t[2], 'unitless/sec')
                elif k == 'HALF_LIFE':
                    hl = fnum(t[1])
                    if len(t) > 2 and t[2][0] not in '!#':
                        hl = hl * units_convert_to_internal(t[2], 'sec')
                    r.rate_constant = -1.0 * math.log(0.5) / hl
                else:
                    raise DeckError('RADIOACTIVE_DECAY_REACTION keyword ' + k)
            if r.rate_constant is None:
                raise DeckError('RATE_CONSTANT or HALF_LIFE must be set in RADIOACTIVE_DECAY_REACTION.')
            chem.radiodecay_rxns.append(r)
        elif kw == 'IMMOBILE_SPECIES':                      # reaction_immobile.F90:30-77
            for t in rd.block():
                chem.immobile_species.append(t[0])
        elif kw == 'IMMOBILE_DECAY_REACTION':               # reaction_immobile.F90:81-160
            name, k = '', None
            for t in rd.block():
                key = t[0].upper()
                if key == 'SPECIES_NAME':
                    name = t[1]
                elif key == 'RATE_CONSTANT':
                    k = fnum(t[1])
                    if len(t) > 2 and t[2][0] not in '!#':
                        k = k * units_convert_to_internal(t[2], '1/sec')
                elif key == 'HALF_LIFE':
                    hl = fnum(t[1])
                    if len(t) > 2 and t[2][0] not in '!#':
                        hl = hl * units_convert_to_internal(t[2], 'sec')
                    k = -1.0 * math.log(0.5) / hl
                else:
                    raise DeckError('IMMOBILE_DECAY_REACTION keyword ' + key)
            if k is None:
                raise DeckError('RATE_CONSTANT or HALF_LIFE must be set in IMMOBILE_DECAY_REACTION.')
            chem.immobile_decay_rxns.append((name, k))
        elif kw == 'MICROBIAL_REACTION':                    # reaction_microbial.F90:30-233
            r = MicrobialRxn()
            for t in rd.block():
                key = t[0].upper()
                if key == 'REACTION':
                    r.reaction = ' '.join(t[1:])
                elif key == 'RATE_CONSTANT':
                    r.rate_constant = fnum(t[1])
                elif key == 'ACTIVATION_ENERGY':
                    r.activation_energy = fnum(t[1])
                    if len(t) > 2 and t[2][0] not in '!#':
                        r.activation_energy = r.activation_energy * units_convert_to_internal(t[2], 'J/mol')
                elif key == 'MONOD':
                    name, K, Cth = '', 0.0, 0.0
                    for u in rd.block():
                        k2 = u[0].upper()
                        if k2 == 'SPECIES_NAME':
                            name = u[1]
                        elif k2 == 'HALF_SATURATION_CONSTANT':
                            K = fnum(u[1])
                        elif k2 == 'THRESHOLD_CONCENTRATION':
                            Cth = fnum(u[1])
                        else:
                            raise DeckError('MICROBIAL_REACTION,MONOD keyword ' + k2)
                    r.monod.append((name, K, Cth))
                elif key == 'INHIBITION':
                    name, itype, Cc, C2 = '', 0, None, 0.0
                    for u in rd.block():
                        k2 = u[0].upper()
                        if k2 == 'SPECIES_NAME':
                            name = u[1]
                        elif k2 == 'TYPE':
                            w = u[1].upper()
                            if w == 'MONOD':
                                itype = INHIBITION_MONOD
                            elif w == 'INVERSE_MONOD':
                                itype = INHIBITION_INVERSE_MONOD
                            elif w == 'THRESHOLD':
                                itype = INHIBITION_THRESHOLD
                                C2 = fnum(u[2])
                            else:
                                raise DeckError('MICROBIAL_REACTION,INHIBITION,TYPE ' + w)
                        elif k2 == 'INHIBITION_CONSTANT':
                            Cc = fnum(u[1])
                        else:
                            raise DeckError('MICROBIAL_REACTION,INHIBITION keyword ' + k2)
                    if len(name) < 2 or itype == 0 or Cc is None:
                        raise DeckError('A SPECIES_NAME, TYPE, and INHIBITION_CONSTANT must be defined for INHIBITION in MICROBIAL_REACTION')
                    r.inhibition.append((name, itype, Cc, C2))
                elif key == 'BIOMASS':
                    name, y = '', 0.0
                    for u in rd.block():
                        k2 = u[0].upper()
                        if k2 == 'SPECIES_NAME':
                            name = u[1]
                        elif k2 == 'YIELD':
                            y = fnum(u[1])
                        else:
                            raise DeckError('MICROBIAL_REACTION,BIOMASS keyword ' + k2)
                    r.biomass = (name, y)
                else:
                    raise DeckError('MICROBIAL_REACTION keyword ' + key)
            chem.microbial_rxns.append(r)
        elif kw in ('COLLOIDS',
                    'REACTION_SANDBOX', 'CLM_REACTION', 'SOLID_SOLUTIONS'):
            chem.unsupported.append(kw)
            rd.skip_block()
        elif kw in ('NO_CHECK_UPDATE', 'NO_RESTART_MINERAL_VOL_FRAC', 'NO_CHECKPOINT_ACT_COEFS',
                    'USE_FULL_GEOCHEMISTRY', 'NUMERICAL_JACOBIAN', 'TRUNCATE_CONCENTRATION',
                    'UPDATE_POROSITY', 'UPDATE_TORTUOSITY', 'UPDATE_PERMEABILITY',
                    'UPDATE_MINERAL_SURFACE_AREA', 'MINIMUM_POROSITY'):
            pass
        else:
            raise DeckError('CHEMISTRY keyword ' + kw)
    if len(chem.database) < 2:
        chem.act_coef_update_frequency = ACT_COEF_FREQUENCY_OFF
    return chem


_CTYPE = {
    'F': CONSTRAINT_FREE, 'FREE': CONSTRAINT_FREE,
    'T': CONSTRAINT_TOTAL, 'TOTAL': CONSTRAINT_TOTAL,
    'TOTAL_SORB': CONSTRAINT_TOTAL_SORB,
    'P': CONSTRAINT_PH, 'PH': CONSTRAINT_PH,
    'L': CONSTRAINT_LOG, 'LOG': CONSTRAINT_LOG,
    'M': CONSTRAINT_MINERAL, 'MINERAL': CONSTRAINT_MINERAL, 'MNRL': CONSTRAINT_MINERAL,
    'G': CONSTRAINT_GAS, 'GAS': CONSTRAINT_GAS,
    'SC': CONSTRAINT_SUPERCRIT_CO2,
    'Z': CONSTRAINT_CHARGE_BAL, 'CHG': CONSTRAINT_CHARGE_BAL,
}


def _read_constraint(rd: LineReader, name: str) -> Constraint:
    c = Constraint(name)
    for toks in rd.block():
        kw = toks[0].upper()
        if kw in ('CONC', 'CONCENTRATIONS'):
            for t in rd.block():
                c.names.append(t[0])
                c.conc.append(fnum(t[1]))
                ctype = CONSTRAINT_TOTAL
                aux = ''
                if len(t) > 2 and t[2][0] not in '!#':
                    ctype = _CTYPE[t[2].upper()]
                    if ctype in (CONSTRAINT_MINERAL, CONSTRAINT_GAS, CONSTRAINT_SUPERCRIT_CO2):
                        aux = t[3]
                c.ctype.append(ctype)
                c.aux.append(aux)
        elif kw == 'FREE_ION_GUESS':
            c.free_ion_guess = {}
            for t in rd.block():
                c.free_ion_guess[t[0]] = fnum(t[1])
        elif kw in ('MNRL', 'MINERALS'):
            for t in rd.block():
                vf = fnum(t[1])
                area = fnum(t[2])
                if len(t) > 3 and t[3][0] not in '!#':
                    area = area * units_convert_to_internal(t[3], 'm^2/m^3')
                c.minerals[t[0]] = (vf, area)
        elif kw == 'IMMOBILE':                              # transport_constraint.F90 (IMMOBILE block: name, concentration)
            for t in rd.block():
                c.immobile[t[0]] = fnum(t[1])
        elif kw in ('SURFACE_COMPLEXES', 'COLLOIDS'):
            rd.skip_block()
        else:
            raise DeckError('CONSTRAINT keyword ' + kw)
    return c


def read_deck(path: str) -> Deck:
    with open(path) as f:
        text = f.read()
    rd = LineReader(text)
    chem = None
    constraints: Dict[str, Constraint] = {}
    porosity = None
    ref_t = 25.0
    ref_p = 101325.0
    ref_den = None
    rock_den = None
    while True:
        toks = rd.next()
        if toks is None:
            break
        kw = toks[0].upper()
        if kw == 'CHEMISTRY':
            chem = _read_chemistry(rd)
        elif kw == 'CONSTRAINT':
            constraints[toks[1]] = _read_constraint(rd, toks[1])
        elif kw == 'MATERIAL_PROPERTY':
            for t in rd.block():
                if t[0].upper() == 'POROSITY' and porosity is None:
                    try:
                        porosity = fnum(t[1])
                    except ValueError:      # POROSITY DATASET name: per-cell values, not ours
                        porosity = None
                elif t[0].upper() == 'ROCK_DENSITY' and rock_den is None:
                    rock_den = fnum(t[1])
                elif t[0].upper() in ('PERMEABILITY', 'SATURATION_FUNCTION'):
                    # nested blocks
                    if len(t) == 1 or t[0].upper() == 'PERMEABILITY':
                        rd.skip_block()
        elif kw == 'REFERENCE_TEMPERATURE':
            ref_t = fnum(toks[1])
        elif kw == 'REFERENCE_PRESSURE':
            ref_p = fnum(toks[1])
        elif kw == 'REFERENCE_DENSITY':
            ref_den = fnum(toks[1])
    if chem is None:
        raise DeckError('no CHEMISTRY card in ' + path)
    return Deck(chem, constraints, porosity, ref_t, ref_p, path, rock_den, ref_den)
