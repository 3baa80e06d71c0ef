"""ctypes mirror of include/rxn_b200.h (struct layouts, enums) and the
ReactionTables -> RxnTablesDesc marshalling.

This is plumbing: it holds no arithmetic.  Field order and types must match the
header exactly; tests/test_abi.py checks sizeof() against the compiled library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)

# RxnStatus
RXN_OK, RXN_ERR_INVALID, RXN_ERR_UNSUPPORTED, RXN_ERR_CUDA, RXN_ERR_NO_DEVICE, RXN_ERR_CELL_FAILED = range(6)
RXN_DT_AS_WRITTEN, RXN_DT_CONSISTENT = 0, 1
RXN_EXIT_RESIDUAL, RXN_EXIT_REL_CHANGE = 1, 2
RXN_FLAG_CAPPED, RXN_FLAG_LU_ZERO_ROW, RXN_FLAG_ACT_DIVERGED, RXN_FLAG_NONFINITE, RXN_FLAG_INACTIVE = \
    1 << 8, 1 << 9, 1 << 10, 1 << 11, 1 << 12
# per-cell status of rxn_equilibrate_constraint_batch / constraint types (include/rxn_b200.h)
RXN_EQ_OK, RXN_EQ_NO_H_ION, RXN_EQ_BAD_CONSTRAINT, RXN_EQ_LU_ZERO_ROW, RXN_EQ_ZERO_CONCENTRATION, RXN_EQ_NOT_CONVERGED = 0, 2, 3, 4, 5, 6

FIELDS = [
    'PRI_MOLAL', 'TOTAL', 'SEC_MOLAL', 'PRI_ACT_COEF', 'SEC_ACT_COEF', 'LN_ACT_H2O',
    'TOTAL_SORB_EQ', 'FREE_SITE_CONC', 'EQSRFCPLX_CONC', 'KINMR_TOTAL_SORB',
    'EQIONX_REF_CATION_SORBED_CONC', 'EQIONX_CONC', 'MNRL_VOLFRAC', 'MNRL_AREA', 'MNRL_RATE',
    'DEN_KG', 'SAT', 'TEMP', 'PRES', 'VOLUME', 'POROSITY', 'SOIL_PARTICLE_DENSITY',
    'DTOTAL', 'DTOTAL_SORB_EQ',
    'KINSRFCPLX_CONC', 'KINSRFCPLX_CONC_KP1', 'KINSRFCPLX_FREE_SITE_CONC',
    'IMMOBILE',
]
F = {name: i for i, name in enumerate(FIELDS)}
RXN_F_COUNT = len(FIELDS)


class RxnSpecList(C.Structure):
    _fields_ = [
        ('id', c_i32p), ('stoich', c_f64p), ('h2oid', c_i32p), ('h2ostoich', c_f64p),
        ('logK', c_f64p), ('logKcoef', c_f64p),
        ('id_ld', C.c_int32), ('stoich_ld', C.c_int32), ('stoich_off', C.c_int32), ('n', C.c_int32),
    ]


class RxnTablesDesc(C.Structure):
    _fields_ = [
        ('struct_size', C.c_int32), ('naqcomp', C.c_int32), ('ncomp', C.c_int32),
        ('logK_mode', C.c_int32), ('num_logK_coef', C.c_int32), ('use_log_formulation', C.c_int32),
        ('act_coef_update_frequency', C.c_int32), ('act_coef_update_algorithm', C.c_int32),
        ('use_activity_h2o', C.c_int32), ('h2o_aq_id', C.c_int32), ('h_ion_id', C.c_int32),
        ('reserved0', C.c_int32),
        ('debyeA', C.c_double), ('debyeB', C.c_double), ('debyeBdot', C.c_double),
        ('max_dlnC', C.c_double), ('max_relative_change_tolerance', C.c_double),
        ('max_residual_tolerance', C.c_double),
        ('primary_spec_Z', c_f64p), ('primary_spec_a0', c_f64p),
        ('eqcplx', RxnSpecList), ('eqcplx_Z', c_f64p), ('eqcplx_a0', c_f64p),
        ('kinmnrl', RxnSpecList),
        ('kinmnrl_rate_constant', c_f64p), ('kinmnrl_activation_energy', c_f64p),
        ('kinmnrl_molar_vol', c_f64p), ('kinmnrl_affinity_threshold', c_f64p),
        ('kinmnrl_rate_limiter', c_f64p), ('kinmnrl_Temkin_const', c_f64p),
        ('kinmnrl_min_scale_factor', c_f64p), ('kinmnrl_affinity_power', c_f64p),
        ('kinmnrl_num_prefactors', c_i32p), ('kinmnrl_pref_rate', c_f64p),
        ('kinmnrl_pref_activation_energy', c_f64p), ('kinmnrl_prefactor_id', c_i32p),
        ('kinmnrl_pref_alpha', c_f64p), ('kinmnrl_pref_beta', c_f64p),
        ('kinmnrl_pref_atten_coef', c_f64p),
        ('max_num_prefactors', C.c_int32), ('max_num_prefactor_species', C.c_int32),
        ('mnrl', RxnSpecList), ('paseq', RxnSpecList),
        ('srfcplx', RxnSpecList), ('srfcplx_free_site_stoich', c_f64p), ('srfcplx_Z', c_f64p),
        ('nsrfcplxrxn', C.c_int32), ('srfcplxrxn_to_complex_ld', C.c_int32),
        ('srfcplxrxn_to_surf', c_i32p), ('srfcplxrxn_surf_type', c_i32p),
        ('srfcplxrxn_to_complex', c_i32p), ('srfcplxrxn_stoich_flag', c_i32p),
        ('srfcplxrxn_site_density', c_f64p),
        ('neqsrfcplxrxn', C.c_int32), ('nkinmrsrfcplxrxn', C.c_int32),
        ('eqsrfcplxrxn_to_srfcplxrxn', c_i32p), ('kinmrsrfcplxrxn_to_srfcplxrxn', c_i32p),
        ('kinmr_nrate', c_i32p), ('kinmr_rate', c_f64p), ('kinmr_frac', c_f64p),
        ('kinmr_ld', C.c_int32), ('nkinsrfcplxrxn', C.c_int32),
        ('neqionxrxn', C.c_int32), ('eqionx_ld', C.c_int32),
        ('eqionx_rxn_cationid', c_i32p), ('eqionx_rxn_k', c_f64p), ('eqionx_rxn_CEC', c_f64p),
        ('eqionx_rxn_Z_flag', c_i32p), ('eqionx_rxn_to_surf', c_i32p),
        ('neqkdrxn', C.c_int32), ('reserved1', C.c_int32),
        ('eqkdspecid', c_i32p), ('eqkdtype', c_i32p), ('eqkdmineral', c_i32p),
        ('eqkddistcoef', c_f64p), ('eqkdlangmuirb', c_f64p), ('eqkdfreundlichn', c_f64p),
        ('nactive_gas', C.c_int32), ('nimmobile', C.c_int32), ('ncoll', C.c_int32),
        ('ngeneral_rxn', C.c_int32), ('nradiodecay_rxn', C.c_int32), ('nmicrobial_rxn', C.c_int32),
        ('nimmobile_decay_rxn', C.c_int32), ('has_sandbox', C.c_int32), ('has_clm', C.c_int32),
        ('has_solid_solution', C.c_int32), ('co2_flow_mode', C.c_int32),
        ('numerical_derivatives', C.c_int32),
        ('general_ld', C.c_int32), ('radiodecay_ld', C.c_int32),
        ('generalspecid', c_i32p), ('generalstoich', c_f64p),
        ('generalforwardspecid', c_i32p), ('generalforwardstoich', c_f64p),
        ('generalbackwardspecid', c_i32p), ('generalbackwardstoich', c_f64p),
        ('general_kf', c_f64p), ('general_kr', c_f64p),
        ('radiodecayspecid', c_i32p), ('radiodecaystoich', c_f64p),
        ('radiodecayforwardspecid', c_i32p), ('radiodecay_kf', c_f64p),
        ('kinsrfcplxrxn_to_srfcplxrxn', c_i32p), ('kinsrfcplx_forward_rate', c_f64p),
        ('kinsrfcplx_backward_rate', c_f64p),
        ('kinsrfcplx_ld', C.c_int32), ('reserved2', C.c_int32),
        ('immobile_decayspecid', c_i32p), ('immobile_decay_rate_constant', c_f64p),
        ('microbial_ld', C.c_int32), ('microbial_monod_ld', C.c_int32), ('microbial_inhibition_ld', C.c_int32),
        ('nmicrobial_monod', C.c_int32), ('nmicrobial_inhibition', C.c_int32), ('reserved3', C.c_int32),
        ('microbial_specid', c_i32p), ('microbial_stoich', c_f64p),
        ('microbial_rate_constant', c_f64p), ('microbial_activation_energy', c_f64p),
        ('microbial_biomassid', c_i32p), ('microbial_biomass_yield', c_f64p),
        ('microbial_monodid', c_i32p), ('microbial_inhibitionid', c_i32p),
        ('microbial_monod_specid', c_i32p), ('microbial_monod_K', c_f64p), ('microbial_monod_Cth', c_f64p),
        ('microbial_inhibition_type', c_i32p), ('microbial_inhibition_specid', c_i32p),
        ('microbial_inhibition_C', c_f64p), ('microbial_inhibition_C2', c_f64p),
    ]


class HostView(C.Structure):
    """SoA host view used by the oracle: f[field] -> [rows][ld] doubles."""
    _fields_ = [('ncells', C.c_int64), ('ld', C.c_int64), ('f', c_f64p * RXN_F_COUNT)]


def _ip(a):
    return a.ctypes.data_as(c_i32p) if a is not None and a.size else c_i32p()


def _fp(a):
    return a.ctypes.data_as(c_f64p) if a is not None and a.size else c_f64p()


def make_desc(t) -> RxnTablesDesc:
    """Build the C descriptor from a ReactionTables.  The returned struct keeps the
    numpy arrays alive in `d._keep`."""
    d = RxnTablesDesc()
    keep: List[np.ndarray] = []

    def I(a):
        a = np.ascontiguousarray(a, dtype=np.int32); keep.append(a); return _ip(a)

    def D(a):
        a = np.ascontiguousarray(a, dtype=np.float64); keep.append(a); return _fp(a)

    def spec(ids, st, off, h2oid, h2ost, logK, coef):
        s = RxnSpecList()
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        st = np.ascontiguousarray(st, dtype=np.float64)
        n = ids.shape[0]
        s.id = I(ids); s.stoich = D(st); s.h2oid = I(h2oid); s.h2ostoich = D(h2ost)
        s.logK = D(logK)
        s.logKcoef = D(coef) if coef is not None else c_f64p()
        s.id_ld = ids.shape[1] if n else 1
        s.stoich_ld = st.shape[1] if n else 1
        s.stoich_off = off
        s.n = n
        return s

    d.struct_size = C.sizeof(RxnTablesDesc)
    d.naqcomp = t.naqcomp; d.ncomp = t.ncomp
    d.logK_mode = t.logK_mode; d.num_logK_coef = t.num_logK_coef
    d.use_log_formulation = t.use_log_formulation
    d.act_coef_update_frequency = t.act_coef_update_frequency
    d.act_coef_update_algorithm = t.act_coef_update_algorithm
    d.use_activity_h2o = t.use_activity_h2o
    d.h2o_aq_id = t.h2o_aq_id; d.h_ion_id = t.h_ion_id
    d.debyeA, d.debyeB, d.debyeBdot = t.debyeA, t.debyeB, t.debyeBdot
    d.max_dlnC = t.max_dlnC
    d.max_relative_change_tolerance = t.max_relative_change_tolerance
    d.max_residual_tolerance = t.max_residual_tolerance
    d.primary_spec_Z = D(t.primary_spec_Z); d.primary_spec_a0 = D(t.primary_spec_a0)
    d.eqcplx = spec(t.eqcplxspecid, t.eqcplxstoich, 0, t.eqcplxh2oid, t.eqcplxh2ostoich,
                    t.eqcplx_logK, t.eqcplx_logKcoef)
    d.eqcplx_Z = D(t.eqcplx_Z); d.eqcplx_a0 = D(t.eqcplx_a0)
    d.kinmnrl = spec(t.kinmnrlspecid, t.kinmnrlstoich, 1, t.kinmnrlh2oid, t.kinmnrlh2ostoich,
                     t.kinmnrl_logK, t.kinmnrl_logKcoef)
    d.kinmnrl_rate_constant = D(t.kinmnrl_rate_constant)
    d.kinmnrl_activation_energy = D(t.kinmnrl_activation_energy)
    d.kinmnrl_molar_vol = D(t.kinmnrl_molar_vol)
    d.kinmnrl_affinity_threshold = D(t.kinmnrl_affinity_threshold)
    d.kinmnrl_rate_limiter = D(t.kinmnrl_rate_limiter)
    d.kinmnrl_Temkin_const = D(t.kinmnrl_Temkin_const) if t.has_Temkin_const else c_f64p()
    d.kinmnrl_min_scale_factor = D(t.kinmnrl_min_scale_factor) if t.has_min_scale_factor else c_f64p()
    d.kinmnrl_affinity_power = D(t.kinmnrl_affinity_power) if t.has_affinity_power else c_f64p()
    d.kinmnrl_num_prefactors = I(t.kinmnrl_num_prefactors)
    d.kinmnrl_pref_rate = D(t.kinmnrl_pref_rate)
    d.kinmnrl_pref_activation_energy = D(t.kinmnrl_pref_activation_energy)
    d.kinmnrl_prefactor_id = I(t.kinmnrl_prefactor_id)
    d.kinmnrl_pref_alpha = D(t.kinmnrl_pref_alpha)
    d.kinmnrl_pref_beta = D(t.kinmnrl_pref_beta)
    d.kinmnrl_pref_atten_coef = D(t.kinmnrl_pref_atten_coef)
    d.max_num_prefactors = t.max_num_prefactors
    d.max_num_prefactor_species = t.max_num_prefactor_species
    d.mnrl = spec(t.mnrlspecid, t.mnrlstoich, 1, t.mnrlh2oid, t.mnrlh2ostoich, t.mnrl_logK, None)
    d.paseq = spec(t.paseqspecid, t.paseqstoich, 0, t.paseqh2oid, t.paseqh2ostoich, t.paseqlogK, None)
    d.srfcplx = spec(t.srfcplxspecid, t.srfcplxstoich, 1, t.srfcplxh2oid, t.srfcplxh2ostoich,
                     t.srfcplx_logK, t.srfcplx_logKcoef)
    d.srfcplx_free_site_stoich = D(t.srfcplx_free_site_stoich); d.srfcplx_Z = D(t.srfcplx_Z)
    d.nsrfcplxrxn = t.nsrfcplxrxn
    d.srfcplxrxn_to_complex_ld = t.srfcplxrxn_to_complex.shape[1] if t.nsrfcplxrxn else 1
    d.srfcplxrxn_to_surf = I(t.srfcplxrxn_to_surf)
    d.srfcplxrxn_surf_type = I(t.srfcplxrxn_surf_type)
    d.srfcplxrxn_to_complex = I(t.srfcplxrxn_to_complex)
    d.srfcplxrxn_stoich_flag = I(t.srfcplxrxn_stoich_flag)
    d.srfcplxrxn_site_density = D(t.srfcplxrxn_site_density)
    d.neqsrfcplxrxn = t.neqsrfcplxrxn; d.nkinmrsrfcplxrxn = t.nkinmrsrfcplxrxn
    d.eqsrfcplxrxn_to_srfcplxrxn = I(t.eqsrfcplxrxn_to_srfcplxrxn)
    d.kinmrsrfcplxrxn_to_srfcplxrxn = I(t.kinmrsrfcplxrxn_to_srfcplxrxn)
    d.kinmr_nrate = I(t.kinmr_nrate); d.kinmr_rate = D(t.kinmr_rate); d.kinmr_frac = D(t.kinmr_frac)
    d.kinmr_ld = t.kinmr_max_nrate; d.nkinsrfcplxrxn = t.nkinsrfcplxrxn
    d.neqionxrxn = t.neqionxrxn
    d.eqionx_ld = t.eqionx_rxn_cationid.shape[1] - 1 if t.neqionxrxn else 0
    d.eqionx_rxn_cationid = I(t.eqionx_rxn_cationid); d.eqionx_rxn_k = D(t.eqionx_rxn_k)
    d.eqionx_rxn_CEC = D(t.eqionx_rxn_CEC); d.eqionx_rxn_Z_flag = I(t.eqionx_rxn_Z_flag)
    d.eqionx_rxn_to_surf = I(t.eqionx_rxn_to_surf)
    d.neqkdrxn = t.neqkdrxn
    d.eqkdspecid = I(t.eqkdspecid); d.eqkdtype = I(t.eqkdtype); d.eqkdmineral = I(t.eqkdmineral)
    d.eqkddistcoef = D(t.eqkddistcoef); d.eqkdlangmuirb = D(t.eqkdlangmuirb)
    d.eqkdfreundlichn = D(t.eqkdfreundlichn)
    uns = getattr(t, 'unsupported', [])
    d.nactive_gas = int(any('ACTIVE_GAS' in u for u in uns))
    # immobile species, their decay and microbial reactions (fixtures older than this feature have none)
    d.nimmobile = int(getattr(t, 'nimmobile', 0))
    d.nimmobile_decay_rxn = int(getattr(t, 'nimmobile_decay_rxn', 0))
    d.nmicrobial_rxn = int(getattr(t, 'nmicrobial_rxn', 0))
    if d.nimmobile_decay_rxn:
        d.immobile_decayspecid = I(t.immobile_decayspecid); d.immobile_decay_rate_constant = D(t.immobile_decay_rate_constant)
    if d.nmicrobial_rxn:
        d.microbial_ld = t.microbial_stoich.shape[1]
        d.microbial_monod_ld = t.microbial_monodid.shape[1] - 1
        d.microbial_inhibition_ld = t.microbial_inhibitionid.shape[1] - 1
        d.nmicrobial_monod = t.microbial_monod_specid.size; d.nmicrobial_inhibition = t.microbial_inhibition_specid.size
        d.microbial_specid = I(t.microbial_specid); d.microbial_stoich = D(t.microbial_stoich)
        d.microbial_rate_constant = D(t.microbial_rate_constant)
        d.microbial_activation_energy = D(t.microbial_activation_energy) if t.has_microbial_activation_energy else c_f64p()
        d.microbial_biomassid = I(t.microbial_biomassid); d.microbial_biomass_yield = D(t.microbial_biomass_yield)
        d.microbial_monodid = I(t.microbial_monodid); d.microbial_inhibitionid = I(t.microbial_inhibitionid)
        d.microbial_monod_specid = I(t.microbial_monod_specid); d.microbial_monod_K = D(t.microbial_monod_K)
        d.microbial_monod_Cth = D(t.microbial_monod_Cth)
        d.microbial_inhibition_type = I(t.microbial_inhibition_type); d.microbial_inhibition_specid = I(t.microbial_inhibition_specid)
        d.microbial_inhibition_C = D(t.microbial_inhibition_C); d.microbial_inhibition_C2 = D(t.microbial_inhibition_C2)
    d.ncoll = int(any('COLLOID' in u for u in uns))
    # general reactions / radioactive decay / kinetic surface complexation (fixtures older than round 2 have none)
    d.ngeneral_rxn = int(getattr(t, 'ngeneral_rxn', 0))
    d.nradiodecay_rxn = int(getattr(t, 'nradiodecay_rxn', 0))
    if d.ngeneral_rxn:
        d.general_ld = t.generalstoich.shape[1]
        d.generalspecid = I(t.generalspecid); d.generalstoich = D(t.generalstoich)
        d.generalforwardspecid = I(t.generalforwardspecid); d.generalforwardstoich = D(t.generalforwardstoich)
        d.generalbackwardspecid = I(t.generalbackwardspecid); d.generalbackwardstoich = D(t.generalbackwardstoich)
        d.general_kf = D(t.general_kf); d.general_kr = D(t.general_kr)
    if d.nradiodecay_rxn:
        d.radiodecay_ld = t.radiodecaystoich.shape[1]
        d.radiodecayspecid = I(t.radiodecayspecid); d.radiodecaystoich = D(t.radiodecaystoich)
        d.radiodecayforwardspecid = I(t.radiodecayforwardspecid); d.radiodecay_kf = D(t.radiodecay_kf)
    if t.nkinsrfcplxrxn:
        d.kinsrfcplx_ld = t.kinsrfcplx_forward_rate.shape[1]
        d.kinsrfcplxrxn_to_srfcplxrxn = I(t.kinsrfcplxrxn_to_srfcplxrxn)
        d.kinsrfcplx_forward_rate = D(t.kinsrfcplx_forward_rate)
        d.kinsrfcplx_backward_rate = D(t.kinsrfcplx_backward_rate)
    d.has_sandbox = int(any('SANDBOX' in u for u in uns))
    d.has_clm = int(any('CLM' in u for u in uns))
    d.has_solid_solution = int(any('SOLID_SOLUTION' in u for u in uns))
    d._keep = keep
    return d


def field_rows(t) -> Dict[str, int]:
    """Rows of every per-cell field for tables t (mirror of rxn_field_rows)."""
    naq = t.naqcomp
    ionx_ld = (t.eqionx_rxn_cationid.shape[1] - 1) if t.neqionxrxn else 0
    return {
        'PRI_MOLAL': naq, 'TOTAL': naq, 'SEC_MOLAL': t.neqcplx, 'PRI_ACT_COEF': naq,
        'SEC_ACT_COEF': t.neqcplx, 'LN_ACT_H2O': 1, 'TOTAL_SORB_EQ': naq,
        'FREE_SITE_CONC': t.nsrfcplxrxn, 'EQSRFCPLX_CONC': t.nsrfcplx,
        'KINMR_TOTAL_SORB': t.nkinmrsrfcplxrxn * (t.kinmr_max_nrate + 1) * naq,
        'EQIONX_REF_CATION_SORBED_CONC': t.neqionxrxn, 'EQIONX_CONC': t.neqionxrxn * ionx_ld,
        'MNRL_VOLFRAC': t.nkinmnrl, 'MNRL_AREA': t.nkinmnrl, 'MNRL_RATE': t.nkinmnrl,
        'DEN_KG': 1, 'SAT': 1, 'TEMP': 1, 'PRES': 1, 'VOLUME': 1, 'POROSITY': 1,
        'SOIL_PARTICLE_DENSITY': 1, 'DTOTAL': naq * naq, 'DTOTAL_SORB_EQ': naq * naq,
        'KINSRFCPLX_CONC': getattr(t, 'nkinsrfcplx', 0), 'KINSRFCPLX_CONC_KP1': getattr(t, 'nkinsrfcplx', 0),
        'KINSRFCPLX_FREE_SITE_CONC': t.nkinsrfcplxrxn,
        'IMMOBILE': getattr(t, 'nimmobile', 0),
    }


class HostState:
    """Host mirror of the per-cell state, SoA float64 [rows, ncells] per field, with the
    reference's initial values (reactive_transport_aux.F90:213-400)."""

    def __init__(self, t, ncells: int):
        self.t = t
        self.ncells = ncells
        self.rows = field_rows(t)
        self.a: Dict[str, np.ndarray] = {}
        for name in FIELDS:
            self.a[name] = np.zeros((self.rows[name], ncells), dtype=np.float64)
        self.a['PRI_ACT_COEF'][:] = 1.0
        self.a['SEC_ACT_COEF'][:] = 1.0
        self.a['FREE_SITE_CONC'][:] = 1.0e-9
        self.a['EQIONX_REF_CATION_SORBED_CONC'][:] = 1.0e-9
        self.active = np.ones(ncells, dtype=np.uint8)

    def __getitem__(self, name):
        return self.a[name]

    def copy(self) -> 'HostState':
        o = HostState(self.t, self.ncells)
        for k in self.a:
            o.a[k] = self.a[k].copy()
        o.active = self.active.copy()
        return o

    def view(self) -> HostView:
        v = HostView()
        v.ncells = self.ncells
        v.ld = self.ncells
        for i, name in enumerate(FIELDS):
            arr = self.a[name]
            v.f[i] = arr.ctypes.data_as(c_f64p) if arr.size else c_f64p()
        return v

# coupler (boundary / source-sink) connection sets, include/rxn_b200.h
RXN_COUPLER_BOUNDARY, RXN_COUPLER_SRC_SINK = 0, 1
RXN_SS_MASS_RATE, RXN_SS_EQUILIBRIUM = 7, 12
