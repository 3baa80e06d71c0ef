// rxn_tab.h — device-side view of the chemistry tables and of the per-cell SoA state.
//
// The host packs RxnTablesDesc (include/rxn_b200.h; reference reaction_type,
// src/pflotran/reaction_aux.F90:142-335) into ONE blob: [doubles][int32s].  Every
// kernel stages that blob into shared memory once per CTA and reads it through the
// offsets below (all accesses are warp-uniform -> shared-memory broadcasts).
// Species lists are CSR (ptr/id/stoich) with 0-based primary ids.
#pragma once
#include <stdint.h>
#include "../../include/rxn_b200.h"

namespace rxn {

struct DSpec {      // one reaction list (aqueous complexes / kinetic minerals / surface complexes)
  int n;            // number of reactions
  int o_ptr;        // int  [n+1]  CSR row pointer
  int o_id;         // int  [nnz]  0-based primary species id
  int o_st;         // dbl  [nnz]  stoichiometry
  int o_h2ost;      // dbl  [n]    H2O stoichiometry (0 when H2O takes no part)
  int o_logK;       // dbl  [n]
  int o_coef;       // dbl  [n][ncoef]  or -1
};

struct DevTab {
  int naq, ncplx, nkin, nsrf, nrxn, neq, nmr, nionx, nkd, neqsorb;
  int logK_mode, ncoef, use_log, act_freq, act_alg, use_act_h2o, h2o_aq_id;
  int has_Temkin, has_scale, has_power, maxpref, maxprefspec, mr_ld, ionx_ld, maxit;
  int ndbl, nint;   // blob sizes
  double debyeA, debyeB, debyeBdot, max_dlnC, rel_tol, res_tol;
  DSpec cplx, kin, srf;
  DSpec mnrl, gas;  // all minerals / passive gases: constraint equilibration only (logK at the reference temperature)
  int h_ion_id;     // species_idx%h_ion_id: > 0 primary, < 0 complex, 0 none
  int o_Z, o_a0, o_cplxZ, o_cplxa0;                                        // dbl
  int o_k_rate, o_k_Ea, o_k_molar_vol, o_k_aff, o_k_lim, o_k_Temkin, o_k_scale, o_k_power;  // dbl [nkin]
  int o_k_npref;                                                            // int [nkin]
  int o_pref_rate, o_pref_Ea, o_pref_alpha, o_pref_beta, o_pref_atten;      // dbl
  int o_pref_id;                                                            // int (nkin, maxpref, maxprefspec+1)
  int o_srf_site_st;                                                        // dbl [nsrf]
  int o_rxn_to_surf, o_rxn_surf_type, o_rxn_flag, o_rxn_cptr, o_rxn_cid;    // int
  int o_rxn_density;                                                        // dbl [nrxn]
  int o_eq_rxn, o_mr_rxn, o_mr_nrate;                                       // int
  int o_mr_rate, o_mr_frac;                                                 // dbl (nmr, mr_ld)
  int o_ionx_ptr, o_ionx_cat, o_ionx_Zflag, o_ionx_to_surf;                 // int
  int o_ionx_k, o_ionx_CEC;                                                 // dbl
  int o_kd_spec, o_kd_type, o_kd_mnrl;                                      // int
  int o_kd_coef, o_kd_b, o_kd_n;                                            // dbl
  // general reactions (RGeneral), radioactive decay (RRadioactiveDecay): CSR lists, 0-based ids
  int ngen, ndecay;
  int o_gen_ptr, o_gen_id, o_genf_ptr, o_genf_id, o_genb_ptr, o_genb_id;   // int
  int o_gen_st, o_genf_st, o_genb_st, o_gen_kf, o_gen_kr;                   // dbl
  int o_dec_ptr, o_dec_id, o_dec_fwd;                                       // int ([ndecay] reactant id, 0-based)
  int o_dec_st, o_dec_kf;                                                   // dbl
  // kinetic surface complexation (RKineticSurfCplx): at most one reaction, = surface complexation reaction kin_rxn (0-based)
  int nkinrxn, nkinsrf, kin_rxn;
  int o_kin_kf, o_kin_kb;                                                   // dbl [nkinsrf]
  // immobile species (dofs naq .. ncomp-1), their decay (RImmobileDecay) and microbial reactions (RMicrobial); 0-based ids,
  // microbial species ids run over the ncomp dofs
  int nim, ncomp, nimdecay, nmic, mic_has_Ea;
  int o_imdec_id;                                                           // int [nimdecay] immobile id
  int o_imdec_k;                                                            // dbl [nimdecay]
  int o_mic_ptr, o_mic_id, o_mic_bio, o_mic_mptr, o_mic_mid, o_mic_iptr, o_mic_iid;   // int (o_mic_bio: immobile id or -1)
  int o_mic_st, o_mic_k, o_mic_Ea, o_mic_yield;                             // dbl
  int o_mon_spec, o_inh_spec, o_inh_type;                                   // int
  int o_mon_K, o_mon_Cth, o_inh_C, o_inh_C2;                                // dbl
};

// SoA FP64 state in HBM: f[field][row*ld + cell]; NULL when the field has no rows or is
// not materialised (DTOTAL/DTOTAL_SORB_EQ until a global-implicit entry point needs them).
struct DevState {
  double *f[RXN_F_COUNT];
  long long ld;
  long long ncells;
  const uint8_t *active;   // NULL = all active
  unsigned int *fail;      // OR of the RXN_FLAG_* raised by the cells of a global-implicit launch (NULL: not collected)
  const int32_t *order = nullptr;   // resident-lane RReact: the work counter's n-th item is batch item order[n] (NULL: identity)
};

// hard limits of the per-thread scratch (tables beyond them are rejected at create time)
enum { RXN_MAX_SRFCPLX_PER_RXN = 32, RXN_MAX_PREF = 10, RXN_MAX_PREF_SPEC = 5, RXN_MAX_NAQ = 24,
       RXN_MAX_IMMOBILE = 4, RXN_MAX_MONOD = 10 /* monod(10), inhibition(10): reaction_microbial.F90:270-271 */ };

}  // namespace rxn
