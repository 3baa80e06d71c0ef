// rxn_pack.h — host-side validation and packing of RxnTablesDesc into the device blob.
// Pure C++ (no CUDA): reference reaction_type / mineral_type / surface_complexation_type
// compressed tables (reaction_aux.F90:142-335, reaction_mineral_aux.F90:77-128,
// reaction_surf_complex_aux.F90:68-128) -> CSR lists with 0-based ids + DevTab offsets.
#pragma once
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rxn_tab.h"

namespace rxn {

struct Packer {
  std::vector<double> d;
  std::vector<int32_t> i;
  int D(const double *p, size_t n) {
    int o = (int)d.size();
    for (size_t k = 0; k < n; ++k) d.push_back(p ? p[k] : 0.0);
    return o;
  }
  int I(const int32_t *p, size_t n) {
    int o = (int)i.size();
    for (size_t k = 0; k < n; ++k) i.push_back(p ? p[k] : 0);
    return o;
  }
};

struct PackResult {
  DevTab h;
  Packer P;
  int rows[RXN_F_COUNT];
  std::string err;
  int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

// RxnSpecList (Fortran compressed form, 1-based ids) -> CSR with 0-based ids
inline int pack_spec(PackResult &R, const RxnSpecList &s, int ncoef, int naq, DSpec &o, const char *what) {
  Packer &P = R.P;
  o.n = s.n;
  std::vector<int32_t> ptr(1, 0), id;
  std::vector<double> st, h2ost((size_t)std::max(s.n, 0), 0.0), logK((size_t)std::max(s.n, 0), 0.0);
  for (int r = 0; r < s.n; ++r) {
    const int ns = s.id[(size_t)r * s.id_ld];
    if (ns < 0 || ns >= s.id_ld + 1) return R.fail(RXN_ERR_INVALID, "%s %d: bad species count %d", what, r, ns);
    for (int k = 1; k <= ns; ++k) {
      const int sp = s.id[(size_t)r * s.id_ld + k];
      if (sp < 1 || sp > naq) return R.fail(RXN_ERR_INVALID, "%s %d: species id %d out of range", what, r, sp);
      id.push_back(sp - 1);
      st.push_back(s.stoich[(size_t)r * s.stoich_ld + k - s.stoich_off]);
    }
    ptr.push_back((int32_t)id.size());
    if (s.h2oid && s.h2oid[r] > 0 && s.h2ostoich) h2ost[r] = s.h2ostoich[r];
    if (s.logK) logK[r] = s.logK[r];
  }
  o.o_ptr = P.I(ptr.data(), ptr.size());
  o.o_id = P.I(id.data(), id.size());
  o.o_st = P.D(st.data(), st.size());
  o.o_h2ost = P.D(h2ost.data(), h2ost.size());
  o.o_logK = P.D(logK.data(), logK.size());
  o.o_coef = (s.logKcoef && ncoef > 0 && s.n > 0) ? P.D(s.logKcoef, (size_t)s.n * ncoef) : -1;
  return RXN_OK;
}

inline int pack_tables(const RxnTablesDesc *d, PackResult &R) {
  if (!d) return R.fail(RXN_ERR_INVALID, "null descriptor");
  if (d->struct_size != (int)sizeof(RxnTablesDesc))
    return R.fail(RXN_ERR_INVALID, "RxnTablesDesc.struct_size %d != %d", d->struct_size, (int)sizeof(RxnTablesDesc));
  // reaction types outside the path
  if (d->nactive_gas || d->ncoll || d->has_sandbox || d->has_clm || d->has_solid_solution || d->co2_flow_mode ||
      d->numerical_derivatives)
    return R.fail(RXN_ERR_UNSUPPORTED,
                  "tables enable a reaction type outside the B200 path (active gas %d, colloids %d, "
                  "sandbox %d, CLM %d, solid solution %d, CO2 flow mode %d, numerical Jacobian %d)",
                  d->nactive_gas, d->ncoll, d->has_sandbox, d->has_clm, d->has_solid_solution, d->co2_flow_mode,
                  d->numerical_derivatives);
  if (d->ngeneral_rxn < 0 || d->nradiodecay_rxn < 0 || d->nkinsrfcplxrxn < 0 || d->nimmobile < 0 || d->nmicrobial_rxn < 0 ||
      d->nimmobile_decay_rxn < 0)
    return R.fail(RXN_ERR_INVALID, "negative reaction count");
  if (d->nkinsrfcplxrxn > 1)
    return R.fail(RXN_ERR_UNSUPPORTED, "more than one KINETIC surface complexation reaction: the reference keeps the kinetic "
                                       "concentrations of one reaction only (kinsrfcplx_conc(:,1), reactive_transport_aux.F90:284-290)");
  if (d->naqcomp < 1 || d->ncomp != d->naqcomp + d->nimmobile)
    return R.fail(RXN_ERR_UNSUPPORTED, "ncomp (%d) must equal naqcomp (%d) >= 1 plus nimmobile (%d): no colloid dofs", d->ncomp, d->naqcomp, d->nimmobile);
  if (d->ncomp > RXN_MAX_NAQ) return R.fail(RXN_ERR_UNSUPPORTED, "ncomp %d > %d", d->ncomp, (int)RXN_MAX_NAQ);
  if (d->nimmobile > RXN_MAX_IMMOBILE) return R.fail(RXN_ERR_UNSUPPORTED, "nimmobile %d > %d", d->nimmobile, (int)RXN_MAX_IMMOBILE);
  if (d->nkinmrsrfcplxrxn > 2) return R.fail(RXN_ERR_UNSUPPORTED, "more than 2 multirate surface complexation reactions");
  if (d->max_num_prefactors > RXN_MAX_PREF || d->max_num_prefactor_species > RXN_MAX_PREF_SPEC)
    return R.fail(RXN_ERR_UNSUPPORTED, "mineral prefactor table too large");
  if (d->logK_mode != RXN_LOGK_FIXED && !((d->logK_mode == RXN_LOGK_FIT5 && d->num_logK_coef == 5) ||
                                          (d->logK_mode == RXN_LOGK_HPT && d->num_logK_coef == 17)))
    return R.fail(RXN_ERR_INVALID, "logK_mode %d with %d coefficients", d->logK_mode, d->num_logK_coef);
  if (d->act_coef_update_algorithm != RXN_ACT_COEF_ALGORITHM_LAG && d->act_coef_update_algorithm != RXN_ACT_COEF_ALGORITHM_NEWTON)
    return R.fail(RXN_ERR_INVALID, "act_coef_update_algorithm %d", d->act_coef_update_algorithm);

  DevTab &h = R.h;
  Packer &P = R.P;
  memset(&h, 0, sizeof h);
  const int naq = d->naqcomp;
  h.naq = naq; h.ncplx = d->eqcplx.n; h.nkin = d->kinmnrl.n; h.nsrf = d->srfcplx.n; h.nrxn = d->nsrfcplxrxn;
  h.neq = d->neqsrfcplxrxn; h.nmr = d->nkinmrsrfcplxrxn; h.nionx = d->neqionxrxn; h.nkd = d->neqkdrxn;
  h.neqsorb = h.neq + h.nionx + h.nkd;
  h.logK_mode = d->logK_mode; h.ncoef = d->num_logK_coef; h.use_log = d->use_log_formulation;
  h.act_freq = d->act_coef_update_frequency; h.act_alg = d->act_coef_update_algorithm;
  h.use_act_h2o = d->use_activity_h2o; h.h2o_aq_id = d->h2o_aq_id;
  h.has_Temkin = d->kinmnrl_Temkin_const != nullptr; h.has_scale = d->kinmnrl_min_scale_factor != nullptr;
  h.has_power = d->kinmnrl_affinity_power != nullptr;
  h.maxpref = d->max_num_prefactors; h.maxprefspec = d->max_num_prefactor_species;
  h.mr_ld = d->kinmr_ld; h.ionx_ld = d->eqionx_ld;
  h.maxit = 10000;
  if (const char *e = getenv("RXN_MAX_NEWTON_ITERATIONS")) h.maxit = std::max(1, atoi(e));
  h.debyeA = d->debyeA; h.debyeB = d->debyeB; h.debyeBdot = d->debyeBdot;
  h.max_dlnC = d->max_dlnC; h.rel_tol = d->max_relative_change_tolerance; h.res_tol = d->max_residual_tolerance;
  int rc;
#define RXN_TRY(x) do { rc = (x); if (rc != RXN_OK) return rc; } while (0)
  h.o_Z = P.D(d->primary_spec_Z, naq); h.o_a0 = P.D(d->primary_spec_a0, naq);
  RXN_TRY(pack_spec(R, d->eqcplx, h.ncoef, naq, h.cplx, "aqueous complex"));
  h.o_cplxZ = P.D(d->eqcplx_Z, h.ncplx); h.o_cplxa0 = P.D(d->eqcplx_a0, h.ncplx);
  RXN_TRY(pack_spec(R, d->kinmnrl, h.ncoef, naq, h.kin, "kinetic mineral"));
  const int nk = h.nkin;
  h.o_k_rate = P.D(d->kinmnrl_rate_constant, nk); h.o_k_Ea = P.D(d->kinmnrl_activation_energy, nk);
  h.o_k_molar_vol = P.D(d->kinmnrl_molar_vol, nk); h.o_k_aff = P.D(d->kinmnrl_affinity_threshold, nk);
  h.o_k_lim = P.D(d->kinmnrl_rate_limiter, nk); h.o_k_Temkin = P.D(d->kinmnrl_Temkin_const, nk);
  h.o_k_scale = P.D(d->kinmnrl_min_scale_factor, nk); h.o_k_power = P.D(d->kinmnrl_affinity_power, nk);
  h.o_k_npref = P.I(d->kinmnrl_num_prefactors, nk);
  {
    const size_t np = (size_t)nk * std::max(h.maxpref, 1), nps = std::max(h.maxprefspec, 1);
    h.o_pref_rate = P.D(d->kinmnrl_pref_rate, np); h.o_pref_Ea = P.D(d->kinmnrl_pref_activation_energy, np);
    h.o_pref_id = P.I(d->kinmnrl_prefactor_id, np * (h.maxprefspec + 1));
    h.o_pref_alpha = P.D(d->kinmnrl_pref_alpha, np * nps); h.o_pref_beta = P.D(d->kinmnrl_pref_beta, np * nps);
    h.o_pref_atten = P.D(d->kinmnrl_pref_atten_coef, np * nps);
    if (d->kinmnrl_prefactor_id && d->kinmnrl_num_prefactors)
      for (int im = 0; im < nk; ++im)
        for (int ip = 0; ip < d->kinmnrl_num_prefactors[im]; ++ip) {
          const int32_t *row = d->kinmnrl_prefactor_id + ((size_t)im * std::max(h.maxpref, 1) + ip) * (h.maxprefspec + 1);
          for (int k = 1; k <= row[0]; ++k)
            if (row[k] < 1 || row[k] > naq)
              return R.fail(RXN_ERR_UNSUPPORTED,
                            "mineral %d prefactor %d on a secondary species (id %d): the reference branch "
                            "(reaction_mineral.F90:977-979) clobbers its loop variables; not supported",
                            im + 1, ip + 1, row[k]);
        }
  }
  RXN_TRY(pack_spec(R, d->mnrl, 0, naq, h.mnrl, "mineral"));
  RXN_TRY(pack_spec(R, d->paseq, 0, naq, h.gas, "passive gas"));
  h.h_ion_id = d->h_ion_id;
  RXN_TRY(pack_spec(R, d->srfcplx, h.ncoef, naq, h.srf, "surface complex"));
  h.o_srf_site_st = P.D(d->srfcplx_free_site_stoich, h.nsrf);
  h.o_rxn_to_surf = P.I(d->srfcplxrxn_to_surf, h.nrxn); h.o_rxn_surf_type = P.I(d->srfcplxrxn_surf_type, h.nrxn);
  h.o_rxn_flag = P.I(d->srfcplxrxn_stoich_flag, h.nrxn); h.o_rxn_density = P.D(d->srfcplxrxn_site_density, h.nrxn);
  {
    std::vector<int32_t> cptr(1, 0), cid;
    for (int r = 0; r < h.nrxn; ++r) {
      const int ld = d->srfcplxrxn_to_complex_ld, n = d->srfcplxrxn_to_complex[(size_t)r * ld];
      if (n < 0 || n > ld - 1) return R.fail(RXN_ERR_INVALID, "surface complexation reaction %d: complex count %d out of range", r + 1, n);
      if (n > RXN_MAX_SRFCPLX_PER_RXN)
        return R.fail(RXN_ERR_UNSUPPORTED, "surface complexation reaction %d has %d complexes (> %d)", r + 1, n, (int)RXN_MAX_SRFCPLX_PER_RXN);
      for (int k = 1; k <= n; ++k) {
        const int ic = d->srfcplxrxn_to_complex[(size_t)r * ld + k];
        if (ic < 1 || ic > h.nsrf) return R.fail(RXN_ERR_INVALID, "surface complexation reaction %d: complex id %d out of range", r + 1, ic);
        cid.push_back(ic - 1);
      }
      cptr.push_back((int32_t)cid.size());
      const int st = d->srfcplxrxn_surf_type[r];
      if (st == RXN_COLLOID_SURFACE) return R.fail(RXN_ERR_UNSUPPORTED, "colloid surface");
      if (st == RXN_MINERAL_SURFACE && (d->srfcplxrxn_to_surf[r] < 1 || d->srfcplxrxn_to_surf[r] > nk))
        return R.fail(RXN_ERR_INVALID, "surface complexation reaction %d: mineral id out of range", r + 1);
    }
    h.o_rxn_cptr = P.I(cptr.data(), cptr.size()); h.o_rxn_cid = P.I(cid.data(), cid.size());
    std::vector<int32_t> eq, mr, mrn;
    for (int i = 0; i < h.neq; ++i) {
      const int ir = d->eqsrfcplxrxn_to_srfcplxrxn[i];
      if (ir < 1 || ir > h.nrxn) return R.fail(RXN_ERR_INVALID, "eqsrfcplxrxn_to_srfcplxrxn[%d] = %d out of range", i, ir);
      eq.push_back(ir - 1);
    }
    for (int i = 0; i < h.nmr; ++i) {
      const int ir = d->kinmrsrfcplxrxn_to_srfcplxrxn[i], nr = d->kinmr_nrate[i + 1];
      if (ir < 1 || ir > h.nrxn) return R.fail(RXN_ERR_INVALID, "kinmrsrfcplxrxn_to_srfcplxrxn[%d] = %d out of range", i, ir);
      if (nr < 0 || nr > h.mr_ld) return R.fail(RXN_ERR_INVALID, "kinmr_nrate[%d] = %d exceeds the leading dimension %d", i + 1, nr, h.mr_ld);
      mr.push_back(ir - 1);
      mrn.push_back(nr);
    }
    h.o_eq_rxn = P.I(eq.data(), eq.size()); h.o_mr_rxn = P.I(mr.data(), mr.size()); h.o_mr_nrate = P.I(mrn.data(), mrn.size());
    h.o_mr_rate = P.D(d->kinmr_rate, (size_t)h.nmr * h.mr_ld); h.o_mr_frac = P.D(d->kinmr_frac, (size_t)h.nmr * h.mr_ld);
  }
  {
    std::vector<int32_t> iptr(1, 0), cat;
    std::vector<double> kk;
    for (int r = 0; r < h.nionx; ++r) {
      const int ld = h.ionx_ld + 1, n = d->eqionx_rxn_cationid[(size_t)r * ld];
      if (n < 0 || n > h.ionx_ld) return R.fail(RXN_ERR_INVALID, "ion exchange reaction %d: cation count %d exceeds the leading dimension %d", r + 1, n, h.ionx_ld);
      for (int k = 1; k <= n; ++k) {
        const int ic = d->eqionx_rxn_cationid[(size_t)r * ld + k];
        if (ic < 1 || ic > naq) return R.fail(RXN_ERR_INVALID, "ion exchange reaction %d: cation id %d out of range", r + 1, ic);
        cat.push_back(ic - 1);
        kk.push_back(d->eqionx_rxn_k[(size_t)r * h.ionx_ld + k - 1]);
      }
      iptr.push_back((int32_t)cat.size());
      if (d->eqionx_rxn_to_surf && d->eqionx_rxn_to_surf[r] > nk)
        return R.fail(RXN_ERR_INVALID, "ion exchange reaction %d: mineral id out of range", r + 1);
    }
    h.o_ionx_ptr = P.I(iptr.data(), iptr.size()); h.o_ionx_cat = P.I(cat.data(), cat.size());
    h.o_ionx_k = P.D(kk.data(), kk.size()); h.o_ionx_CEC = P.D(d->eqionx_rxn_CEC, h.nionx);
    h.o_ionx_Zflag = P.I(d->eqionx_rxn_Z_flag, h.nionx); h.o_ionx_to_surf = P.I(d->eqionx_rxn_to_surf, h.nionx);
  }
  for (int r = 0; r < h.nkd; ++r) {
    if (d->eqkdspecid[r] < 1 || d->eqkdspecid[r] > naq) return R.fail(RXN_ERR_INVALID, "KD reaction %d: species id out of range", r + 1);
    if (d->eqkdmineral && d->eqkdmineral[r] > nk) return R.fail(RXN_ERR_INVALID, "KD reaction %d: mineral id out of range", r + 1);
  }
  h.o_kd_spec = P.I(d->eqkdspecid, h.nkd); h.o_kd_type = P.I(d->eqkdtype, h.nkd); h.o_kd_mnrl = P.I(d->eqkdmineral, h.nkd);
  h.o_kd_coef = P.D(d->eqkddistcoef, h.nkd); h.o_kd_b = P.D(d->eqkdlangmuirb, h.nkd); h.o_kd_n = P.D(d->eqkdfreundlichn, h.nkd);
  // general reactions / radioactive decay: Fortran (0:m, n) id tables -> CSR with 0-based ids
  {
    auto csr = [&](const int32_t *ids, const double *st, int ld, int nr, int &o_ptr, int &o_id, int &o_st, const char *what) {
      std::vector<int32_t> ptr(1, 0), id;
      std::vector<double> sv;
      for (int r = 0; r < nr; ++r) {
        const int n = ids ? ids[(size_t)r * (ld + 1)] : 0;
        if (n < 0 || n > ld) return R.fail(RXN_ERR_INVALID, "%s %d: species count %d out of range", what, r + 1, n);
        for (int k = 1; k <= n; ++k) {
          const int sp = ids[(size_t)r * (ld + 1) + k];
          if (sp < 1 || sp > naq) return R.fail(RXN_ERR_INVALID, "%s %d: species id %d out of range", what, r + 1, sp);
          id.push_back(sp - 1);
          sv.push_back(st ? st[(size_t)r * ld + k - 1] : 0.0);
        }
        ptr.push_back((int32_t)id.size());
      }
      o_ptr = P.I(ptr.data(), ptr.size()); o_id = P.I(id.data(), id.size()); o_st = P.D(sv.data(), sv.size());
      return (int)RXN_OK;
    };
    h.ngen = d->ngeneral_rxn; h.ndecay = d->nradiodecay_rxn;
    if (h.ngen > 0 && (!d->generalspecid || !d->generalstoich || !d->generalforwardspecid || !d->generalforwardstoich ||
                       !d->generalbackwardspecid || !d->generalbackwardstoich || !d->general_kf || !d->general_kr || d->general_ld < 1))
      return R.fail(RXN_ERR_INVALID, "ngeneral_rxn = %d without the general reaction tables", h.ngen);
    if (h.ndecay > 0 && (!d->radiodecayspecid || !d->radiodecaystoich || !d->radiodecayforwardspecid || !d->radiodecay_kf || d->radiodecay_ld < 1))
      return R.fail(RXN_ERR_INVALID, "nradiodecay_rxn = %d without the radioactive decay tables", h.ndecay);
    RXN_TRY(csr(d->generalspecid, d->generalstoich, d->general_ld, h.ngen, h.o_gen_ptr, h.o_gen_id, h.o_gen_st, "general reaction"));
    RXN_TRY(csr(d->generalforwardspecid, d->generalforwardstoich, d->general_ld, h.ngen, h.o_genf_ptr, h.o_genf_id, h.o_genf_st, "general reaction (forward)"));
    RXN_TRY(csr(d->generalbackwardspecid, d->generalbackwardstoich, d->general_ld, h.ngen, h.o_genb_ptr, h.o_genb_id, h.o_genb_st, "general reaction (backward)"));
    h.o_gen_kf = P.D(d->general_kf, h.ngen); h.o_gen_kr = P.D(d->general_kr, h.ngen);
    RXN_TRY(csr(d->radiodecayspecid, d->radiodecaystoich, d->radiodecay_ld, h.ndecay, h.o_dec_ptr, h.o_dec_id, h.o_dec_st, "radioactive decay reaction"));
    std::vector<int32_t> fwd;
    for (int r = 0; r < h.ndecay; ++r) {
      const int sp = d->radiodecayforwardspecid ? d->radiodecayforwardspecid[r] : 0;
      if (sp < 1 || sp > naq) return R.fail(RXN_ERR_INVALID, "radioactive decay reaction %d: reactant id %d out of range", r + 1, sp);
      fwd.push_back(sp - 1);
    }
    h.o_dec_fwd = P.I(fwd.data(), fwd.size()); h.o_dec_kf = P.D(d->radiodecay_kf, h.ndecay);
  }
  // kinetic surface complexation: RKineticSurfCplx indexes numerator_sum / denominator_sum / srfcplxrxn_site_density /
  // kinsrfcplx_free_site_conc with isite = srfcplxrxn_to_surf(irxn) and the rate tables with the GLOBAL complex id
  // (reaction_surf_complex.F90:1022-1075) although they are dimensioned by kinetic reaction and by position in the reaction.
  // Those indices are in bounds and mean what the routine intends only when the kinetic reaction is surface complexation
  // reaction 1 (its complexes are then 1..n in the master list) on mineral surface 1: anything else is rejected.
  h.nkinrxn = d->nkinsrfcplxrxn; h.nkinsrf = 0; h.kin_rxn = 0;
  if (h.nkinrxn == 1) {
    if (!d->kinsrfcplxrxn_to_srfcplxrxn || !d->kinsrfcplx_forward_rate || !d->kinsrfcplx_backward_rate || h.nrxn < 1)
      return R.fail(RXN_ERR_INVALID, "nkinsrfcplxrxn = 1 without the kinetic surface complexation tables");
    const int ir = d->kinsrfcplxrxn_to_srfcplxrxn ? d->kinsrfcplxrxn_to_srfcplxrxn[0] : 0;
    if (ir != 1 || h.nrxn < 1)
      return R.fail(RXN_ERR_UNSUPPORTED, "the KINETIC surface complexation reaction must be the first surface complexation reaction (it is %d)", ir);
    if (d->srfcplxrxn_surf_type[0] != RXN_MINERAL_SURFACE || d->srfcplxrxn_to_surf[0] != 1)
      return R.fail(RXN_ERR_UNSUPPORTED, "the KINETIC surface complexation reaction must sit on kinetic mineral 1 "
                                         "(RKineticSurfCplx uses the mineral id as site index)");
    const int nc = d->srfcplxrxn_to_complex[0];
    for (int k = 1; k <= nc; ++k)
      if (d->srfcplxrxn_to_complex[k] != k) return R.fail(RXN_ERR_UNSUPPORTED, "complexes of the KINETIC reaction must be complexes 1..n of the master list");
    if (nc > d->kinsrfcplx_ld) return R.fail(RXN_ERR_INVALID, "kinsrfcplx_ld %d < %d complexes", d->kinsrfcplx_ld, nc);
    h.nkinsrf = nc; h.kin_rxn = 0;
    h.o_kin_kf = P.D(d->kinsrfcplx_forward_rate, nc); h.o_kin_kb = P.D(d->kinsrfcplx_backward_rate, nc);
  }
  // immobile decay (reaction_immobile.F90:240-293) and microbial reactions (reaction_microbial.F90:236-450)
  h.nim = d->nimmobile; h.ncomp = d->ncomp; h.nimdecay = d->nimmobile_decay_rxn; h.nmic = d->nmicrobial_rxn;
  {
    std::vector<int32_t> ids;
    if (h.nimdecay > 0 && (!d->immobile_decayspecid || !d->immobile_decay_rate_constant))
      return R.fail(RXN_ERR_INVALID, "nimmobile_decay_rxn = %d without the immobile decay tables", h.nimdecay);
    for (int r = 0; r < h.nimdecay; ++r) {
      const int sp = d->immobile_decayspecid[r];
      if (sp < 1 || sp > h.nim) return R.fail(RXN_ERR_INVALID, "immobile decay reaction %d: immobile species id %d out of range", r + 1, sp);
      ids.push_back(sp - 1);
    }
    h.o_imdec_id = P.I(ids.data(), ids.size()); h.o_imdec_k = P.D(d->immobile_decay_rate_constant, h.nimdecay);
  }
  if (h.nmic > 0) {
    if (!d->microbial_specid || !d->microbial_stoich || !d->microbial_rate_constant || !d->microbial_biomassid || !d->microbial_biomass_yield ||
        !d->microbial_monodid || !d->microbial_inhibitionid || d->microbial_ld < 1 || d->nmicrobial_monod < 0 || d->nmicrobial_inhibition < 0 ||
        (d->nmicrobial_monod > 0 && (!d->microbial_monod_specid || !d->microbial_monod_K || !d->microbial_monod_Cth)) ||
        (d->nmicrobial_inhibition > 0 && (!d->microbial_inhibition_type || !d->microbial_inhibition_specid || !d->microbial_inhibition_C ||
                                          !d->microbial_inhibition_C2)))
      return R.fail(RXN_ERR_INVALID, "nmicrobial_rxn = %d without the microbial tables", h.nmic);
    std::vector<int32_t> ptr(1, 0), id, bio, mptr(1, 0), mid, iptr(1, 0), iid, mspec, ispec;
    std::vector<double> sv;
    for (int r = 0; r < h.nmic; ++r) {
      const int ld = d->microbial_ld, n = d->microbial_specid[(size_t)r * (ld + 1)];
      if (n < 0 || n > ld) return R.fail(RXN_ERR_INVALID, "microbial reaction %d: species count %d out of range", r + 1, n);
      for (int k = 1; k <= n; ++k) {
        const int sp = d->microbial_specid[(size_t)r * (ld + 1) + k];
        if (sp < 1 || sp > h.ncomp) return R.fail(RXN_ERR_INVALID, "microbial reaction %d: species id %d out of range", r + 1, sp);
        id.push_back(sp - 1);
        sv.push_back(d->microbial_stoich[(size_t)r * ld + k - 1]);
      }
      ptr.push_back((int32_t)id.size());
      const int b = d->microbial_biomassid[r];
      if (b < 0 || b > h.nim) return R.fail(RXN_ERR_INVALID, "microbial reaction %d: biomass id %d out of range", r + 1, b);
      bio.push_back(b - 1);
      const int nm = d->microbial_monodid[(size_t)r * (d->microbial_monod_ld + 1)], ni = d->microbial_inhibitionid[(size_t)r * (d->microbial_inhibition_ld + 1)];
      if (nm < 0 || nm > d->microbial_monod_ld || ni < 0 || ni > d->microbial_inhibition_ld)
        return R.fail(RXN_ERR_INVALID, "microbial reaction %d: Monod / inhibition count out of range", r + 1);
      if (nm > RXN_MAX_MONOD || ni > RXN_MAX_MONOD)
        return R.fail(RXN_ERR_UNSUPPORTED, "microbial reaction %d: more than %d Monod or inhibition terms (the reference's monod(10), inhibition(10))", r + 1, (int)RXN_MAX_MONOD);
      for (int k = 1; k <= nm; ++k) {
        const int m = d->microbial_monodid[(size_t)r * (d->microbial_monod_ld + 1) + k];
        if (m < 1 || m > d->nmicrobial_monod) return R.fail(RXN_ERR_INVALID, "microbial reaction %d: Monod id %d out of range", r + 1, m);
        mid.push_back(m - 1);
      }
      mptr.push_back((int32_t)mid.size());
      for (int k = 1; k <= ni; ++k) {
        const int m = d->microbial_inhibitionid[(size_t)r * (d->microbial_inhibition_ld + 1) + k];
        if (m < 1 || m > d->nmicrobial_inhibition) return R.fail(RXN_ERR_INVALID, "microbial reaction %d: inhibition id %d out of range", r + 1, m);
        iid.push_back(m - 1);
      }
      iptr.push_back((int32_t)iid.size());
    }
    for (int k = 0; k < d->nmicrobial_monod; ++k) {
      const int sp = d->microbial_monod_specid[k];
      if (sp < 1 || sp > naq) return R.fail(RXN_ERR_INVALID, "Monod term %d: species id %d out of range", k + 1, sp);
      mspec.push_back(sp - 1);
    }
    for (int k = 0; k < d->nmicrobial_inhibition; ++k) {
      const int sp = d->microbial_inhibition_specid[k], ty = d->microbial_inhibition_type[k];
      if (sp < 1 || sp > naq) return R.fail(RXN_ERR_INVALID, "inhibition term %d: species id %d out of range", k + 1, sp);
      if (ty != RXN_INHIBITION_THRESHOLD && ty != RXN_INHIBITION_MONOD && ty != RXN_INHIBITION_INVERSE_MONOD)
        return R.fail(RXN_ERR_UNSUPPORTED, "inhibition term %d: type %d has no branch in RMicrobial (reaction_microbial.F90:322-339)", k + 1, ty);
      ispec.push_back(sp - 1);
    }
    h.o_mic_ptr = P.I(ptr.data(), ptr.size()); h.o_mic_id = P.I(id.data(), id.size()); h.o_mic_st = P.D(sv.data(), sv.size());
    h.o_mic_bio = P.I(bio.data(), bio.size());
    h.o_mic_mptr = P.I(mptr.data(), mptr.size()); h.o_mic_mid = P.I(mid.data(), mid.size());
    h.o_mic_iptr = P.I(iptr.data(), iptr.size()); h.o_mic_iid = P.I(iid.data(), iid.size());
    h.o_mic_k = P.D(d->microbial_rate_constant, h.nmic);
    h.mic_has_Ea = d->microbial_activation_energy != nullptr; h.o_mic_Ea = P.D(d->microbial_activation_energy, h.nmic);
    h.o_mic_yield = P.D(d->microbial_biomass_yield, h.nmic);
    h.o_mon_spec = P.I(mspec.data(), mspec.size()); h.o_mon_K = P.D(d->microbial_monod_K, d->nmicrobial_monod);
    h.o_mon_Cth = P.D(d->microbial_monod_Cth, d->nmicrobial_monod);
    h.o_inh_spec = P.I(ispec.data(), ispec.size()); h.o_inh_type = P.I(d->microbial_inhibition_type, d->nmicrobial_inhibition);
    h.o_inh_C = P.D(d->microbial_inhibition_C, d->nmicrobial_inhibition); h.o_inh_C2 = P.D(d->microbial_inhibition_C2, d->nmicrobial_inhibition);
  }
#undef RXN_TRY
  if (P.i.size() & 1) P.i.push_back(0);
  h.ndbl = (int)P.d.size(); h.nint = (int)P.i.size();

  const int maxrate = d->kinmr_ld;
  int *r = R.rows;
  r[RXN_F_PRI_MOLAL] = naq; r[RXN_F_TOTAL] = naq; r[RXN_F_SEC_MOLAL] = h.ncplx; r[RXN_F_PRI_ACT_COEF] = naq;
  r[RXN_F_SEC_ACT_COEF] = h.ncplx; r[RXN_F_LN_ACT_H2O] = 1; r[RXN_F_TOTAL_SORB_EQ] = naq;
  r[RXN_F_FREE_SITE_CONC] = h.nrxn; r[RXN_F_EQSRFCPLX_CONC] = h.nsrf;
  r[RXN_F_KINMR_TOTAL_SORB] = h.nmr * (maxrate + 1) * naq;
  r[RXN_F_EQIONX_REF_CATION_SORBED_CONC] = h.nionx; r[RXN_F_EQIONX_CONC] = h.nionx * h.ionx_ld;
  r[RXN_F_MNRL_VOLFRAC] = nk; r[RXN_F_MNRL_AREA] = nk; r[RXN_F_MNRL_RATE] = nk;
  r[RXN_F_DEN_KG] = r[RXN_F_SAT] = r[RXN_F_TEMP] = r[RXN_F_PRES] = r[RXN_F_VOLUME] = r[RXN_F_POROSITY] =
      r[RXN_F_SOIL_PARTICLE_DENSITY] = 1;
  r[RXN_F_DTOTAL] = naq * naq; r[RXN_F_DTOTAL_SORB_EQ] = naq * naq;
  r[RXN_F_KINSRFCPLX_CONC] = h.nkinsrf; r[RXN_F_KINSRFCPLX_CONC_KP1] = h.nkinsrf; r[RXN_F_KINSRFCPLX_FREE_SITE_CONC] = h.nkinrxn;
  r[RXN_F_IMMOBILE] = h.nim;
  return RXN_OK;
}

inline std::vector<unsigned char> blob_bytes(const PackResult &R) {
  std::vector<unsigned char> blob((size_t)R.h.ndbl * 8 + (size_t)R.h.nint * 4);
  memcpy(blob.data(), R.P.d.data(), (size_t)R.h.ndbl * 8);
  memcpy(blob.data() + (size_t)R.h.ndbl * 8, R.P.i.data(), (size_t)R.h.nint * 4);
  return blob;
}

inline int variant_for(int naq) { return naq <= 4 ? 4 : naq <= 8 ? 8 : naq <= 16 ? 16 : 24; }

}  // namespace rxn
