// rxn_oracle.cpp — CPU ORACLE for the PFLOTRAN per-cell reaction path.
//
// TEST INFRASTRUCTURE ONLY.  This is a scalar C++ restatement of the reference's
// Fortran arithmetic, used as the checker for the CUDA path and as the timed CPU
// baseline in bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  The product
// (pflotran_b200/csrc) never links, includes or calls anything in oracle/.
//
// Parity pinning: the restatement is checked against the reference's own
// 14-digit regression gold files (tests/test_oracle_gold.py; list in DESIGN.md).
// RReact itself is dead code in the reference snapshot and has no gold file; it is
// pinned indirectly (every routine it calls is pinned) — see DESIGN.md "Oracle".
//
// Each function cites the reference file:line it follows (paths relative to the
// reference root, src/pflotran/...).  Loop order, operation order and constants
// are the reference's (LOG_TO_LN is the TRUNCATED 2.30258509299d0 of
// pflotran_constants.F90:48).  Compile with -O2 -ffp-contract=off, no fast-math.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

#include "../include/rxn_b200.h"

namespace {

// Scratch of the per-cell routines: the Fortran uses automatic (stack) arrays (e.g. reaction.F90:3339-3346,
// utility.F90:405); so does the port - the heap is only touched by chemistries beyond CAP entries.
template <class T, int CAP> struct Stk {
  T buf[CAP];
  T *p;
  size_t n;
  explicit Stk(size_t n_, T v = T()) : p(n_ <= (size_t)CAP ? buf : new T[n_]), n(n_) { for (size_t i = 0; i < n; ++i) p[i] = v; }
  ~Stk() { if (p != buf) delete[] p; }
  Stk(const Stk &) = delete;
  Stk &operator=(const Stk &) = delete;
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
  T *data() { return p; }
  T *begin() { return p; }
  T *end() { return p + n; }
  size_t size() const { return n; }
};

const double LOG_TO_LN = 2.30258509299;     // pflotran_constants.F90:48
const double IDEAL_GAS_CONSTANT = 8.31446;  // pflotran_constants.F90:53

// transport_constraint.F90:20-29
enum { CONSTRAINT_NULL = 0, CONSTRAINT_FREE = 1, CONSTRAINT_TOTAL = 2, CONSTRAINT_LOG = 3,
       CONSTRAINT_PH = 4, CONSTRAINT_MINERAL = 5, CONSTRAINT_GAS = 6, CONSTRAINT_CHARGE_BAL = 7,
       CONSTRAINT_TOTAL_SORB = 9 };

struct SpecList {
  int n = 0;
  std::vector<std::vector<int>> id;      // 0-based primary ids
  std::vector<std::vector<double>> st;
  std::vector<int> h2oid;
  std::vector<double> h2ost, logK;
  std::vector<std::vector<double>> coef;
  void load(const RxnSpecList &s, int ncoef) {
    n = s.n;
    id.assign(n, {}); st.assign(n, {}); h2oid.assign(n, 0); h2ost.assign(n, 0.0);
    logK.assign(n, 0.0); coef.assign(n, {});
    for (int r = 0; r < n; ++r) {
      int ns = s.id[r * s.id_ld];
      for (int i = 1; i <= ns; ++i) {
        id[r].push_back(s.id[r * s.id_ld + i] - 1);
        st[r].push_back(s.stoich[r * s.stoich_ld + i - s.stoich_off]);
      }
      if (s.h2oid) h2oid[r] = s.h2oid[r];
      if (s.h2ostoich) h2ost[r] = s.h2ostoich[r];
      if (s.logK) logK[r] = s.logK[r];
      if (s.logKcoef && ncoef > 0) coef[r].assign(s.logKcoef + (size_t)r * ncoef, s.logKcoef + (size_t)(r + 1) * ncoef);
    }
  }
};

struct Tables {
  int naq = 0, ncomp = 0, logK_mode = 0, ncoef = 0;
  bool use_log = false, use_act_h2o = false;
  int act_freq = 0, act_alg = 3, h2o_aq_id = 0, h_ion_id = 0;
  double debyeA = 0, debyeB = 0, debyeBdot = 0, max_dlnC = 5, rel_tol = 1e-6, res_tol = 1e-12;
  std::vector<double> Z, a0;
  SpecList cplx; std::vector<double> cplx_Z, cplx_a0;
  SpecList kin;
  std::vector<double> k_rate, k_Ea, k_molar_vol, k_aff_thresh, k_rate_lim, k_Temkin, k_scale, k_power;
  bool has_Temkin = false, has_scale = false, has_power = false;
  std::vector<int> k_npref; int maxpref = 0, maxprefspec = 0;
  std::vector<double> pref_rate, pref_Ea, pref_alpha, pref_beta, pref_atten; std::vector<int> pref_id;
  SpecList mnrl, gas;
  SpecList srf; std::vector<double> srf_site_stoich, srf_Z;
  int nrxn = 0; std::vector<int> rxn_to_surf, rxn_surf_type, rxn_stoich_flag; std::vector<std::vector<int>> rxn_cplx;
  std::vector<double> rxn_site_density;
  std::vector<int> eq_rxn, mr_rxn; std::vector<int> mr_nrate; int mr_ld = 0; std::vector<double> mr_rate, mr_frac;
  int nionx = 0, ionx_ld = 0; std::vector<std::vector<int>> ionx_cat; std::vector<std::vector<double>> ionx_k;
  std::vector<double> ionx_CEC; std::vector<int> ionx_Zflag, ionx_to_surf;
  int nkd = 0; std::vector<int> kd_spec, kd_type, kd_mnrl; std::vector<double> kd_coef, kd_b, kd_n;
  // general / radioactive decay reactions: species lists (0-based ids) as the reference's (0:m,n) tables
  int ngen = 0, ndecay = 0;
  std::vector<std::vector<int>> gen_id, genf_id, genb_id, dec_id; std::vector<std::vector<double>> gen_st, genf_st, genb_st, dec_st;
  std::vector<double> gen_kf, gen_kr, dec_kf; std::vector<int> dec_fwd;
  // kinetic surface complexation (one reaction): its srfcplxrxn (0-based), rates per complex
  int nkinrxn = 0; std::vector<int> kin_rxn; std::vector<double> kin_kf, kin_kb; int kin_ld = 0;
  int nkinsrf() const { return nkinrxn ? (int)rxn_cplx[kin_rxn[0]].size() : 0; }
  // immobile species (dofs naq .. ncomp-1) and their decay; microbial reactions (0-based ids; species ids over the ncomp dofs)
  int nim = 0, nimdecay = 0; std::vector<int> imdec_id; std::vector<double> imdec_k;
  int nmic = 0; bool mic_has_Ea = false;
  std::vector<std::vector<int>> mic_id, mic_monod, mic_inhib; std::vector<std::vector<double>> mic_st;
  std::vector<double> mic_k, mic_Ea, mic_yield; std::vector<int> mic_biomass;   // mic_biomass: 0-based immobile id or -1
  std::vector<int> monod_spec, inhib_spec, inhib_type; std::vector<double> monod_K, monod_Cth, inhib_C, inhib_C2;
  int neqsorb() const { return nionx + nkd + (int)eq_rxn.size(); }
  int nkinmr() const { return (int)mr_rxn.size(); }
};

template <class T> void cp(std::vector<T> &v, const T *p, size_t n) { if (p && n) v.assign(p, p + n); else v.assign(n, T()); }

Tables *load_tables(const RxnTablesDesc *d) {
  Tables *t = new Tables();
  t->naq = d->naqcomp; t->ncomp = d->ncomp; t->logK_mode = d->logK_mode; t->ncoef = d->num_logK_coef;
  t->use_log = d->use_log_formulation != 0; t->use_act_h2o = d->use_activity_h2o != 0;
  t->act_freq = d->act_coef_update_frequency; t->act_alg = d->act_coef_update_algorithm;
  t->h2o_aq_id = d->h2o_aq_id; t->h_ion_id = d->h_ion_id;
  t->debyeA = d->debyeA; t->debyeB = d->debyeB; t->debyeBdot = d->debyeBdot;
  t->max_dlnC = d->max_dlnC; t->rel_tol = d->max_relative_change_tolerance; t->res_tol = d->max_residual_tolerance;
  cp(t->Z, d->primary_spec_Z, t->naq); cp(t->a0, d->primary_spec_a0, t->naq);
  t->cplx.load(d->eqcplx, t->ncoef); cp(t->cplx_Z, d->eqcplx_Z, t->cplx.n); cp(t->cplx_a0, d->eqcplx_a0, t->cplx.n);
  t->kin.load(d->kinmnrl, t->ncoef);
  int nk = t->kin.n;
  cp(t->k_rate, d->kinmnrl_rate_constant, nk); cp(t->k_Ea, d->kinmnrl_activation_energy, nk);
  cp(t->k_molar_vol, d->kinmnrl_molar_vol, nk); cp(t->k_aff_thresh, d->kinmnrl_affinity_threshold, nk);
  cp(t->k_rate_lim, d->kinmnrl_rate_limiter, nk);
  t->has_Temkin = d->kinmnrl_Temkin_const != nullptr; cp(t->k_Temkin, d->kinmnrl_Temkin_const, nk);
  t->has_scale = d->kinmnrl_min_scale_factor != nullptr; cp(t->k_scale, d->kinmnrl_min_scale_factor, nk);
  t->has_power = d->kinmnrl_affinity_power != nullptr; cp(t->k_power, d->kinmnrl_affinity_power, nk);
  cp(t->k_npref, d->kinmnrl_num_prefactors, nk);
  t->maxpref = d->max_num_prefactors; t->maxprefspec = d->max_num_prefactor_species;
  size_t np = (size_t)nk * std::max(t->maxpref, 1);
  cp(t->pref_rate, d->kinmnrl_pref_rate, np); cp(t->pref_Ea, d->kinmnrl_pref_activation_energy, np);
  cp(t->pref_id, d->kinmnrl_prefactor_id, np * (t->maxprefspec + 1));
  cp(t->pref_alpha, d->kinmnrl_pref_alpha, np * std::max(t->maxprefspec, 1));
  cp(t->pref_beta, d->kinmnrl_pref_beta, np * std::max(t->maxprefspec, 1));
  cp(t->pref_atten, d->kinmnrl_pref_atten_coef, np * std::max(t->maxprefspec, 1));
  t->mnrl.load(d->mnrl, 0); t->gas.load(d->paseq, 0);
  t->srf.load(d->srfcplx, t->ncoef); cp(t->srf_site_stoich, d->srfcplx_free_site_stoich, t->srf.n); cp(t->srf_Z, d->srfcplx_Z, t->srf.n);
  t->nrxn = d->nsrfcplxrxn;
  cp(t->rxn_to_surf, d->srfcplxrxn_to_surf, t->nrxn); cp(t->rxn_surf_type, d->srfcplxrxn_surf_type, t->nrxn);
  cp(t->rxn_stoich_flag, d->srfcplxrxn_stoich_flag, t->nrxn); cp(t->rxn_site_density, d->srfcplxrxn_site_density, t->nrxn);
  t->rxn_cplx.assign(t->nrxn, {});
  for (int r = 0; r < t->nrxn; ++r) {
    int ld = d->srfcplxrxn_to_complex_ld, n = d->srfcplxrxn_to_complex[r * ld];
    for (int k = 1; k <= n; ++k) t->rxn_cplx[r].push_back(d->srfcplxrxn_to_complex[r * ld + k] - 1);
  }
  for (int i = 0; i < d->neqsrfcplxrxn; ++i) t->eq_rxn.push_back(d->eqsrfcplxrxn_to_srfcplxrxn[i] - 1);
  for (int i = 0; i < d->nkinmrsrfcplxrxn; ++i) t->mr_rxn.push_back(d->kinmrsrfcplxrxn_to_srfcplxrxn[i] - 1);
  t->mr_ld = d->kinmr_ld;
  for (int i = 0; i < d->nkinmrsrfcplxrxn; ++i) t->mr_nrate.push_back(d->kinmr_nrate[i + 1]);
  cp(t->mr_rate, d->kinmr_rate, (size_t)d->nkinmrsrfcplxrxn * t->mr_ld);
  cp(t->mr_frac, d->kinmr_frac, (size_t)d->nkinmrsrfcplxrxn * t->mr_ld);
  t->nionx = d->neqionxrxn; t->ionx_ld = d->eqionx_ld;
  t->ionx_cat.assign(t->nionx, {}); t->ionx_k.assign(t->nionx, {});
  for (int r = 0; r < t->nionx; ++r) {
    int ld = t->ionx_ld + 1, n = d->eqionx_rxn_cationid[r * ld];
    for (int k = 1; k <= n; ++k) {
      t->ionx_cat[r].push_back(d->eqionx_rxn_cationid[r * ld + k] - 1);
      t->ionx_k[r].push_back(d->eqionx_rxn_k[r * t->ionx_ld + k - 1]);
    }
  }
  cp(t->ionx_CEC, d->eqionx_rxn_CEC, t->nionx); cp(t->ionx_Zflag, d->eqionx_rxn_Z_flag, t->nionx);
  cp(t->ionx_to_surf, d->eqionx_rxn_to_surf, t->nionx);
  t->nkd = d->neqkdrxn;
  cp(t->kd_spec, d->eqkdspecid, t->nkd); cp(t->kd_type, d->eqkdtype, t->nkd); cp(t->kd_mnrl, d->eqkdmineral, t->nkd);
  cp(t->kd_coef, d->eqkddistcoef, t->nkd); cp(t->kd_b, d->eqkdlangmuirb, t->nkd); cp(t->kd_n, d->eqkdfreundlichn, t->nkd);
  auto lists = [](const int32_t *ids, const double *st, int ld, int nr, std::vector<std::vector<int>> &oid, std::vector<std::vector<double>> &ost) {
    oid.assign(nr, {}); ost.assign(nr, {});
    for (int r = 0; r < nr; ++r) {
      int n = ids[(size_t)r * (ld + 1)];
      for (int k = 1; k <= n; ++k) { oid[r].push_back(ids[(size_t)r * (ld + 1) + k] - 1); ost[r].push_back(st[(size_t)r * ld + k - 1]); }
    }
  };
  t->ngen = d->ngeneral_rxn; t->ndecay = d->nradiodecay_rxn;
  if (t->ngen > 0) {
    lists(d->generalspecid, d->generalstoich, d->general_ld, t->ngen, t->gen_id, t->gen_st);
    lists(d->generalforwardspecid, d->generalforwardstoich, d->general_ld, t->ngen, t->genf_id, t->genf_st);
    lists(d->generalbackwardspecid, d->generalbackwardstoich, d->general_ld, t->ngen, t->genb_id, t->genb_st);
    cp(t->gen_kf, d->general_kf, t->ngen); cp(t->gen_kr, d->general_kr, t->ngen);
  }
  if (t->ndecay > 0) {
    lists(d->radiodecayspecid, d->radiodecaystoich, d->radiodecay_ld, t->ndecay, t->dec_id, t->dec_st);
    cp(t->dec_kf, d->radiodecay_kf, t->ndecay);
    for (int r = 0; r < t->ndecay; ++r) t->dec_fwd.push_back(d->radiodecayforwardspecid[r] - 1);
  }
  t->nkinrxn = d->nkinsrfcplxrxn;
  if (t->nkinrxn > 0) {
    for (int i = 0; i < t->nkinrxn; ++i) t->kin_rxn.push_back(d->kinsrfcplxrxn_to_srfcplxrxn[i] - 1);
    t->kin_ld = d->kinsrfcplx_ld;
    cp(t->kin_kf, d->kinsrfcplx_forward_rate, (size_t)t->kin_ld * t->nkinrxn);
    cp(t->kin_kb, d->kinsrfcplx_backward_rate, (size_t)t->kin_ld * t->nkinrxn);
  }
  t->nim = d->nimmobile; t->nimdecay = d->nimmobile_decay_rxn;
  for (int r = 0; r < t->nimdecay; ++r) { t->imdec_id.push_back(d->immobile_decayspecid[r] - 1); t->imdec_k.push_back(d->immobile_decay_rate_constant[r]); }
  t->nmic = d->nmicrobial_rxn;
  if (t->nmic > 0) {
    lists(d->microbial_specid, d->microbial_stoich, d->microbial_ld, t->nmic, t->mic_id, t->mic_st);
    cp(t->mic_k, d->microbial_rate_constant, t->nmic);
    t->mic_has_Ea = d->microbial_activation_energy != nullptr; cp(t->mic_Ea, d->microbial_activation_energy, t->nmic);
    cp(t->mic_yield, d->microbial_biomass_yield, t->nmic);
    t->mic_monod.assign(t->nmic, {}); t->mic_inhib.assign(t->nmic, {});
    for (int r = 0; r < t->nmic; ++r) {
      t->mic_biomass.push_back(d->microbial_biomassid ? d->microbial_biomassid[r] - 1 : -1);
      const int nm = d->microbial_monodid ? d->microbial_monodid[(size_t)r * (d->microbial_monod_ld + 1)] : 0;
      for (int k = 1; k <= nm; ++k) t->mic_monod[r].push_back(d->microbial_monodid[(size_t)r * (d->microbial_monod_ld + 1) + k] - 1);
      const int ni = d->microbial_inhibitionid ? d->microbial_inhibitionid[(size_t)r * (d->microbial_inhibition_ld + 1)] : 0;
      for (int k = 1; k <= ni; ++k) t->mic_inhib[r].push_back(d->microbial_inhibitionid[(size_t)r * (d->microbial_inhibition_ld + 1) + k] - 1);
    }
    for (int k = 0; k < d->nmicrobial_monod; ++k) t->monod_spec.push_back(d->microbial_monod_specid[k] - 1);
    cp(t->monod_K, d->microbial_monod_K, d->nmicrobial_monod); cp(t->monod_Cth, d->microbial_monod_Cth, d->nmicrobial_monod);
    for (int k = 0; k < d->nmicrobial_inhibition; ++k) t->inhib_spec.push_back(d->microbial_inhibition_specid[k] - 1);
    cp(t->inhib_type, d->microbial_inhibition_type, d->nmicrobial_inhibition);
    cp(t->inhib_C, d->microbial_inhibition_C, d->nmicrobial_inhibition); cp(t->inhib_C2, d->microbial_inhibition_C2, d->nmicrobial_inhibition);
  }
  return t;
}

// reactive_transport_auxvar_type + the global/material scalars the path reads
struct AuxVar {
  std::vector<double> pri_molal, total, sec_molal, pri_act_coef, sec_act_coef;
  std::vector<double> dtotal;            // naq x naq column-major
  double ln_act_h2o = 0.0;
  std::vector<double> total_sorb_eq, dtotal_sorb_eq, free_site_conc, eqsrfcplx_conc;
  std::vector<double> kinmr_total_sorb;  // [rxn][rate 0..maxrate][naq]
  std::vector<double> ionx_ref_sorbed, ionx_conc;
  std::vector<double> mnrl_volfrac, mnrl_area, mnrl_rate;
  std::vector<double> kinsrfcplx_conc, kinsrfcplx_conc_kp1, kinsrfcplx_free_site_conc;   // (nkinsrfcplx,1), (nkinsrfcplx,1), (nkinsrfcplxrxn)
  std::vector<double> immobile;          // nimmobile [mol/m^3 bulk]
  double den_kg = 0, sat = 0, temp = 0, pres = 0, volume = 0, porosity = 0, soil_density = 0;
  int flags = 0;
};

void init_auxvar(const Tables &t, AuxVar &a) {
  int n = t.naq;
  a.pri_molal.assign(n, 0); a.total.assign(n, 0); a.sec_molal.assign(t.cplx.n, 0);
  a.pri_act_coef.assign(n, 1.0); a.sec_act_coef.assign(t.cplx.n, 1.0);
  a.dtotal.assign((size_t)n * n, 0); a.total_sorb_eq.assign(n, 0); a.dtotal_sorb_eq.assign((size_t)n * n, 0);
  a.free_site_conc.assign(t.nrxn, 1e-9); a.eqsrfcplx_conc.assign(t.srf.n, 0);
  a.kinmr_total_sorb.assign((size_t)t.nkinmr() * (t.mr_ld + 1) * n, 0);
  a.ionx_ref_sorbed.assign(t.nionx, 1e-9); a.ionx_conc.assign((size_t)t.nionx * std::max(t.ionx_ld, 1), 0);
  a.mnrl_volfrac.assign(t.kin.n, 0); a.mnrl_area.assign(t.kin.n, 0); a.mnrl_rate.assign(t.kin.n, 0);
  a.kinsrfcplx_conc.assign(t.nkinsrf(), 0); a.kinsrfcplx_conc_kp1.assign(t.nkinsrf(), 0); a.kinsrfcplx_free_site_conc.assign(t.nkinrxn, 0);
  a.immobile.assign(t.nim, 0);
}

// ---------------------------------------------------------------- utility.F90:393-476
int ludcmp(double *A, int N, int *indx) {  // A column-major N x N; returns 1 on all-zero row
  const double tiny = 1.0e-20;
  Stk<double, 32> vv(N);
#define AA(i, j) A[(i) + (size_t)(j) * N]
  for (int i = 0; i < N; ++i) {
    double aamax = 0.0;
    for (int j = 0; j < N; ++j) if (std::fabs(AA(i, j)) > aamax) aamax = std::fabs(AA(i, j));
    if (aamax <= 0.0) return 1;
    vv[i] = 1.0 / aamax;
  }
  int imax = 0;
  for (int j = 0; j < N; ++j) {
    for (int i = 0; i < j; ++i) {
      double sum = AA(i, j);
      for (int k = 0; k < i; ++k) sum = sum - AA(i, k) * AA(k, j);
      AA(i, j) = sum;
    }
    double aamax = 0.0;
    for (int i = j; i < N; ++i) {
      double sum = AA(i, j);
      for (int k = 0; k < j; ++k) sum = sum - AA(i, k) * AA(k, j);
      AA(i, j) = sum;
      double dum = vv[i] * std::fabs(sum);
      if (dum >= aamax) { imax = i; aamax = dum; }
    }
    if (j != imax) {
      for (int k = 0; k < N; ++k) { double dum = AA(imax, k); AA(imax, k) = AA(j, k); AA(j, k) = dum; }
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (AA(j, j) == 0.0) AA(j, j) = tiny;
    if (j != N - 1) {
      double dum = 1.0 / AA(j, j);
      for (int i = j + 1; i < N; ++i) AA(i, j) = AA(i, j) * dum;
    }
  }
  return 0;
}

// ---------------------------------------------------------------- utility.F90:480-523
void lubksb(const double *A, int N, const int *indx, double *B) {
  int ii = -1;
  for (int i = 0; i < N; ++i) {
    int ll = indx[i];
    double sum = B[ll];
    B[ll] = B[i];
    if (ii != -1) {
      for (int j = ii; j < i; ++j) sum = sum - AA(i, j) * B[j];
    } else if (sum != 0.0) {
      ii = i;
    }
    B[i] = sum;
  }
  for (int i = N - 1; i >= 0; --i) {
    double sum = B[i];
    for (int j = i + 1; j < N; ++j) sum = sum - AA(i, j) * B[j];
    B[i] = sum / AA(i, i);
  }
#undef AA
}

// ---------------------------------------------------------------- reaction.F90:4835-4880
int RSolve(double *Res, double *Jac, const double *conc, double *update, int n, bool use_log) {
  Stk<int, 32> indices(n);
  Stk<double, 32> rhs(n);
  for (int i = 0; i < n; ++i) {
    double mx = 0.0;
    for (int j = 0; j < n; ++j) mx = std::max(mx, std::fabs(Jac[i + (size_t)j * n]));
    double norm = std::max(1.0, mx);
    norm = 1.0 / norm;
    rhs[i] = Res[i] * norm;
    for (int j = 0; j < n; ++j) Jac[i + (size_t)j * n] = Jac[i + (size_t)j * n] * norm;
  }
  if (use_log) {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) Jac[i + (size_t)j * n] = Jac[i + (size_t)j * n] * conc[j];
  }
  if (ludcmp(Jac, n, indices.data())) return 1;
  lubksb(Jac, n, indices.data(), rhs.data());
  for (int i = 0; i < n; ++i) update[i] = rhs[i];
  return 0;
}

// ------------------------------------------ reaction_aux.F90:1461-1488 / :1529-1571
double interp_logK(const std::vector<double> &c, double temp) {
  double tk = temp + 273.15;
  return c[0] * std::log(tk) + c[1] + c[2] * tk + c[3] / tk + c[4] / (tk * tk);
}
double interp_logK_hpt(const std::vector<double> &c, double temp, double pres) {
  double tk = temp + 273.15, tr = tk / 273.15, pr = pres / 1.0e7;
  double logtr = std::log(tr) / std::log(10.0);
  return c[0] + c[1] * tr + c[2] / tr + c[3] * logtr + c[4] * tr * tr + c[5] / tr / tr +
         c[6] * std::sqrt(tr) + c[7] * pr + c[8] * pr * tr + c[9] * pr / tr + c[10] * pr * logtr +
         c[11] / pr + c[12] / pr * tr + c[13] / pr / tr + c[14] * pr * pr + c[15] * pr * pr * tr +
         c[16] * pr * pr / tr;
}

// ---------------------------------------------------------------- reaction.F90:5433-5524
// The reference overwrites the shared tables per cell; here `t` is the calling thread's copy.
void RUpdateTempDependentCoefs(Tables &t, const AuxVar &a) {
  if (t.logK_mode == RXN_LOGK_FIXED) return;
  auto upd = [&](SpecList &s) {
    for (int r = 0; r < s.n; ++r) {
      if (s.coef[r].empty()) continue;
      s.logK[r] = (t.logK_mode == RXN_LOGK_HPT) ? interp_logK_hpt(s.coef[r], a.temp, a.pres)
                                                 : interp_logK(s.coef[r], a.temp);
    }
  };
  upd(t.cplx);
  upd(t.kin);   // MineralUpdateTempDepCoefs, reaction_mineral.F90:1227-1278
  if (t.logK_mode == RXN_LOGK_FIT5) upd(t.srf);  // hpt: not implemented in the reference (:5517-5521)
}

// ---------------------------------------------------------------- reaction.F90:3812-4053
void RActivityCoefficients(const Tables &t, AuxVar &a) {
  const int naq = t.naq, ncplx = t.cplx.n;
  double sum_pri_molal = 0.0;
  if (t.use_act_h2o) {
    for (int j = 0; j < naq; ++j)
      if (j + 1 != t.h2o_aq_id) sum_pri_molal = sum_pri_molal + a.pri_molal[j];
  }
  if (t.act_alg == RXN_ACT_COEF_ALGORITHM_NEWTON) {
    Stk<double, 32> ln_conc(naq), ln_act(naq);
    for (int j = 0; j < naq; ++j) { ln_conc[j] = std::log(a.pri_molal[j]); ln_act[j] = ln_conc[j] + std::log(a.pri_act_coef[j]); }
    double fpri = 0.0;
    for (int j = 0; j < naq; ++j) fpri = fpri + a.pri_molal[j] * t.Z[j] * t.Z[j];
    int it = 0;
    double II = 0.0, I = 0.0, f = 0.0;
    for (;;) {
      it = it + 1;
      if (it > 50) {
        double NaN = std::nan("");
        for (auto &x : a.pri_molal) x = NaN;
        for (auto &x : a.pri_act_coef) x = NaN;
        for (auto &x : a.sec_act_coef) x = NaN;
        a.flags |= RXN_FLAG_ACT_DIVERGED;
        return;  // the reference keeps looping on NaNs until the job dies
      }
      I = fpri;
      for (int k = 0; k < ncplx; ++k) I = I + a.sec_molal[k] * t.cplx_Z[k] * t.cplx_Z[k];
      I = 0.5 * I;
      f = I;
      if (std::fabs(I - II) < 1.0e-6 * I) break;
      if (ncplx > 0) {
        double didi = 0.0;
        double sqrt_I = std::sqrt(I);
        for (int k = 0; k < ncplx; ++k) {
          if (std::fabs(t.cplx_Z[k]) > 0.0) {
            double tmp = 1.0 + t.debyeB * t.cplx_a0[k] * sqrt_I;
            double sum = 0.5 * t.debyeA * t.cplx_Z[k] * t.cplx_Z[k] / (sqrt_I * (tmp * tmp)) - t.debyeBdot;
            for (size_t jc = 0; jc < t.cplx.id[k].size(); ++jc) {
              int j = t.cplx.id[k][jc];
              if (std::fabs(t.Z[j]) > 0.0) {
                double tp = 1.0 + t.debyeB * t.a0[j] * sqrt_I;
                double dgamdi = -0.5 * t.debyeA * (t.Z[j] * t.Z[j]) / (sqrt_I * (tp * tp)) + t.debyeBdot;
                sum = sum + t.cplx.st[k][jc] * dgamdi;
              }
            }
            double dcdi = a.sec_molal[k] * LOG_TO_LN * sum;
            didi = didi + 0.5 * t.cplx_Z[k] * t.cplx_Z[k] * dcdi;
          }
        }
        double den = 1.0 - didi;
        if (std::fabs(den) > 0.0) II = (f - I * didi) / den; else II = f;
      } else {
        II = f;
      }
      I = II;
      double sqrt_I = std::sqrt(I);
      for (int i = 0; i < naq; ++i) {
        if (std::fabs(t.Z[i]) > 0.0)
          a.pri_act_coef[i] = std::exp((-t.Z[i] * t.Z[i] * sqrt_I * t.debyeA / (1.0 + t.a0[i] * t.debyeB * sqrt_I) + t.debyeBdot * I) * LOG_TO_LN);
        else
          a.pri_act_coef[i] = 1.0;
      }
      double sum_sec_molal = 0.0;
      for (int k = 0; k < ncplx; ++k) {
        if (std::fabs(t.cplx_Z[k]) > 0.0)
          a.sec_act_coef[k] = std::exp((-t.cplx_Z[k] * t.cplx_Z[k] * sqrt_I * t.debyeA / (1.0 + t.cplx_a0[k] * t.debyeB * sqrt_I) + t.debyeBdot * I) * LOG_TO_LN);
        else
          a.sec_act_coef[k] = 1.0;
        double lnQK = -t.cplx.logK[k] * LOG_TO_LN;
        if (t.cplx.h2oid[k] > 0) lnQK = lnQK + t.cplx.h2ost[k] * a.ln_act_h2o;
        for (size_t jc = 0; jc < t.cplx.id[k].size(); ++jc) lnQK = lnQK + t.cplx.st[k][jc] * ln_act[t.cplx.id[k][jc]];
        a.sec_molal[k] = std::exp(lnQK) / a.sec_act_coef[k];
        sum_sec_molal = sum_sec_molal + a.sec_molal[k];
      }
      if (t.use_act_h2o) {
        a.ln_act_h2o = 1.0 - 0.017 * (sum_pri_molal + sum_sec_molal);
        if (a.ln_act_h2o > 0.0) a.ln_act_h2o = std::log(a.ln_act_h2o); else a.ln_act_h2o = 0.0;
      }
    }
  } else {
    double I = 0.0;
    for (int i = 0; i < naq; ++i) I = I + a.pri_molal[i] * t.Z[i] * t.Z[i];
    for (int k = 0; k < ncplx; ++k) I = I + a.sec_molal[k] * t.cplx_Z[k] * t.cplx_Z[k];
    I = 0.5 * I;
    double sqrt_I = std::sqrt(I);
    for (int i = 0; i < naq; ++i) {
      if (std::fabs(t.Z[i]) > 1.0e-10)
        a.pri_act_coef[i] = std::exp((-t.Z[i] * t.Z[i] * sqrt_I * t.debyeA / (1.0 + t.a0[i] * t.debyeB * sqrt_I) + t.debyeBdot * I) * LOG_TO_LN);
      else
        a.pri_act_coef[i] = 1.0;
    }
    double sum_sec_molal = 0.0;
    for (int k = 0; k < ncplx; ++k) {
      if (std::fabs(t.cplx_Z[k]) > 1.0e-10)
        a.sec_act_coef[k] = std::exp((-t.cplx_Z[k] * t.cplx_Z[k] * sqrt_I * t.debyeA / (1.0 + t.cplx_a0[k] * t.debyeB * sqrt_I) + t.debyeBdot * I) * LOG_TO_LN);
      else
        a.sec_act_coef[k] = 1.0;
      sum_sec_molal = sum_sec_molal + a.sec_molal[k];
    }
    if (t.use_act_h2o) {
      a.ln_act_h2o = 1.0 - 0.017 * (sum_pri_molal + sum_sec_molal);
      if (a.ln_act_h2o > 0.0) a.ln_act_h2o = std::log(a.ln_act_h2o); else a.ln_act_h2o = 0.0;
    }
  }
}

// ---------------------------------------------------------------- reaction.F90:4057-4158
void RTotal(const Tables &t, AuxVar &a) {
  const int naq = t.naq;
  Stk<double, 32> ln_conc(naq), ln_act(naq);
  const double xmass = 1.0;
  const double den_kg_per_L = a.den_kg * xmass * 1.0e-3;
  for (int i = 0; i < naq; ++i) { ln_conc[i] = std::log(a.pri_molal[i]); ln_act[i] = ln_conc[i] + std::log(a.pri_act_coef[i]); }
  for (int i = 0; i < naq; ++i) a.total[i] = a.pri_molal[i];
  std::fill(a.dtotal.begin(), a.dtotal.end(), 0.0);
  for (int i = 0; i < naq; ++i) a.dtotal[i + (size_t)i * naq] = 1.0;
  for (int k = 0; k < t.cplx.n; ++k) {
    double lnQK = -t.cplx.logK[k] * LOG_TO_LN;
    if (t.cplx.h2oid[k] > 0) lnQK = lnQK + t.cplx.h2ost[k] * a.ln_act_h2o;
    const std::vector<int> &id = t.cplx.id[k];
    const std::vector<double> &st = t.cplx.st[k];
    const int ncomp = (int)id.size();
    for (int i = 0; i < ncomp; ++i) lnQK = lnQK + st[i] * ln_act[id[i]];
    a.sec_molal[k] = std::exp(lnQK) / a.sec_act_coef[k];
    for (int i = 0; i < ncomp; ++i) a.total[id[i]] = a.total[id[i]] + st[i] * a.sec_molal[k];
    for (int j = 0; j < ncomp; ++j) {
      int jcomp = id[j];
      double tempreal = st[j] * std::exp(lnQK - ln_conc[jcomp]) / a.sec_act_coef[k];
      for (int i = 0; i < ncomp; ++i) {
        int icomp = id[i];
        a.dtotal[icomp + (size_t)jcomp * naq] = a.dtotal[icomp + (size_t)jcomp * naq] + st[i] * tempreal;
      }
    }
  }
  for (int i = 0; i < naq; ++i) a.total[i] = a.total[i] * den_kg_per_L;
  for (auto &x : a.dtotal) x = x * den_kg_per_L;
}

// ------------------------------------------------- reaction_surf_complex.F90:658-934
void RTotalSorbEqSurfCplx1(const Tables &t, AuxVar &a, int irxn, double &external_free_site_conc,
                           double *external_srfcplx_conc, double *external_total_sorb,
                           double *external_dtotal_sorb) {
  const int naq = t.naq;
  const double tol = 1.0e-12;
  Stk<double, 32> ln_conc(naq), ln_act(naq), dSx_dmi(naq);
  Stk<double, 256> srfcplx_conc(t.srf.n, 0.0);
  for (int i = 0; i < naq; ++i) { ln_conc[i] = std::log(a.pri_molal[i]); ln_act[i] = ln_conc[i] + std::log(a.pri_act_coef[i]); }
  const std::vector<int> &cl = t.rxn_cplx[irxn];
  const int ncplx = (int)cl.size();
  double free_site_conc = external_free_site_conc;
  double site_density = 0.0;
  switch (t.rxn_surf_type[irxn]) {
    case RXN_MINERAL_SURFACE: site_density = t.rxn_site_density[irxn] * a.mnrl_volfrac[t.rxn_to_surf[irxn] - 1]; break;
    case RXN_ROCK_SURFACE: site_density = t.rxn_site_density[irxn] * a.soil_density * (1.0 - a.porosity); break;
    default: site_density = t.rxn_site_density[irxn]; break;  // NULL_SURFACE (colloids unsupported)
  }
  if (site_density < 1.0e-40) return;
  bool one_more = false;
  int num_iterations = 0;
  double damping_factor = 1.0;
  double total;
  for (;;) {
    num_iterations = num_iterations + 1;
    total = free_site_conc;
    double ln_free_site = std::log(free_site_conc);
    for (int j = 0; j < ncplx; ++j) {
      int icplx = cl[j];
      double lnQK = -t.srf.logK[icplx] * LOG_TO_LN;
      if (t.srf.h2oid[icplx] > 0) lnQK = lnQK + t.srf.h2ost[icplx] * a.ln_act_h2o;
      lnQK = lnQK + t.srf_site_stoich[icplx] * ln_free_site;
      for (size_t i = 0; i < t.srf.id[icplx].size(); ++i) lnQK = lnQK + t.srf.st[icplx][i] * ln_act[t.srf.id[icplx][i]];
      srfcplx_conc[icplx] = std::exp(lnQK);
      total = total + t.srf_site_stoich[icplx] * srfcplx_conc[icplx];
    }
    if (one_more) break;
    if (t.rxn_stoich_flag[irxn]) {
      double res = site_density - total;
      double dres_dfree_site = 1.0;
      for (int j = 0; j < ncplx; ++j) {
        int icplx = cl[j];
        dres_dfree_site = dres_dfree_site + t.srf_site_stoich[icplx] * srfcplx_conc[icplx] / free_site_conc;
      }
      double dfree_site_conc = res / dres_dfree_site;
      if (num_iterations > 1000) damping_factor = 0.5;
      free_site_conc = free_site_conc + damping_factor * dfree_site_conc;
      double rel_change = std::fabs(dfree_site_conc / free_site_conc);
      if (rel_change < tol) one_more = true;
      if (num_iterations > 100000) { a.flags |= RXN_FLAG_CAPPED; one_more = true; }  // oracle-only guard
    } else {
      total = total / free_site_conc;
      free_site_conc = site_density / total;
      one_more = true;
    }
  }
  external_free_site_conc = free_site_conc;

  std::fill(dSx_dmi.begin(), dSx_dmi.end(), 0.0);
  double tempreal = 0.0;
  for (int j = 0; j < ncplx; ++j) {
    int icplx = cl[j];
    for (size_t i = 0; i < t.srf.id[icplx].size(); ++i) {
      int icomp = t.srf.id[icplx][i];
      dSx_dmi[icomp] = dSx_dmi[icomp] + t.srf.st[icplx][i] * t.srf_site_stoich[icplx] * srfcplx_conc[icplx];
    }
    tempreal = tempreal + t.srf_site_stoich[icplx] * t.srf_site_stoich[icplx] * srfcplx_conc[icplx];
  }
  tempreal = tempreal / free_site_conc;
  tempreal = tempreal + 1.0;
  for (int i = 0; i < naq; ++i) dSx_dmi[i] = -dSx_dmi[i] / tempreal;
  for (int i = 0; i < naq; ++i) dSx_dmi[i] = dSx_dmi[i] / a.pri_molal[i];

  if (external_srfcplx_conc)
    for (int i = 0; i < t.srf.n; ++i) external_srfcplx_conc[i] = external_srfcplx_conc[i] + srfcplx_conc[i];

  for (int k = 0; k < ncplx; ++k) {
    int icplx = cl[k];
    const std::vector<int> &id = t.srf.id[icplx];
    const std::vector<double> &st = t.srf.st[icplx];
    int ncomp = (int)id.size();
    for (int i = 0; i < ncomp; ++i) external_total_sorb[id[i]] = external_total_sorb[id[i]] + st[i] * srfcplx_conc[icplx];
    double nui_Si_over_Sx = t.srf_site_stoich[icplx] * srfcplx_conc[icplx] / free_site_conc;
    for (int j = 0; j < ncomp; ++j) {
      int jcomp = id[j];
      double tr = st[j] * srfcplx_conc[icplx] / a.pri_molal[jcomp] + nui_Si_over_Sx * dSx_dmi[jcomp];
      for (int i = 0; i < ncomp; ++i) {
        int icomp = id[i];
        external_dtotal_sorb[icomp + (size_t)jcomp * naq] = external_dtotal_sorb[icomp + (size_t)jcomp * naq] + st[i] * tr;
      }
    }
  }
}

// ---------------------------------------------------------------- reaction.F90:4305-4535
void RTotalSorbEqIonx(const Tables &t, AuxVar &a) {
  const int naq = t.naq;
  const double tol = 1.0e-12;
  std::fill(a.ionx_conc.begin(), a.ionx_conc.end(), 0.0);
  for (int irxn = 0; irxn < t.nionx; ++irxn) {
    const std::vector<int> &cat = t.ionx_cat[irxn];
    const std::vector<double> &kk = t.ionx_k[irxn];
    int ncomp = (int)cat.size();
    double omega;
    if (t.ionx_to_surf[irxn] > 0) omega = std::max(t.ionx_CEC[irxn] * a.mnrl_volfrac[t.ionx_to_surf[irxn] - 1], 1.0e-40);
    else omega = t.ionx_CEC[irxn];
    Stk<double, 32> cation_X(naq, 0.0);
    if (t.ionx_Zflag[irxn]) {
      int icomp = cat[0];
      double ref_cation_conc = a.pri_molal[icomp] * a.pri_act_coef[icomp];
      double ref_cation_Z = t.Z[icomp];
      double ref_cation_k = kk[0];
      double ref_cation_X = ref_cation_Z * a.ionx_ref_sorbed[irxn] / omega;
      bool one_more = false;
      double KDj = ref_cation_X / (ref_cation_k * ref_cation_conc);
      int it = 0;
      for (;;) {
        it = it + 1;
        if (it > 20000) { a.flags |= RXN_FLAG_CAPPED; break; }
        ref_cation_X = KDj * (ref_cation_k * ref_cation_conc);
        cation_X[0] = ref_cation_X;
        double total = ref_cation_X;
        double dres_dKDj = 0.0;
        for (int j = 1; j < ncomp; ++j) {
          int ic = cat[j];
          cation_X[j] = kk[j] * a.pri_molal[ic] * a.pri_act_coef[ic] * std::pow(KDj, t.Z[ic] / ref_cation_Z);
          total = total + cation_X[j];
          dres_dKDj = dres_dKDj + cation_X[j] / KDj * t.Z[ic];
        }
        dres_dKDj = dres_dKDj / ref_cation_Z + (ref_cation_k * ref_cation_conc);
        double res = 1.0 - total;
        if (one_more) break;
        double delta_KDj = res / dres_dKDj;
        KDj = KDj + delta_KDj;
        KDj = std::max(KDj, 1.0e-40);
        if (std::fabs(delta_KDj / KDj) < tol) one_more = true;
      }
      a.ionx_ref_sorbed[irxn] = ref_cation_X * omega / ref_cation_Z;
    } else {
      double sumkm = 0.0;
      for (int j = 0; j < ncomp; ++j) {
        int ic = cat[j];
        cation_X[j] = a.pri_molal[ic] * a.pri_act_coef[ic] * kk[j];
        sumkm = sumkm + cation_X[j];
      }
      for (int j = 0; j < naq; ++j) cation_X[j] = cation_X[j] / sumkm;
    }
    double sumZX = 0.0;
    for (int i = 0; i < ncomp; ++i) sumZX = sumZX + t.Z[cat[i]] * cation_X[i];
    for (int i = 0; i < ncomp; ++i) {
      int icomp = cat[i];
      double tempreal1 = cation_X[i] * omega / t.Z[icomp];
      a.ionx_conc[(size_t)irxn * t.ionx_ld + i] = a.ionx_conc[(size_t)irxn * t.ionx_ld + i] + tempreal1;
      a.total_sorb_eq[icomp] = a.total_sorb_eq[icomp] + tempreal1;
      double tempreal2 = t.Z[icomp] / sumZX;
      for (int j = 0; j < ncomp; ++j) {
        int jcomp = cat[j];
        size_t e = icomp + (size_t)jcomp * naq;
        if (i == j) a.dtotal_sorb_eq[e] = a.dtotal_sorb_eq[e] + tempreal1 * (1.0 - (tempreal2 * cation_X[j])) / a.pri_molal[jcomp];
        else a.dtotal_sorb_eq[e] = a.dtotal_sorb_eq[e] + (-tempreal1) * tempreal2 * cation_X[j] / a.pri_molal[jcomp];
      }
    }
  }
}

// ---------------------------------------------------------------- reaction.F90:4220-4301
void RTotalSorbKD(const Tables &t, AuxVar &a) {
  const int naq = t.naq;
  for (int irxn = 0; irxn < t.nkd; ++irxn) {
    int icomp = t.kd_spec[irxn] - 1;
    double molality = a.pri_molal[icomp];
    double kd_kgw_m3b;
    if (t.kd_mnrl[irxn] > 0)
      kd_kgw_m3b = t.kd_coef[irxn] * a.den_kg * (1.0 - a.porosity) * a.soil_density * 1.0e-3 * (a.mnrl_volfrac[t.kd_mnrl[irxn] - 1]);
    else
      kd_kgw_m3b = t.kd_coef[irxn];
    double res, dres_dc;
    switch (t.kd_type[irxn]) {
      case RXN_SORPTION_LINEAR: res = kd_kgw_m3b * molality; dres_dc = kd_kgw_m3b; break;
      case RXN_SORPTION_LANGMUIR: {
        double tempreal = kd_kgw_m3b * molality;
        res = tempreal * t.kd_b[irxn] / (1.0 + tempreal);
        dres_dc = res / molality - res / (1.0 + tempreal) * tempreal / molality;
      } break;
      case RXN_SORPTION_FREUNDLICH: {
        double one_over_n = 1.0 / t.kd_n[irxn];
        res = kd_kgw_m3b * std::pow(molality, one_over_n);
        dres_dc = res / molality * one_over_n;
      } break;
      default: res = 0.0; dres_dc = 0.0;
    }
    a.total_sorb_eq[icomp] = a.total_sorb_eq[icomp] + res;
    a.dtotal_sorb_eq[icomp + (size_t)icomp * naq] = a.dtotal_sorb_eq[icomp + (size_t)icomp * naq] + dres_dc;
  }
}

// -------------------------------- reaction.F90:4162-4216 (RZeroSorb, RTotalSorb); surf_complex :441-502
void RTotalSorb(const Tables &t, AuxVar &a) {
  std::fill(a.total_sorb_eq.begin(), a.total_sorb_eq.end(), 0.0);
  std::fill(a.dtotal_sorb_eq.begin(), a.dtotal_sorb_eq.end(), 0.0);
  std::fill(a.eqsrfcplx_conc.begin(), a.eqsrfcplx_conc.end(), 0.0);
  for (size_t ieq = 0; ieq < t.eq_rxn.size(); ++ieq) {
    int irxn = t.eq_rxn[ieq];
    RTotalSorbEqSurfCplx1(t, a, irxn, a.free_site_conc[irxn], a.eqsrfcplx_conc.data(), a.total_sorb_eq.data(), a.dtotal_sorb_eq.data());
  }
  if (t.nionx > 0) RTotalSorbEqIonx(t, a);
  if (t.nkd > 0) RTotalSorbKD(t, a);
}

// ------------------------------------------------- reaction_surf_complex.F90:506-562
void RTotalSorbMultiRateAsEQ(const Tables &t, AuxVar &a) {
  const int naq = t.naq;
  Stk<double, 32> total_sorb_eq(naq);
  Stk<double, 1024> dtotal_sorb_eq((size_t)naq * naq);
  for (int ikr = 0; ikr < t.nkinmr(); ++ikr) {
    int irxn = t.mr_rxn[ikr];
    std::fill(total_sorb_eq.begin(), total_sorb_eq.end(), 0.0);
    std::fill(dtotal_sorb_eq.begin(), dtotal_sorb_eq.end(), 0.0);
    RTotalSorbEqSurfCplx1(t, a, irxn, a.free_site_conc[irxn], nullptr, total_sorb_eq.data(), dtotal_sorb_eq.data());
    double *S = &a.kinmr_total_sorb[(size_t)ikr * (t.mr_ld + 1) * naq];
    for (int i = 0; i < naq; ++i) S[i] = total_sorb_eq[i];
  }
}

// ---------------------------------------------------------------- reaction.F90:4969-5006
void RTAuxVarCompute(const Tables &t, AuxVar &a) {
  RTotal(t, a);
  if (t.neqsorb() > 0) RTotalSorb(t, a);
}

// ---------------------------------------------------------------- reaction.F90:5072-5148
void RTAccumulation(const Tables &t, const AuxVar &a, double *Res) {
  double psv_t = a.porosity * a.sat * 1000.0 * a.volume;
  for (int i = 0; i < t.ncomp; ++i) Res[i] = 0.0;
  for (int i = 0; i < t.naq; ++i) Res[i] = psv_t * a.total[i];
  for (int iimob = 0; iimob < t.nim; ++iimob) Res[t.naq + iimob] = Res[t.naq + iimob] + a.immobile[iimob] * a.volume;   // :5117-5125
}
// ---------------------------------------------------------------- reaction.F90:5152-5232
void RTAccumulationDerivative(const Tables &t, const AuxVar &a, double tran_dt, double *J) {
  const int n = t.ncomp, naq = t.naq;
  for (size_t e = 0; e < (size_t)n * n; ++e) J[e] = 0.0;
  double psvd_t = a.porosity * a.sat * 1000.0 * a.volume / tran_dt;
  for (int j = 0; j < naq; ++j)
    for (int i = 0; i < naq; ++i) J[i + (size_t)j * n] = a.dtotal[i + (size_t)j * naq] * psvd_t;
  for (int iimob = 0; iimob < t.nim; ++iimob) J[(naq + iimob) + (size_t)(naq + iimob) * n] = a.volume / tran_dt;        // :5201-5206
}
// ---------------------------------------------------------------- reaction.F90:4539-4568
void RAccumulationSorb(const Tables &t, const AuxVar &a, double *Res) {
  for (int i = 0; i < t.naq; ++i) Res[i] = Res[i] + a.total_sorb_eq[i] * a.volume;
}
// ---------------------------------------------------------------- reaction.F90:4572-4603
void RAccumulationSorbDerivative(const Tables &t, const AuxVar &a, double tran_dt, double *J) {
  const int n = t.ncomp, naq = t.naq;
  double v_t = a.volume / tran_dt;
  for (int j = 0; j < naq; ++j)
    for (int i = 0; i < naq; ++i) J[i + (size_t)j * n] = J[i + (size_t)j * n] + a.dtotal_sorb_eq[i + (size_t)j * naq] * v_t;
}

// ------------------------------------------------- reaction_mineral.F90:564-1000
void RKineticMineral(const Tables &t, AuxVar &a, double *Res, double *Jac, bool compute_derivative) {
  const int naq = t.naq, n = t.ncomp, ncplx_all = t.cplx.n;
  Stk<double, 32> ln_conc(naq), ln_act(naq);
  Stk<double, 512> ln_sec_act(ncplx_all);
  for (int i = 0; i < naq; ++i) { ln_conc[i] = std::log(a.pri_molal[i]); ln_act[i] = ln_conc[i] + std::log(a.pri_act_coef[i]); }
  for (int k = 0; k < ncplx_all; ++k) ln_sec_act[k] = std::log(a.sec_molal[k]) + std::log(a.sec_act_coef[k]);
  for (int im = 0; im < t.kin.n; ++im) a.mnrl_rate[im] = 0.0;
  const int mp = std::max(t.maxpref, 1), mps = std::max(t.maxprefspec, 1);
  for (int imnrl = 0; imnrl < t.kin.n; ++imnrl) {
    double lnQK = -t.kin.logK[imnrl] * LOG_TO_LN;
    if (t.kin.h2oid[imnrl] > 0) lnQK = lnQK + t.kin.h2ost[imnrl] * a.ln_act_h2o;
    const std::vector<int> &id = t.kin.id[imnrl];
    const std::vector<double> &st = t.kin.st[imnrl];
    int ncomp = (int)id.size();
    for (int i = 0; i < ncomp; ++i) lnQK = lnQK + st[i] * ln_act[id[i]];
    double QK;
    if (lnQK <= 6.90776) QK = std::exp(lnQK); else QK = 1.0e3;
    double affinity_factor;
    if (t.has_Temkin) {
      if (t.has_scale) affinity_factor = 1.0 - std::pow(QK, 1.0 / (t.k_scale[imnrl] * t.k_Temkin[imnrl]));
      else affinity_factor = 1.0 - std::pow(QK, 1.0 / t.k_Temkin[imnrl]);
    } else if (t.has_scale) {
      affinity_factor = 1.0 - std::pow(QK, 1.0 / t.k_scale[imnrl]);
    } else {
      affinity_factor = 1.0 - QK;
    }
    double sign_ = std::copysign(1.0, affinity_factor);
    double Im, Im_const, sum_prefactor_rate;
    double prefactor[10];
    double ln_prefactor_spec[10][5];
    if (a.mnrl_volfrac[imnrl] > 0 || sign_ < 0.0) {
      if (t.k_aff_thresh[imnrl] > 0.0) {
        if (sign_ < 0.0 && QK < t.k_aff_thresh[imnrl]) continue;
      }
      if (t.k_rate_lim[imnrl] > 0.0) affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / t.k_rate_lim[imnrl]);
      if (t.k_npref[imnrl] > 0) {
        sum_prefactor_rate = 0.0;
        for (int i = 0; i < 10; ++i) { prefactor[i] = 0.0; for (int j = 0; j < 5; ++j) ln_prefactor_spec[i][j] = 0.0; }
        for (int ipref = 0; ipref < t.k_npref[imnrl]; ++ipref) {
          double ln_prefactor = 0.0;
          size_t pb = (size_t)imnrl * mp + ipref;
          int nps = t.pref_id[pb * (t.maxprefspec + 1)];
          for (int ips = 0; ips < nps; ++ips) {
            int icomp = t.pref_id[pb * (t.maxprefspec + 1) + ips + 1];
            double ln_spec_act = (icomp > 0) ? ln_act[icomp - 1] : ln_sec_act[-icomp - 1];
            double ln_numerator = t.pref_alpha[pb * mps + ips] * ln_spec_act;
            double ln_denominator = std::log(1.0 + std::exp(std::log(t.pref_atten[pb * mps + ips]) + t.pref_beta[pb * mps + ips] * ln_spec_act));
            ln_prefactor = ln_prefactor + ln_numerator;
            ln_prefactor = ln_prefactor - ln_denominator;
            ln_prefactor_spec[ipref][ips] = ln_numerator - ln_denominator;
          }
          prefactor[ipref] = std::exp(ln_prefactor);
          double arrhenius_factor = 1.0;
          if (t.pref_Ea[pb] > 0.0)
            arrhenius_factor = std::exp(t.pref_Ea[pb] / IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (a.temp + 273.15)));
          sum_prefactor_rate = sum_prefactor_rate + prefactor[ipref] * t.pref_rate[pb] * arrhenius_factor;
        }
      } else {
        double arrhenius_factor = 1.0;
        if (t.k_Ea[imnrl] > 0.0)
          arrhenius_factor = std::exp(t.k_Ea[imnrl] / IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (a.temp + 273.15)));
        sum_prefactor_rate = t.k_rate[imnrl] * arrhenius_factor;
      }
      Im_const = -a.mnrl_area[imnrl];
      if (t.has_scale) Im_const = Im_const / t.k_scale[imnrl];
      if (t.has_power) Im = Im_const * sign_ * std::pow(std::fabs(affinity_factor), t.k_power[imnrl]) * sum_prefactor_rate;
      else Im = Im_const * sign_ * std::fabs(affinity_factor) * sum_prefactor_rate;
      a.mnrl_rate[imnrl] = Im;
    } else {
      continue;
    }
    Im_const = Im_const * a.volume;
    Im = Im * a.volume;
    for (int i = 0; i < ncomp; ++i) Res[id[i]] = Res[id[i]] + st[i] * Im;
    if (!compute_derivative) continue;

    double dIm_dQK;
    if (t.has_power) dIm_dQK = -Im * t.k_power[imnrl] / std::fabs(affinity_factor);
    else dIm_dQK = -Im_const * sum_prefactor_rate;
    if (t.has_Temkin) {
      if (t.has_scale) dIm_dQK = dIm_dQK * (1.0 / (t.k_scale[imnrl] * t.k_Temkin[imnrl])) / QK * (1.0 - affinity_factor);
      else dIm_dQK = dIm_dQK * (1.0 / t.k_Temkin[imnrl]) / QK * (1.0 - affinity_factor);
    } else if (t.has_scale) {
      dIm_dQK = dIm_dQK * (1.0 / t.k_scale[imnrl]) / QK * (1.0 - affinity_factor);
    }
    if (t.k_rate_lim[imnrl] <= 0.0) {
      for (int j = 0; j < ncomp; ++j) {
        int jcomp = id[j];
        double dQK_dCj = st[j] * QK * std::exp(-ln_conc[jcomp]);
        double dQK_dmj = dQK_dCj * a.den_kg * 1.0e-3;
        for (int i = 0; i < ncomp; ++i) {
          int icomp = id[i];
          Jac[icomp + (size_t)jcomp * n] = Jac[icomp + (size_t)jcomp * n] + st[i] * dIm_dQK * dQK_dmj;
        }
      }
    } else {
      double den = 1.0 + (1.0 - affinity_factor) / t.k_rate_lim[imnrl];
      for (int j = 0; j < ncomp; ++j) {
        int jcomp = id[j];
        double dQK_dCj = st[j] * QK * std::exp(-ln_conc[jcomp]);
        double dQK_dmj = dQK_dCj * a.den_kg * 1.0e-3;
        for (int i = 0; i < ncomp; ++i) {
          int icomp = id[i];
          Jac[icomp + (size_t)jcomp * n] = Jac[icomp + (size_t)jcomp * n] +
              st[i] * dIm_dQK * (1.0 + QK / t.k_rate_lim[imnrl] / den) * dQK_dmj / den;
        }
      }
    }
    if (t.k_npref[imnrl] > 0) {
      double dIm_dsum_prefactor_rate = Im / sum_prefactor_rate;
      for (int ipref = 0; ipref < t.k_npref[imnrl]; ++ipref) {
        size_t pb = (size_t)imnrl * mp + ipref;
        double arrhenius_factor = 1.0;
        if (t.pref_Ea[pb] > 0.0)
          arrhenius_factor = std::exp(t.pref_Ea[pb] / IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (a.temp + 273.15)));
        double ln_prefactor = std::log(prefactor[ipref]);
        int nps = t.pref_id[pb * (t.maxprefspec + 1)];
        for (int ips = 0; ips < nps; ++ips) {
          double dprefactor_dprefactor_spec = std::exp(ln_prefactor - ln_prefactor_spec[ipref][ips]);
          int icomp = t.pref_id[pb * (t.maxprefspec + 1) + ips + 1];
          double ln_spec_act, spec_act_coef;
          if (icomp > 0) { ln_spec_act = ln_act[icomp - 1]; spec_act_coef = a.pri_act_coef[icomp - 1]; }
          else { ln_spec_act = ln_sec_act[-icomp - 1]; spec_act_coef = a.sec_act_coef[-icomp - 1]; }
          double alpha = t.pref_alpha[pb * mps + ips], beta = t.pref_beta[pb * mps + ips], atten = t.pref_atten[pb * mps + ips];
          double dnum = alpha * std::exp(ln_prefactor_spec[ipref][ips] - ln_spec_act);
          double ln_gam_m_beta = beta * ln_spec_act;
          double denominator = 1.0 + std::exp(std::log(atten) + ln_gam_m_beta);
          double dden = -1.0 * std::exp(ln_prefactor_spec[ipref][ips]) / denominator * atten * beta * std::exp(ln_gam_m_beta - ln_spec_act);
          double dprefactor_spec_dspec = dnum + dden;
          dprefactor_spec_dspec = dprefactor_spec_dspec * spec_act_coef;
          double dIm_dspec = dIm_dsum_prefactor_rate * dprefactor_dprefactor_spec * dprefactor_spec_dspec * t.pref_rate[pb] * arrhenius_factor;
          if (icomp > 0) {
            for (int i = 0; i < ncomp; ++i) {
              int jcomp = id[i];
              Jac[jcomp + (size_t)(icomp - 1) * n] = Jac[jcomp + (size_t)(icomp - 1) * n] + st[i] * dIm_dspec;
            }
          } else {
            // Secondary-species prefactor: the reference clobbers its loop variables ncomp/icomp
            // here (reaction_mineral.F90:977-979), changing later trip counts for this mineral.
            // Tables with secondary prefactor species are rejected by the product; the oracle
            // mirrors the arithmetic of the branch with the complex's own species list.
            int icplx = -icomp - 1;
            double lnQKc = -t.cplx.logK[icplx] * LOG_TO_LN;
            if (t.cplx.h2oid[icplx] > 0) lnQKc = lnQKc + t.cplx.h2ost[icplx] * a.ln_act_h2o;
            const std::vector<int> &cid = t.cplx.id[icplx];
            const std::vector<double> &cst = t.cplx.st[icplx];
            for (size_t i = 0; i < cid.size(); ++i) lnQKc = lnQKc + cst[i] * ln_act[cid[i]];
            for (size_t j = 0; j < cid.size(); ++j) {
              int jcomp = cid[j];
              double tempreal = cst[j] * std::exp(lnQKc - ln_conc[jcomp]) / a.sec_act_coef[icplx];
              for (size_t i = 0; i < cid.size(); ++i)
                Jac[cid[i] + (size_t)jcomp * n] = Jac[cid[i] + (size_t)jcomp * n] + cst[i] * tempreal * dIm_dspec;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------- reaction_surf_complex.F90:566-654
void RMultiRateSorption(const Tables &t, AuxVar &a, double tran_dt, double *Res, double *Jac, bool compute_derivative) {
  const int naq = t.naq, n = t.ncomp;
  Stk<double, 32> total_sorb_eq(naq);
  Stk<double, 1024> dtotal_sorb_eq((size_t)naq * naq);
  const size_t blk = (size_t)(t.mr_ld + 1) * naq;
  for (int ikr = 0; ikr < t.nkinmr(); ++ikr)
    for (int i = 0; i < naq; ++i) a.kinmr_total_sorb[ikr * blk + i] = 0.0;
  for (int ikr = 0; ikr < t.nkinmr(); ++ikr) {
    int irxn = t.mr_rxn[ikr];
    std::fill(total_sorb_eq.begin(), total_sorb_eq.end(), 0.0);
    std::fill(dtotal_sorb_eq.begin(), dtotal_sorb_eq.end(), 0.0);
    RTotalSorbEqSurfCplx1(t, a, irxn, a.free_site_conc[irxn], nullptr, total_sorb_eq.data(), dtotal_sorb_eq.data());
    for (int irate = 0; irate < t.mr_nrate[ikr]; ++irate) {
      double rate = t.mr_rate[(size_t)ikr * t.mr_ld + irate], frac = t.mr_frac[(size_t)ikr * t.mr_ld + irate];
      double kdt = rate * tran_dt;
      double one_plus_kdt = 1.0 + kdt;
      double k_over_one_plus_kdt = rate / one_plus_kdt;
      const double *Sr = &a.kinmr_total_sorb[ikr * blk + (size_t)(irate + 1) * naq];
      for (int i = 0; i < naq; ++i) Res[i] = Res[i] + a.volume * k_over_one_plus_kdt * (frac * total_sorb_eq[i] - Sr[i]);
      if (compute_derivative) {
        double c = a.volume * k_over_one_plus_kdt * frac;
        for (int j = 0; j < naq; ++j)
          for (int i = 0; i < naq; ++i) Jac[i + (size_t)j * n] = Jac[i + (size_t)j * n] + c * dtotal_sorb_eq[i + (size_t)j * naq];
      }
    }
    for (int i = 0; i < naq; ++i) a.kinmr_total_sorb[ikr * blk + i] = total_sorb_eq[i];
  }
}

// ---------------------------------------------------------------- reaction_surf_complex.F90:938-1137
// One kinetic reaction on mineral surface 1, = surface complexation reaction 1 (the only configuration in which the
// reference's indices are in bounds, rxn_pack.h): isite = ikinrxn = 1, global complex id = position in the reaction.
void RKineticSurfCplx(const Tables &t, AuxVar &a, double dt, double *Res, double *Jac, bool compute_derivative) {
  const int n = t.ncomp, naq = t.naq;
  Stk<double, 32> ln_conc(naq), ln_act(naq);
  for (int i = 0; i < naq; ++i) { ln_conc[i] = std::log(a.pri_molal[i]); ln_act[i] = ln_conc[i] + std::log(a.pri_act_coef[i]); }
  const int irxn = t.kin_rxn[0];
  const std::vector<int> &cplx = t.rxn_cplx[irxn];
  const int ncplx = (int)cplx.size();
  Stk<double, 256> lnQ(t.srf.n, 0.0), Q(t.srf.n, 0.0);
  for (int k = 0; k < ncplx; ++k) {
    const int icplx = cplx[k];
    if (t.srf.h2oid[icplx] > 0) lnQ[icplx] = lnQ[icplx] + t.srf.h2ost[icplx] * a.ln_act_h2o;
    for (size_t i = 0; i < t.srf.id[icplx].size(); ++i) lnQ[icplx] = lnQ[icplx] + t.srf.st[icplx][i] * ln_act[t.srf.id[icplx][i]];
    Q[icplx] = std::exp(lnQ[icplx]);
  }
  const double *kf = t.kin_kf.data(), *kb = t.kin_kb.data();       // (icplx, ikinrxn = 1)
  double numerator_sum = 0.0;
  for (int k = 0; k < ncplx; ++k) {
    const int icplx = cplx[k];
    numerator_sum = numerator_sum + a.kinsrfcplx_conc[icplx] / (1.0 + kb[icplx] * dt);
  }
  numerator_sum = t.rxn_site_density[0] - numerator_sum;           // srfcplxrxn_site_density(isite), isite = 1
  double denominator_sum = 1.0;
  for (int k = 0; k < ncplx; ++k) {
    const int icplx = cplx[k];
    denominator_sum = denominator_sum + (kf[icplx] * dt) / (1.0 + kb[icplx] * dt) * Q[icplx];
  }
  std::vector<double> conc_kp1(t.srf.n, 0.0);
  for (int k = 0; k < ncplx; ++k) {
    const int icplx = cplx[k];
    const double conc_k = a.kinsrfcplx_conc[icplx];
    const double denominator = 1.0 + kb[icplx] * dt;
    conc_kp1[icplx] = (conc_k + kf[icplx] * dt * numerator_sum / denominator_sum * Q[icplx]) / denominator;
    a.kinsrfcplx_conc_kp1[icplx] = conc_kp1[icplx];
  }
  a.kinsrfcplx_free_site_conc[0] = numerator_sum / denominator_sum;
  for (int k = 0; k < ncplx; ++k) {
    const int icplx = cplx[k];
    for (size_t i = 0; i < t.srf.id[icplx].size(); ++i) {
      const int icomp = t.srf.id[icplx][i];
      Res[icomp] = Res[icomp] + t.srf.st[icplx][i] * (conc_kp1[icplx] - a.kinsrfcplx_conc[icplx]) / dt * a.volume;
    }
  }
  if (compute_derivative) {
    Stk<double, 32> fac_sum(naq, 0.0);
    for (int k = 0; k < ncplx; ++k) {
      const int icplx = cplx[k];
      const double denominator = 1.0 + kb[icplx] * dt;
      const double fac = kf[icplx] / denominator;
      for (size_t j = 0; j < t.srf.id[icplx].size(); ++j)
        fac_sum[t.srf.id[icplx][j]] = fac_sum[t.srf.id[icplx][j]] + t.srf.st[icplx][j] * fac * Q[icplx];
    }
    for (int k = 0; k < ncplx; ++k) {
      const int icplx = cplx[k];
      const double denominator = 1.0 + kb[icplx] * dt;
      const double fac = kf[icplx] / denominator;
      for (size_t j = 0; j < t.srf.id[icplx].size(); ++j) {
        const int jcomp = t.srf.id[icplx][j];
        for (size_t l = 0; l < t.srf.id[icplx].size(); ++l) {
          const int lcomp = t.srf.id[icplx][l];
          Jac[jcomp + (size_t)lcomp * n] = Jac[jcomp + (size_t)lcomp * n] +
              (t.srf.st[icplx][j] * fac * numerator_sum * Q[icplx] * (t.srf.st[icplx][l] - dt * fac_sum[lcomp] / denominator_sum)) /
                  denominator_sum * std::exp(-ln_conc[lcomp]) * a.volume;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- reaction.F90:4607-4690
void RRadioactiveDecay(const Tables &t, AuxVar &a, double *Res, double *Jac, bool compute_derivative) {
  const int n = t.ncomp, naq = t.naq;
  const double L_water = a.porosity * a.sat * a.volume * 1.0e3;
  for (int irxn = 0; irxn < t.ndecay; ++irxn) {
    int icomp = t.dec_fwd[irxn];
    double sum = a.total[icomp] * L_water;
    sum = sum + a.total_sorb_eq[icomp] * a.volume;              // total_sorb_eq is always associated here (zero without sorption)
    const double rate = sum * t.dec_kf[irxn];
    const int ncomp = (int)t.dec_id[irxn].size();
    for (int i = 0; i < ncomp; ++i) {
      icomp = t.dec_id[irxn][i];
      Res[icomp] = Res[icomp] - t.dec_st[irxn][i] * rate;
    }
    if (!compute_derivative) continue;
    const double tempreal = -1.0 * t.dec_kf[irxn];
    const int jcomp = t.dec_fwd[irxn];
    for (int i = 0; i < ncomp; ++i) {
      icomp = t.dec_id[irxn][i];
      for (int j = 0; j < naq; ++j)
        Jac[icomp + (size_t)j * n] = Jac[icomp + (size_t)j * n] +
            tempreal * t.dec_st[irxn][i] * (a.dtotal[jcomp + (size_t)j * naq] * L_water + a.dtotal_sorb_eq[jcomp + (size_t)j * naq] * a.volume);
    }
  }
}

// ---------------------------------------------------------------- reaction.F90:4694-4831
void RGeneral(const Tables &t, AuxVar &a, double *Res, double *Jac, bool compute_derivative) {
  const int n = t.ncomp, naq = t.naq;
  Stk<double, 32> ln_conc(naq), ln_act(naq);
  for (int i = 0; i < naq; ++i) { ln_conc[i] = std::log(a.pri_molal[i]); ln_act[i] = ln_conc[i] + std::log(a.pri_act_coef[i]); }
  for (int irxn = 0; irxn < t.ngen; ++irxn) {
    const double kf = t.gen_kf[irxn], kr = t.gen_kr[irxn];
    double Qkf, lnQkf = 0.0, Qkr, lnQkr = 0.0;
    if (kf > 0.0) {
      lnQkf = std::log(kf);
      for (size_t i = 0; i < t.genf_id[irxn].size(); ++i) lnQkf = lnQkf + t.genf_st[irxn][i] * ln_act[t.genf_id[irxn][i]];
      Qkf = std::exp(lnQkf);
    } else {
      Qkf = 0.0;
    }
    if (kr > 0.0) {
      lnQkr = std::log(kr);
      for (size_t i = 0; i < t.genb_id[irxn].size(); ++i) lnQkr = lnQkr + t.genb_st[irxn][i] * ln_act[t.genb_id[irxn][i]];
      Qkr = std::exp(lnQkr);
    } else {
      Qkr = 0.0;
    }
    const double por_den_sat_vol = a.porosity * a.den_kg * a.sat * a.volume;
    for (size_t i = 0; i < t.gen_id[irxn].size(); ++i) {
      const int icomp = t.gen_id[irxn][i];
      Res[icomp] = Res[icomp] - t.gen_st[irxn][i] * (Qkf - Qkr) * por_den_sat_vol;
    }
    if (!compute_derivative) continue;
    if (kf > 0.0) {
      for (size_t j = 0; j < t.genf_id[irxn].size(); ++j) {
        const int jcomp = t.genf_id[irxn][j];
        const double tempreal = -1.0 * t.genf_st[irxn][j] * std::exp(lnQkf - ln_conc[jcomp]) * por_den_sat_vol;
        for (size_t i = 0; i < t.gen_id[irxn].size(); ++i) {
          const int icomp = t.gen_id[irxn][i];
          Jac[icomp + (size_t)jcomp * n] = Jac[icomp + (size_t)jcomp * n] + t.gen_st[irxn][i] * tempreal;
        }
      }
    }
    if (kr > 0.0) {
      for (size_t j = 0; j < t.genb_id[irxn].size(); ++j) {
        const int jcomp = t.genb_id[irxn][j];
        const double tempreal = t.genb_st[irxn][j] * std::exp(lnQkr - ln_conc[jcomp]) * por_den_sat_vol;
        for (size_t i = 0; i < t.gen_id[irxn].size(); ++i) {
          const int icomp = t.gen_id[irxn][i];
          Jac[icomp + (size_t)jcomp * n] = Jac[icomp + (size_t)jcomp * n] + t.gen_st[irxn][i] * tempreal;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- reaction_microbial.F90:236-450
void RMicrobial(const Tables &t, AuxVar &a, double *Res, double *Jac, bool compute_derivative) {
  const int n = t.ncomp, naq = t.naq;
  const double PI = 3.14159265359;               // pflotran_constants.F90:55 (truncated on purpose, as LOG_TO_LN)
  double monod[10], inhibition[10];
  for (int irxn = 0; irxn < t.nmic; ++irxn) {
    const int ncomp = (int)t.mic_id[irxn].size();
    double rate_constant = t.mic_k[irxn];
    double Im = rate_constant;
    if (t.mic_has_Ea)
      Im = Im * std::exp(t.mic_Ea[irxn] / IDEAL_GAS_CONSTANT * (1.0 / 298.15 - 1.0 / (a.temp + 273.15)));
    double yield = 0.0, biomass_conc = 0.0;
    const std::vector<int> &mon = t.mic_monod[irxn], &inh = t.mic_inhib[irxn];
    for (size_t ii = 0; ii < mon.size(); ++ii) {
      const int imonod = mon[ii], icomp = t.monod_spec[imonod];
      const double activity = a.pri_molal[icomp] * a.pri_act_coef[icomp];
      monod[ii] = (activity - t.monod_Cth[imonod]) / (t.monod_K[imonod] + activity - t.monod_Cth[imonod]);
      Im = Im * monod[ii];
    }
    for (size_t ii = 0; ii < inh.size(); ++ii) {
      const int iinhibition = inh[ii], icomp = t.inhib_spec[iinhibition];
      const double activity = a.pri_molal[icomp] * a.pri_act_coef[icomp];
      switch (t.inhib_type[iinhibition]) {
        case RXN_INHIBITION_MONOD: inhibition[ii] = t.inhib_C[iinhibition] / (t.inhib_C[iinhibition] + activity); break;
        case RXN_INHIBITION_INVERSE_MONOD: inhibition[ii] = activity / (t.inhib_C[iinhibition] + activity); break;
        case RXN_INHIBITION_THRESHOLD:
          inhibition[ii] = 0.5 + std::atan((activity - t.inhib_C[iinhibition]) * t.inhib_C2[iinhibition]) / PI;
          break;
        default: inhibition[ii] = 0.0;   // the reference's select has no default (a stale value): such tables are rejected at create time
      }
      Im = Im * inhibition[ii];
    }
    const int ibiomass = t.mic_biomass[irxn];
    int immobile_id = -1;
    if (ibiomass >= 0) {
      immobile_id = naq + ibiomass;
      biomass_conc = a.immobile[ibiomass];
      yield = t.mic_yield[irxn];
      Im = Im * biomass_conc;
    }
    const double por_sat_vol = a.porosity * a.sat * a.volume;
    Im = Im * 1.0e3 * por_sat_vol;
    for (int i = 0; i < ncomp; ++i) { const int icomp = t.mic_id[irxn][i]; Res[icomp] = Res[icomp] - t.mic_st[irxn][i] * Im; }
    if (ibiomass >= 0) Res[immobile_id] = Res[immobile_id] - yield * Im;
    if (!compute_derivative) continue;
    for (size_t ii = 0; ii < mon.size(); ++ii) {
      const int imonod = mon[ii], jcomp = t.monod_spec[imonod];
      const double act_coef = a.pri_act_coef[jcomp], activity = a.pri_molal[jcomp] * act_coef;
      const double dR_dX = Im / monod[ii];
      const double denominator = t.monod_K[imonod] + activity - t.monod_Cth[imonod];
      const double dX_dc = act_coef / denominator - act_coef * (activity - t.monod_Cth[imonod]) / (denominator * denominator);
      const double dR_dc = -1.0 * dR_dX * dX_dc;
      for (int i = 0; i < ncomp; ++i) {
        const int icomp = t.mic_id[irxn][i];
        Jac[icomp + (size_t)jcomp * n] = Jac[icomp + (size_t)jcomp * n] + t.mic_st[irxn][i] * dR_dc;
      }
      if (ibiomass >= 0) Jac[immobile_id + (size_t)jcomp * n] = Jac[immobile_id + (size_t)jcomp * n] + yield * dR_dc;
    }
    for (size_t ii = 0; ii < inh.size(); ++ii) {
      const int iinhibition = inh[ii], jcomp = t.inhib_spec[iinhibition];
      const double act_coef = a.pri_act_coef[jcomp], activity = a.pri_molal[jcomp] * act_coef;
      const double dR_dX = Im / inhibition[ii];
      double dX_dc = 0.0;
      switch (t.inhib_type[iinhibition]) {
        case RXN_INHIBITION_MONOD: {
          const double denominator = t.inhib_C[iinhibition] + activity;
          dX_dc = -1.0 * act_coef * t.inhib_C[iinhibition] / (denominator * denominator);
        } break;
        case RXN_INHIBITION_INVERSE_MONOD: {
          const double denominator = t.inhib_C[iinhibition] + activity;
          dX_dc = act_coef / denominator - act_coef * activity / (denominator * denominator);
        } break;
        case RXN_INHIBITION_THRESHOLD: {
          const double tempreal = (activity - t.inhib_C[iinhibition]) * t.inhib_C2[iinhibition];
          dX_dc = (t.inhib_C2[iinhibition] * act_coef / (1.0 + tempreal * tempreal)) / PI;
        } break;
      }
      const double dR_dc = -1.0 * dR_dX * dX_dc;
      for (int i = 0; i < ncomp; ++i) {
        const int icomp = t.mic_id[irxn][i];
        Jac[icomp + (size_t)jcomp * n] = Jac[icomp + (size_t)jcomp * n] + t.mic_st[irxn][i] * dR_dc;
      }
      if (ibiomass >= 0) Jac[immobile_id + (size_t)jcomp * n] = Jac[immobile_id + (size_t)jcomp * n] + yield * dR_dc;
    }
    if (ibiomass >= 0) {
      const double dR_dbiomass = -1.0 * Im / biomass_conc;
      for (int i = 0; i < ncomp; ++i) {
        const int icomp = t.mic_id[irxn][i];
        Jac[icomp + (size_t)immobile_id * n] = Jac[icomp + (size_t)immobile_id * n] + t.mic_st[irxn][i] * dR_dbiomass;
      }
      Jac[immobile_id + (size_t)immobile_id * n] = Jac[immobile_id + (size_t)immobile_id * n] + yield * dR_dbiomass;
    }
  }
}

// ---------------------------------------------------------------- reaction_immobile.F90:240-293
void RImmobileDecay(const Tables &t, AuxVar &a, double *Res, double *Jac, bool compute_derivative) {
  const int n = t.ncomp;
  const double volume = a.volume;
  for (int irxn = 0; irxn < t.nimdecay; ++irxn) {
    const int icomp = t.imdec_id[irxn];
    const double rate_constant = t.imdec_k[irxn] * volume;
    const double rate = rate_constant * a.immobile[icomp];
    const int immobile_id = t.naq + icomp;
    Res[immobile_id] = Res[immobile_id] + rate;
    if (!compute_derivative) continue;
    Jac[immobile_id + (size_t)immobile_id * n] = Jac[immobile_id + (size_t)immobile_id * n] + rate_constant;
  }
}

// ---------------------------------------------------------------- reaction.F90:3515-3584
void RReaction(const Tables &t, AuxVar &a, double tran_dt, double *Res, double *Jac, bool derivative) {
  if (t.kin.n > 0) RKineticMineral(t, a, Res, Jac, derivative);
  if (t.nkinmr() > 0) RMultiRateSorption(t, a, tran_dt, Res, Jac, derivative);
  if (t.nkinrxn > 0) RKineticSurfCplx(t, a, tran_dt, Res, Jac, derivative);
  if (t.ndecay > 0) RRadioactiveDecay(t, a, Res, Jac, derivative);
  if (t.ngen > 0) RGeneral(t, a, Res, Jac, derivative);
  if (t.nmic > 0) RMicrobial(t, a, Res, Jac, derivative);
  if (t.nimdecay > 0) RImmobileDecay(t, a, Res, Jac, derivative);
}

// ---------------------------------------------------------------- reaction.F90:3322-3511
int RReact(Tables &t, AuxVar &a, double *tran_xx, double tran_dt, int dt_mode, int maxit, int *exit_reason) {
  const int n = t.ncomp, naq = t.naq;
  Stk<double, 32> residual(n), prev_solution(n), new_solution(n), update(n), fixed_accum(n);
  Stk<double, 1024> J((size_t)n * n);
  int num_iterations = 0;
  *exit_reason = 0;
  for (int i = 0; i < naq; ++i) a.total[i] = tran_xx[i];
  for (int i = 0; i < t.nim; ++i) a.immobile[i] = tran_xx[naq + i];                        // :3386-3392
  RUpdateTempDependentCoefs(t, a);
  RTAccumulation(t, a, fixed_accum.data());
  if (t.neqsorb() > 0) RAccumulationSorb(t, a, fixed_accum.data());
  if (t.act_freq != RXN_ACT_COEF_FREQUENCY_OFF) RActivityCoefficients(t, a);
  for (;;) {
    num_iterations = num_iterations + 1;
    if (t.act_freq == RXN_ACT_COEF_FREQUENCY_NEWTON_ITER) RActivityCoefficients(t, a);
    RTAuxVarCompute(t, a);
    RTAccumulation(t, a, residual.data());
    for (int i = 0; i < n; ++i) residual[i] = residual[i] - fixed_accum[i];
    RTAccumulationDerivative(t, a, tran_dt, J.data());
    if (t.neqsorb() > 0) {
      RAccumulationSorb(t, a, residual.data());
      RAccumulationSorbDerivative(t, a, tran_dt, J.data());
    }
    if (dt_mode == RXN_DT_CONSISTENT)
      for (int i = 0; i < n; ++i) residual[i] = residual[i] / tran_dt;
    RReaction(t, a, tran_dt, residual.data(), J.data(), true);
    double mx = 0.0;
    bool nonfinite = false;
    for (int i = 0; i < n; ++i) { mx = std::max(mx, std::fabs(residual[i])); if (!std::isfinite(residual[i])) nonfinite = true; }
    if (nonfinite) { a.flags |= RXN_FLAG_NONFINITE; break; }
    if (mx < t.res_tol) { *exit_reason = RXN_EXIT_RESIDUAL; break; }
    for (int i = 0; i < naq; ++i) prev_solution[i] = a.pri_molal[i];
    for (int i = 0; i < t.nim; ++i) prev_solution[naq + i] = a.immobile[i];                // :3448-3452
    // DEVIATION (immobile species + LOG_FORMULATION only): the reference passes rt_auxvar%pri_molal (naqcomp values) as RSolve's
    // conc(ncomp) (:3445), so the immobile columns are scaled by whatever lies behind that array; the immobile concentration -
    // what the log formulation means - is used here (prev_solution is [pri_molal, immobile]).
    if (RSolve(residual.data(), J.data(), prev_solution.data(), update.data(), n, t.use_log)) { a.flags |= RXN_FLAG_LU_ZERO_ROW; break; }
    if (t.use_log) {
      for (int i = 0; i < n; ++i) update[i] = std::copysign(1.0, update[i]) * std::min(std::fabs(update[i]), t.max_dlnC);
      for (int i = 0; i < n; ++i) new_solution[i] = prev_solution[i] * std::exp(-update[i]);
    } else {
      double min_ratio = 1.0e20;
      for (int i = 0; i < n; ++i) {
        if (prev_solution[i] <= update[i]) {
          double ratio = std::fabs(prev_solution[i] / update[i]);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      if (min_ratio < 1.0) for (int i = 0; i < n; ++i) update[i] = update[i] * min_ratio * 0.99;
      for (int i = 0; i < n; ++i) new_solution[i] = prev_solution[i] - update[i];
    }
    double maximum_relative_change = 0.0;
    for (int i = 0; i < n; ++i) {
      double r = std::fabs((new_solution[i] - prev_solution[i]) / prev_solution[i]);
      if (!(r <= maximum_relative_change)) maximum_relative_change = r;  // NaN propagates like maxval never would; flagged below
    }
    if (!std::isfinite(maximum_relative_change)) { a.flags |= RXN_FLAG_NONFINITE; break; }
    if (maximum_relative_change < t.rel_tol) { *exit_reason = RXN_EXIT_REL_CHANGE; break; }
    if (num_iterations > 50) {
      double scale = 0.1;  // the >100/>150/>500 branches are unreachable (reaction.F90:3480-3491)
      for (int i = 0; i < n; ++i) new_solution[i] = scale * (new_solution[i] - prev_solution[i]) + prev_solution[i];
    }
    for (int i = 0; i < naq; ++i) a.pri_molal[i] = new_solution[i];
    for (int i = 0; i < t.nim; ++i) a.immobile[i] = new_solution[naq + i];                 // :3499-3502
    if (num_iterations >= maxit) { a.flags |= RXN_FLAG_CAPPED; break; }  // oracle/GPU-only guard
  }
  RTAuxVarCompute(t, a);
  for (int i = 0; i < naq; ++i) tran_xx[i] = a.pri_molal[i];  // reactive_transport.F90:1711
  // :1712-1716 writes the immobile values at tran_xx_p(offset_immobile : ...) without the cell's offset (a reference bug,
  // SURVEY 8a quirks); the cell's own slots are written here
  for (int i = 0; i < t.nim; ++i) tran_xx[naq + i] = a.immobile[i];
  return num_iterations;
}

// ---------------------------------------------------------------- reaction.F90:5320-5429
void RUpdateKineticState(const Tables &t, AuxVar &a, double tran_dt) {
  const int n = t.ncomp, naq = t.naq;
  if (t.kin.n > 0) {
    Stk<double, 32> res(n, 0.0);
    Stk<double, 1024> jac((size_t)n * n, 0.0);
    RKineticMineral(t, a, res.data(), jac.data(), false);
    for (int im = 0; im < t.kin.n; ++im) {
      double delta_volfrac = a.mnrl_rate[im] * t.k_molar_vol[im] * tran_dt;
      a.mnrl_volfrac[im] = a.mnrl_volfrac[im] + delta_volfrac;
      if (a.mnrl_volfrac[im] < 0.0) a.mnrl_volfrac[im] = 0.0;
    }
  }
  const size_t blk = (size_t)(t.mr_ld + 1) * naq;
  for (int ikr = 0; ikr < t.nkinmr(); ++ikr) {
    for (int irate = 0; irate < t.mr_nrate[ikr]; ++irate) {
      double rate = t.mr_rate[(size_t)ikr * t.mr_ld + irate], frac = t.mr_frac[(size_t)ikr * t.mr_ld + irate];
      double kdt = rate * tran_dt;
      double one_plus_kdt = 1.0 + kdt;
      double *Sr = &a.kinmr_total_sorb[ikr * blk + (size_t)(irate + 1) * naq];
      const double *S0 = &a.kinmr_total_sorb[ikr * blk];
      for (int i = 0; i < naq; ++i) Sr[i] = (Sr[i] + kdt * frac * S0[i]) / one_plus_kdt;
    }
  }
  if (t.nkinrxn > 0) {                                          // :5411-5419
    for (int icplx : t.rxn_cplx[t.kin_rxn[0]]) a.kinsrfcplx_conc[icplx] = a.kinsrfcplx_conc_kp1[icplx];
  }
}

// ---------------------------------------------------------------- reaction.F90:1308-2046
int ReactionEquilibrateConstraint(Tables &t, AuxVar &a, const int *constraint_type, const double *conc_in,
                                  const int *constraint_id, const double *free_ion_guess,
                                  int use_prev_soln_as_guess, int initialize_with_molality,
                                  double *basis_molarity, int *num_iterations_out) {
  const int naq = t.naq;
  std::vector<double> conc(conc_in, conc_in + naq), Res(naq), update(naq), total_conc(naq, 0.0), free_conc(naq),
      Jac((size_t)naq * naq), prev_molal(naq);
  double convert_molal_to_molar, convert_molar_to_molal;
  const double xmass = 1.0;
  if (initialize_with_molality) { convert_molal_to_molar = a.den_kg * xmass / 1000.0; convert_molar_to_molal = 1.0; }
  else { convert_molal_to_molar = 1.0; convert_molar_to_molal = 1000.0 / a.den_kg / xmass; }
  RUpdateTempDependentCoefs(t, a);
  // NB: mineral logK for constraint minerals is only temperature-updated for kinetic minerals here.
  if (use_prev_soln_as_guess) free_conc = a.pri_molal;
  else if (free_ion_guess) free_conc.assign(free_ion_guess, free_ion_guess + naq);
  else free_conc.assign(naq, 1.0e-9);
  for (int i = 0; i < naq; ++i) {
    switch (constraint_type[i]) {
      case CONSTRAINT_NULL: case CONSTRAINT_TOTAL: total_conc[i] = conc[i] * convert_molal_to_molar; break;
      case CONSTRAINT_TOTAL_SORB: total_conc[i] = conc[i]; break;
      case CONSTRAINT_FREE: free_conc[i] = conc[i] * convert_molar_to_molal; break;
      case CONSTRAINT_LOG: free_conc[i] = std::pow(10.0, conc[i]) * convert_molar_to_molal; break;
      case CONSTRAINT_CHARGE_BAL: if (!use_prev_soln_as_guess) free_conc[i] = conc[i] * convert_molar_to_molal; break;
      case CONSTRAINT_PH:
        if (t.h_ion_id == 0) return -2;
        free_conc[i] = std::pow(10.0, -conc[i]);
        break;
      case CONSTRAINT_MINERAL: if (!use_prev_soln_as_guess) free_conc[i] = conc[i] * convert_molar_to_molal; break;
      case CONSTRAINT_GAS: if (conc[i] <= 0.0) conc[i] = std::pow(10.0, conc[i]); break;
      default: return -3;
    }
  }
  a.pri_molal = free_conc;
  int num_iterations = 0, num_it_act_coef_turned_on = 0;
  bool compute_activity_coefs = use_prev_soln_as_guess != 0;
  bool charge_balance_warning_flag = false;
  for (;;) {
    for (int i = 0; i < naq; ++i)
      if (constraint_type[i] == CONSTRAINT_FREE || constraint_type[i] == CONSTRAINT_LOG) a.pri_molal[i] = free_conc[i];
    if (t.act_freq != RXN_ACT_COEF_FREQUENCY_OFF && compute_activity_coefs) RActivityCoefficients(t, a);
    RTotal(t, a);
    if (t.neqsorb() + t.nkinmr() > 0) {
      if (t.neqsorb() > 0) RTotalSorb(t, a);
      if (t.nkinmr() > 0) RTotalSorbMultiRateAsEQ(t, a);
    }
    std::fill(Jac.begin(), Jac.end(), 0.0);
#define JAC(i, j) Jac[(i) + (size_t)(j) * naq]
    for (int icomp = 0; icomp < naq; ++icomp) {
      switch (constraint_type[icomp]) {
        case CONSTRAINT_NULL: case CONSTRAINT_TOTAL:
          Res[icomp] = a.total[icomp] - total_conc[icomp];
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = a.dtotal[icomp + (size_t)j * naq];
          break;
        case CONSTRAINT_TOTAL_SORB:
          Res[icomp] = a.total_sorb_eq[icomp] - total_conc[icomp];
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = a.dtotal_sorb_eq[icomp + (size_t)j * naq];
          break;
        case CONSTRAINT_FREE: case CONSTRAINT_LOG:
          Res[icomp] = 0.0;
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = 0.0;
          JAC(icomp, icomp) = 1.0;
          break;
        case CONSTRAINT_CHARGE_BAL:
          Res[icomp] = 0.0;
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = 0.0;
          for (int jcomp = 0; jcomp < naq; ++jcomp) {
            Res[icomp] = Res[icomp] + t.Z[jcomp] * a.total[jcomp];
            for (int kcomp = 0; kcomp < naq; ++kcomp)
              JAC(icomp, jcomp) = JAC(icomp, jcomp) + t.Z[kcomp] * a.dtotal[kcomp + (size_t)jcomp * naq];
          }
          if (a.pri_molal[icomp] < 1.0e-20 && !charge_balance_warning_flag) {
            if ((Res[icomp] > 0.0 && t.Z[icomp] > 0.0) || (Res[icomp] < 0.0 && t.Z[icomp] < 0.0)) {
              charge_balance_warning_flag = true;
              a.pri_molal[icomp] = (double)1.e-3f;  // reference literal is single precision `1.e-3`
            }
          }
          break;
        case CONSTRAINT_PH:
          Res[icomp] = 0.0;
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = 0.0;
          if (t.h_ion_id > 0) {
            a.pri_molal[icomp] = std::pow(10.0, -conc[icomp]) / a.pri_act_coef[icomp];
            JAC(icomp, icomp) = 1.0;
          } else {
            int icplx = std::abs(t.h_ion_id) - 1;
            double lnQK = -t.cplx.logK[icplx] * LOG_TO_LN;
            if (t.cplx.h2oid[icplx] > 0) lnQK = lnQK + t.cplx.h2ost[icplx] * a.ln_act_h2o;
            for (size_t j = 0; j < t.cplx.id[icplx].size(); ++j) {
              int c = t.cplx.id[icplx][j];
              lnQK = lnQK + t.cplx.st[icplx][j] * std::log(a.pri_molal[c] * a.pri_act_coef[c]);
            }
            lnQK = lnQK + conc[icomp] * LOG_TO_LN;
            double QK = std::exp(lnQK);
            Res[icomp] = 1.0 - QK;
            for (size_t j = 0; j < t.cplx.id[icplx].size(); ++j) {
              int c = t.cplx.id[icplx][j];
              JAC(icomp, c) = -QK / a.pri_molal[c] * t.cplx.st[icplx][j];
            }
          }
          break;
        case CONSTRAINT_MINERAL: {
          int imnrl = constraint_id[icomp] - 1;
          double lnQK = -t.mnrl.logK[imnrl] * LOG_TO_LN;
          if (t.mnrl.h2oid[imnrl] > 0) lnQK = lnQK + t.mnrl.h2ost[imnrl] * a.ln_act_h2o;
          for (size_t j = 0; j < t.mnrl.id[imnrl].size(); ++j) {
            int c = t.mnrl.id[imnrl][j];
            lnQK = lnQK + t.mnrl.st[imnrl][j] * std::log(a.pri_molal[c] * a.pri_act_coef[c]);
          }
          Res[icomp] = lnQK;
          for (size_t j = 0; j < t.mnrl.id[imnrl].size(); ++j) {
            int c = t.mnrl.id[imnrl][j];
            JAC(icomp, c) = t.mnrl.st[imnrl][j] / a.pri_molal[c];
          }
        } break;
        case CONSTRAINT_GAS: {
          int igas = constraint_id[icomp] - 1;
          double lnQK = -t.gas.logK[igas] * LOG_TO_LN;
          if (t.gas.h2oid[igas] > 0) lnQK = lnQK + t.gas.h2ost[igas] * a.ln_act_h2o;
          for (size_t j = 0; j < t.gas.id[igas].size(); ++j) {
            int c = t.gas.id[igas][j];
            lnQK = lnQK + t.gas.st[igas][j] * std::log(a.pri_molal[c] * a.pri_act_coef[c]);
          }
          Res[icomp] = lnQK - std::log(conc[icomp]);
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = 0.0;
          for (size_t j = 0; j < t.gas.id[igas].size(); ++j) {
            int c = t.gas.id[igas][j];
            JAC(icomp, c) = t.gas.st[igas][j] / a.pri_molal[c];
          }
        } break;
      }
    }
#undef JAC
    double maximum_residual = 0.0;
    for (int i = 0; i < naq; ++i) maximum_residual = std::max(maximum_residual, std::fabs(Res[i]));
    bool use_log_formulation;
    if (t.use_log) {
      if (num_iterations > 3 && num_iterations < 9) use_log_formulation = (num_iterations % 2 == 0);
      else use_log_formulation = true;
    } else {
      use_log_formulation = false;
    }
    if (RSolve(Res.data(), Jac.data(), a.pri_molal.data(), update.data(), naq, use_log_formulation)) return -4;
    prev_molal = a.pri_molal;
    if (use_log_formulation) {
      for (int i = 0; i < naq; ++i) update[i] = std::copysign(1.0, update[i]) * std::min(std::fabs(update[i]), t.max_dlnC);
      for (int i = 0; i < naq; ++i) a.pri_molal[i] = a.pri_molal[i] * std::exp(-update[i]);
    } else {
      double min_ratio = 1.0e20;
      for (int i = 0; i < naq; ++i) {
        if (prev_molal[i] <= update[i]) {
          double ratio = std::fabs(prev_molal[i] / update[i]);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      if (min_ratio <= 1.0) for (int i = 0; i < naq; ++i) update[i] = update[i] * min_ratio * 0.99;
      for (int i = 0; i < naq; ++i) a.pri_molal[i] = prev_molal[i] - update[i];
    }
    double mn = a.pri_molal[0];
    for (int i = 1; i < naq; ++i) mn = std::min(mn, a.pri_molal[i]);
    if (!(mn > 0.0)) return -5;  // "Zero concentrations found in constraint"
    double maximum_relative_change = 0.0;
    for (int i = 0; i < naq; ++i) maximum_relative_change = std::max(maximum_relative_change, std::fabs((a.pri_molal[i] - prev_molal[i]) / prev_molal[i]));
    num_iterations = num_iterations + 1;
    if (num_iterations >= 10000) return -6;
    if (maximum_residual < t.res_tol && maximum_relative_change < t.rel_tol) {
      if (compute_activity_coefs && num_iterations - num_it_act_coef_turned_on > 1) break;
      if (!compute_activity_coefs) num_it_act_coef_turned_on = num_iterations;
      compute_activity_coefs = true;
    }
  }
  if (t.neqsorb() + t.nkinmr() > 0) {
    if (t.neqsorb() > 0) RTotalSorb(t, a);
    if (t.nkinmr() > 0) RTotalSorbMultiRateAsEQ(t, a);
  }
  const size_t blk = (size_t)(t.mr_ld + 1) * naq;
  for (int ikr = 0; ikr < t.nkinmr(); ++ikr)
    for (int irate = 0; irate < t.mr_nrate[ikr]; ++irate) {
      double frac = t.mr_frac[(size_t)ikr * t.mr_ld + irate];
      for (int i = 0; i < naq; ++i)
        a.kinmr_total_sorb[ikr * blk + (size_t)(irate + 1) * naq + i] = frac * a.kinmr_total_sorb[ikr * blk + i];
    }
  if (basis_molarity) for (int i = 0; i < naq; ++i) basis_molarity[i] = a.pri_molal[i] * a.den_kg / 1000.0;
  *num_iterations_out = num_iterations;
  return 0;
}

// ------------------------------------------------------------------ SoA <-> per-cell AuxVar
struct View { int64_t ncells, ld; double *f[RXN_F_COUNT]; };

inline double *fp(const View &v, int field, int64_t row) { return v.f[field] ? v.f[field] + row * v.ld : nullptr; }

void gather(const Tables &t, const View &v, int64_t c, AuxVar &a) {
  const int naq = t.naq;
  auto g = [&](int field, std::vector<double> &dst) {
    if (!v.f[field]) return;
    for (size_t r = 0; r < dst.size(); ++r) dst[r] = v.f[field][r * v.ld + c];
  };
  auto g1 = [&](int field, double &dst) { if (v.f[field]) dst = v.f[field][c]; };
  g(RXN_F_PRI_MOLAL, a.pri_molal); g(RXN_F_TOTAL, a.total); g(RXN_F_SEC_MOLAL, a.sec_molal);
  g(RXN_F_PRI_ACT_COEF, a.pri_act_coef); g(RXN_F_SEC_ACT_COEF, a.sec_act_coef); g1(RXN_F_LN_ACT_H2O, a.ln_act_h2o);
  g(RXN_F_TOTAL_SORB_EQ, a.total_sorb_eq); g(RXN_F_FREE_SITE_CONC, a.free_site_conc); g(RXN_F_EQSRFCPLX_CONC, a.eqsrfcplx_conc);
  g(RXN_F_KINMR_TOTAL_SORB, a.kinmr_total_sorb); g(RXN_F_EQIONX_REF_CATION_SORBED_CONC, a.ionx_ref_sorbed);
  if (t.nionx > 0) g(RXN_F_EQIONX_CONC, a.ionx_conc);
  g(RXN_F_MNRL_VOLFRAC, a.mnrl_volfrac); g(RXN_F_MNRL_AREA, a.mnrl_area); g(RXN_F_MNRL_RATE, a.mnrl_rate);
  g1(RXN_F_DEN_KG, a.den_kg); g1(RXN_F_SAT, a.sat); g1(RXN_F_TEMP, a.temp); g1(RXN_F_PRES, a.pres);
  g1(RXN_F_VOLUME, a.volume); g1(RXN_F_POROSITY, a.porosity); g1(RXN_F_SOIL_PARTICLE_DENSITY, a.soil_density);
  g(RXN_F_DTOTAL, a.dtotal); g(RXN_F_DTOTAL_SORB_EQ, a.dtotal_sorb_eq);
  g(RXN_F_KINSRFCPLX_CONC, a.kinsrfcplx_conc); g(RXN_F_KINSRFCPLX_CONC_KP1, a.kinsrfcplx_conc_kp1);
  g(RXN_F_KINSRFCPLX_FREE_SITE_CONC, a.kinsrfcplx_free_site_conc);
  g(RXN_F_IMMOBILE, a.immobile);
  (void)naq;
  a.flags = 0;
}

void scatter(const Tables &t, const View &v, int64_t c, const AuxVar &a) {
  auto s = [&](int field, const std::vector<double> &src) {
    if (!v.f[field]) return;
    for (size_t r = 0; r < src.size(); ++r) v.f[field][r * v.ld + c] = src[r];
  };
  s(RXN_F_PRI_MOLAL, a.pri_molal); s(RXN_F_TOTAL, a.total); s(RXN_F_SEC_MOLAL, a.sec_molal);
  s(RXN_F_PRI_ACT_COEF, a.pri_act_coef); s(RXN_F_SEC_ACT_COEF, a.sec_act_coef);
  if (v.f[RXN_F_LN_ACT_H2O]) v.f[RXN_F_LN_ACT_H2O][c] = a.ln_act_h2o;
  s(RXN_F_TOTAL_SORB_EQ, a.total_sorb_eq); s(RXN_F_FREE_SITE_CONC, a.free_site_conc); s(RXN_F_EQSRFCPLX_CONC, a.eqsrfcplx_conc);
  s(RXN_F_KINMR_TOTAL_SORB, a.kinmr_total_sorb); s(RXN_F_EQIONX_REF_CATION_SORBED_CONC, a.ionx_ref_sorbed);
  if (t.nionx > 0) s(RXN_F_EQIONX_CONC, a.ionx_conc);
  s(RXN_F_MNRL_VOLFRAC, a.mnrl_volfrac); s(RXN_F_MNRL_RATE, a.mnrl_rate);
  s(RXN_F_DTOTAL, a.dtotal); s(RXN_F_DTOTAL_SORB_EQ, a.dtotal_sorb_eq);
  s(RXN_F_KINSRFCPLX_CONC, a.kinsrfcplx_conc); s(RXN_F_KINSRFCPLX_CONC_KP1, a.kinsrfcplx_conc_kp1);
  s(RXN_F_KINSRFCPLX_FREE_SITE_CONC, a.kinsrfcplx_free_site_conc);
  s(RXN_F_IMMOBILE, a.immobile);
}

template <class F> void parallel_cells(int64_t n, int nthreads, F body) {
  if (nthreads <= 1 || n < 2) { body(0, n, 0); return; }
  std::vector<std::thread> th;
  int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int i = 0; i < nthreads; ++i) {
    int64_t b = i * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([=]() { body(b, e, i); });
  }
  for (auto &x : th) x.join();
}

}  // namespace

// ============================================================================ C API (ctypes)
extern "C" {

void *orc_create(const RxnTablesDesc *d) { return load_tables(d); }
void orc_destroy(void *h) { delete (Tables *)h; }

// replaces RTReact loop for the checker: tran_xx AoS [ncells][ncomp]
int orc_react_batch(void *h, const View *v, double *tran_xx, const uint8_t *active, double dt, int dt_mode,
                    int32_t *iters, int32_t *flags, int maxit, int nthreads) {
  const Tables &t0 = *(Tables *)h;
  parallel_cells(v->ncells, nthreads, [&](int64_t b, int64_t e, int) {
    Tables t = t0;  // per-thread copy: RUpdateTempDependentCoefs mutates logK
    AuxVar a; init_auxvar(t, a);
    for (int64_t c = b; c < e; ++c) {
      if (active && !active[c]) { if (iters) iters[c] = 0; if (flags) flags[c] = RXN_FLAG_INACTIVE; continue; }
      gather(t, *v, c, a);
      int reason = 0;
      int it = RReact(t, a, tran_xx + c * t.ncomp, dt, dt_mode, maxit, &reason);
      scatter(t, *v, c, a);
      if (iters) iters[c] = it;
      if (flags) flags[c] = reason | a.flags;
    }
  });
  return 0;
}

// RTUpdateAuxVars cells part (reactive_transport.F90:3790-3846)
int orc_update_auxvars_batch(void *h, const View *v, const double *xx_loc, const uint8_t *active, int update_act_coefs, int nthreads) {
  const Tables &t0 = *(Tables *)h;
  parallel_cells(v->ncells, nthreads, [&](int64_t b, int64_t e, int) {
    Tables t = t0;
    AuxVar a; init_auxvar(t, a);
    for (int64_t c = b; c < e; ++c) {
      if (active && !active[c]) continue;
      gather(t, *v, c, a);
      if (xx_loc) for (int i = 0; i < t.naq; ++i) a.pri_molal[i] = xx_loc[c * t.ncomp + i];
      if (xx_loc) for (int i = 0; i < t.nim; ++i) a.immobile[i] = xx_loc[c * t.ncomp + t.naq + i];   // :3801-3805
      RUpdateTempDependentCoefs(t, a);
      if (update_act_coefs) RActivityCoefficients(t, a);
      RTAuxVarCompute(t, a);
      scatter(t, *v, c, a);
    }
  });
  return 0;
}

// RTUpdateFixedAccumulation (reactive_transport.F90:786-843)
int orc_fixed_accum_batch(void *h, const View *v, const double *xx, const uint8_t *active, double *accum_out, int nthreads) {
  const Tables &t0 = *(Tables *)h;
  parallel_cells(v->ncells, nthreads, [&](int64_t b, int64_t e, int) {
    Tables t = t0;
    AuxVar a; init_auxvar(t, a);
    for (int64_t c = b; c < e; ++c) {
      if (active && !active[c]) continue;
      gather(t, *v, c, a);
      if (xx) for (int i = 0; i < t.naq; ++i) a.pri_molal[i] = xx[c * t.ncomp + i];
      if (xx) for (int i = 0; i < t.nim; ++i) a.immobile[i] = xx[c * t.ncomp + t.naq + i];           // :809-813
      RUpdateTempDependentCoefs(t, a);
      RTAuxVarCompute(t, a);
      RTAccumulation(t, a, accum_out + c * t.ncomp);
      if (t.neqsorb() > 0) RAccumulationSorb(t, a, accum_out + c * t.ncomp);
      scatter(t, *v, c, a);
    }
  });
  return 0;
}

// RTResidualNonFlux accumulation + reaction loops (reactive_transport.F90:2545-2586, 2735-2758)
// and RTJacobianNonFlux (3342-3389, 3445-3465); state must be current (update_auxvars first).
int orc_residual_jacobian_batch(void *h, const View *v, const uint8_t *active, double dt, double *res_out, double *jac_out, int nthreads) {
  const Tables &t0 = *(Tables *)h;
  parallel_cells(v->ncells, nthreads, [&](int64_t b, int64_t e, int) {
    Tables t = t0;
    AuxVar a; init_auxvar(t, a);
    const int n = t.ncomp;
    std::vector<double> Res(n), Res2(n), J((size_t)n * n), J2((size_t)n * n);
    for (int64_t c = b; c < e; ++c) {
      if (active && !active[c]) continue;
      gather(t, *v, c, a);
      RUpdateTempDependentCoefs(t, a);
      RTAccumulation(t, a, Res.data());
      if (t.neqsorb() > 0) RAccumulationSorb(t, a, Res.data());
      for (int i = 0; i < n; ++i) Res[i] = Res[i] / dt;
      std::fill(Res2.begin(), Res2.end(), 0.0); std::fill(J2.begin(), J2.end(), 0.0);
      RTAccumulationDerivative(t, a, dt, J.data());
      if (t.neqsorb() > 0) RAccumulationSorbDerivative(t, a, dt, J.data());
      RReaction(t, a, dt, Res2.data(), J2.data(), true);
      if (res_out) for (int i = 0; i < n; ++i) res_out[c * n + i] = Res[i] + Res2[i];
      if (jac_out) for (size_t k = 0; k < (size_t)n * n; ++k) jac_out[c * (size_t)n * n + k] = J[k] + J2[k];
      scatter(t, *v, c, a);
    }
  });
  return 0;
}

// RTUpdateKineticState loop (reactive_transport.F90:692-705)
int orc_update_kinetic_state_batch(void *h, const View *v, const uint8_t *active, double dt, int nthreads) {
  const Tables &t0 = *(Tables *)h;
  parallel_cells(v->ncells, nthreads, [&](int64_t b, int64_t e, int) {
    Tables t = t0;
    AuxVar a; init_auxvar(t, a);
    for (int64_t c = b; c < e; ++c) {
      if (active && !active[c]) continue;
      gather(t, *v, c, a);
      RUpdateTempDependentCoefs(t, a);
      RUpdateKineticState(t, a, dt);
      scatter(t, *v, c, a);
    }
  });
  return 0;
}

int orc_activity_coefficients_batch(void *h, const View *v, int nthreads) {
  const Tables &t0 = *(Tables *)h;
  parallel_cells(v->ncells, nthreads, [&](int64_t b, int64_t e, int) {
    Tables t = t0; AuxVar a; init_auxvar(t, a);
    for (int64_t c = b; c < e; ++c) { gather(t, *v, c, a); RActivityCoefficients(t, a); scatter(t, *v, c, a); }
  });
  return 0;
}

// ReactionEquilibrateConstraint on cell `cell` of the view
int orc_equilibrate_constraint(void *h, const View *v, int64_t cell, const int32_t *ctype, const double *conc,
                               const int32_t *cid, const double *free_ion_guess, int use_prev_soln_as_guess,
                               int initialize_with_molality, double *basis_molarity, int32_t *num_iterations) {
  Tables t = *(Tables *)h;
  AuxVar a; init_auxvar(t, a);
  gather(t, *v, cell, a);
  std::vector<int> ct(ctype, ctype + t.naq), ci(cid, cid + t.naq);
  int nit = 0;
  int rc = ReactionEquilibrateConstraint(t, a, ct.data(), conc, ci.data(), free_ion_guess, use_prev_soln_as_guess,
                                         initialize_with_molality, basis_molarity, &nit);
  if (num_iterations) *num_iterations = nit;
  scatter(t, *v, cell, a);
  return rc;
}

// dense solve exposed for unit tests (RSolve + ludcmp/lubksb)
int orc_rsolve(double *Res, double *Jac, const double *conc, double *update, int n, int use_log) {
  return RSolve(Res, Jac, conc, update, n, use_log != 0);
}

// ---------------------------------------------------------------- flux side
// The reference has no unit test of TFlux/TFluxDerivative (they are exercised only through whole transport
// regressions): this part of the oracle is UNPINNED by gold files; tests pin it with hand-computed connections
// and conservation.
//
// TFluxCoef, transport.F90:756-819 (liquid phase; ngas = 0).  disp: harmonic_dispersion_over_dist(:,1) per
// connection [nconn][naq]; velocity: internal_velocities(1,conn); fraction_upwind: dist(-1,conn).
// Out: T_up / T_dn [nconn][naq] in L water / s.
int orc_flux_coefs(int naq, int64_t nconn, const double *area, const double *velocity, const double *disp,
                   const double *fraction_upwind, int use_upwinding, double *T_up, double *T_dn) {
  for (int64_t c = 0; c < nconn; ++c) {
    const double q = velocity[c];
    for (int i = 0; i < naq; ++i) {
      const double hd = disp[c * naq + i];
      double cu, cd;
      if (use_upwinding) {
        if (q > 0.0) { cu = hd + q; cd = -hd; }           // :794-796
        else { cu = hd; cd = -hd + q; }                  // :797-799
      } else {
        cu = hd + (1.0 - fraction_upwind[c]) * q;        // :804-807
        cd = -hd + fraction_upwind[c] * q;
      }
      T_up[c * naq + i] = cu * area[c] * 1000.0;         // :813-814
      T_dn[c * naq + i] = cd * area[c] * 1000.0;
    }
  }
  return 0;
}

// Interior-connection loop of RTResidualFlux, reactive_transport.F90:2252-2310, with TFlux, transport.F90:368-439.
// id_up/id_dn: ghosted ids (0-based); g2l: ghosted -> local (0-based, < 0 for ghost cells; NULL = identity);
// active: imat > 0 per ghosted cell (NULL = all).  r [nlocal][naq] is zeroed first (r_p = 0.d0, :2249).
int orc_flux_residual(const View *v, const uint8_t *active, int naq, int64_t nconn, const int32_t *id_up,
                      const int32_t *id_dn, const int32_t *g2l, const double *T_up, const double *T_dn, int64_t nlocal,
                      double *r) {
  std::fill(r, r + nlocal * naq, 0.0);
  const double *tot = v->f[RXN_F_TOTAL];
  std::vector<double> Res(naq);
  for (int64_t c = 0; c < nconn; ++c) {
    const int64_t gu = id_up[c], gd = id_dn[c];
    if (active && (!active[gu] || !active[gd])) continue;               // :2264-2265
    for (int i = 0; i < naq; ++i)                                        // transport.F90:402-403
      Res[i] = T_up[c * naq + i] * tot[i * v->ld + gu] + T_dn[c * naq + i] * tot[i * v->ld + gd];
    const int64_t lu = g2l ? g2l[gu] : gu, ld_ = g2l ? g2l[gd] : gd;
    if (lu >= 0) for (int i = 0; i < naq; ++i) r[lu * naq + i] = r[lu * naq + i] + Res[i];     // :2298-2302
    if (ld_ >= 0) for (int i = 0; i < naq; ++i) r[ld_ * naq + i] = r[ld_ * naq + i] - Res[i];  // :2304-2308
  }
  return 0;
}

// Interior-connection loop of RTJacobianFlux, reactive_transport.F90:3094-3140, with TFluxDerivative,
// transport.F90:529-622: Jup(i,j) = dtotal_up(i,j) coef_up(i), Jdn likewise; row up gets (up,up) += Jup,
// (up,dn) += Jdn; row dn gets (dn,dn) += -Jdn, (dn,up) += -Jup (MatSetValuesBlockedLocal, ADD_VALUES).
// The matrix is returned as block CSR over the local rows: slot 0 of a row is its diagonal block, then one slot
// per connection of the row in connection order; blocks column-major, col = ghosted id.  Returns the number of
// blocks (row_ptr[nlocal]); col/val may be NULL to size the arrays.
int64_t orc_flux_jacobian(const View *v, const uint8_t *active, int naq, int64_t nconn, const int32_t *id_up,
                          const int32_t *id_dn, const int32_t *g2l, const double *T_up, const double *T_dn, int64_t nlocal,
                          int64_t nghosted, int32_t *row_ptr, int32_t *col, double *val) {
  std::vector<int64_t> deg(nlocal, 1);
  auto skip = [&](int64_t c) { return active && (!active[id_up[c]] || !active[id_dn[c]]); };
  auto loc = [&](int64_t g) { return g2l ? (int64_t)g2l[g] : g; };
  for (int64_t c = 0; c < nconn; ++c) {
    if (skip(c)) continue;
    if (loc(id_up[c]) >= 0) ++deg[loc(id_up[c])];
    if (loc(id_dn[c]) >= 0) ++deg[loc(id_dn[c])];
  }
  row_ptr[0] = 0;
  for (int64_t r = 0; r < nlocal; ++r) row_ptr[r + 1] = (int32_t)(row_ptr[r] + deg[r]);
  const int64_t nnzb = row_ptr[nlocal];
  if (!col || !val) return nnzb;
  const int nn = naq * naq;
  std::fill(val, val + nnzb * nn, 0.0);
  for (int64_t g = 0; g < nghosted; ++g) if (loc(g) >= 0) col[row_ptr[loc(g)]] = (int32_t)g;
  std::vector<int64_t> cur(nlocal);
  for (int64_t r = 0; r < nlocal; ++r) cur[r] = row_ptr[r] + 1;
  const double *D = v->f[RXN_F_DTOTAL];
  std::vector<double> Jup(nn), Jdn(nn);
  for (int64_t c = 0; c < nconn; ++c) {
    if (skip(c)) continue;
    const int64_t gu = id_up[c], gd = id_dn[c];
    for (int j = 0; j < naq; ++j)
      for (int i = 0; i < naq; ++i) {                                   // transport.F90:575-582
        Jup[j * naq + i] = D[(int64_t)(j * naq + i) * v->ld + gu] * T_up[c * naq + i];
        Jdn[j * naq + i] = D[(int64_t)(j * naq + i) * v->ld + gd] * T_dn[c * naq + i];
      }
    const int64_t lu = loc(gu), ld_ = loc(gd);
    if (lu >= 0) {                                                       // :3122-3127
      double *dg = val + (int64_t)row_ptr[lu] * nn, *od = val + cur[lu] * nn;
      for (int e = 0; e < nn; ++e) { dg[e] = dg[e] + Jup[e]; od[e] = od[e] + Jdn[e]; }
      col[cur[lu]++] = (int32_t)gd;
    }
    if (ld_ >= 0) {                                                      // :3129-3137
      double *dg = val + (int64_t)row_ptr[ld_] * nn, *od = val + cur[ld_] * nn;
      for (int e = 0; e < nn; ++e) { dg[e] = dg[e] + (-Jdn[e]); od[e] = od[e] + (-Jup[e]); }
      col[cur[ld_]++] = (int32_t)gu;
    }
  }
  return nnzb;
}

// TSrcSinkCoef, transport.F90:901-954, liquid phase.  type: RXN_SS_* (include/rxn_b200.h).
int orc_ss_coefs(int64_t nconn, const double *qsrc, const int32_t *type, double *T_in, double *T_out) {
  for (int64_t c = 0; c < nconn; ++c) {
    T_in[c] = 0.0; T_out[c] = 0.0;                                       // :922-923
    switch (type[c]) {
      case RXN_SS_EQUILIBRIUM:                                            // :926-932
        T_in[c] = 1.0e-3;
        T_out[c] = -1.0 * T_in[c];
        break;
      case RXN_SS_MASS_RATE:                                              // :933-936
        T_in[c] = 0.0;
        T_out[c] = -1.0;
        break;
      default:                                                            // :937-947
        if (qsrc[c] > 0.0) { T_in[c] = 0.0; T_out[c] = -1.0 * qsrc[c] * 1000.0; }
        else { T_out[c] = 0.0; T_in[c] = -1.0 * qsrc[c] * 1000.0; }
    }
  }
  return 0;
}

// Boundary-connection loop of RTResidualFlux, reactive_transport.F90:2347-2430 (kind 0: r_p = r_p - Res with
// Res = TFlux(up = boundary auxvar, dn = cell), patch%boundary_tran_fluxes = -Res) and the source/sink loop of
// RTResidualNonFlux, :2623-2672 (kind 1: Res = coef_in total_cell + coef_out total_ss, r_p = r_p + Res,
// patch%ss_tran_fluxes = Res).  id_dn: ghosted id of the cell of each connection; ext_total [nconn][naq]: total of the
// boundary auxvar / of the source-sink constraint; c_ext / c_cell [nconn][naq]: coef_up / coef_dn (kind 0), coef_out /
// coef_in repeated per component (kind 1).  r [nlocal][naq] is updated in place; flux_out may be NULL.
int orc_coupler_residual(const View *v, const uint8_t *active, int kind, int naq, int64_t nconn, const int32_t *id_dn,
                         const int32_t *g2l, const double *ext_total, const double *c_ext, const double *c_cell, int64_t nlocal,
                         double *r, double *flux_out) {
  const double *tot = v->f[RXN_F_TOTAL];
  std::vector<double> Res(naq);
  for (int64_t c = 0; c < nconn; ++c) {
    const int64_t g = id_dn[c];
    if (active && !active[g]) continue;                                 // :2360, :2636
    const int64_t l = g2l ? g2l[g] : g;
    if (l < 0 || l >= nlocal) return 1;                                 // coupler connections sit on local cells
    if (kind == 0) {
      for (int i = 0; i < naq; ++i)                                      // transport.F90:402-403
        Res[i] = c_ext[c * naq + i] * ext_total[c * naq + i] + c_cell[c * naq + i] * tot[i * v->ld + g];
      for (int i = 0; i < naq; ++i) r[l * naq + i] = r[l * naq + i] - Res[i];                 // :2382
      if (flux_out) for (int i = 0; i < naq; ++i) flux_out[c * naq + i] = -Res[i];            // :2421-2424
    } else {
      for (int i = 0; i < naq; ++i)                                      // :2648-2653
        Res[i] = c_cell[c * naq + i] * tot[i * v->ld + g] + c_ext[c * naq + i] * ext_total[c * naq + i];
      for (int i = 0; i < naq; ++i) r[l * naq + i] = r[l * naq + i] + Res[i];                 // :2661
      if (flux_out) for (int i = 0; i < naq; ++i) flux_out[c * naq + i] = Res[i];             // :2662-2664
    }
  }
  return 0;
}

// Boundary-connection loop of RTJacobianFlux, reactive_transport.F90:3176-3240 (kind 0: Jdn(i,j) = dtotal(i,j) coef_dn(i),
// Jdn = -Jdn, added to the diagonal block) and the source/sink loop of RTJacobianNonFlux, :3394-3436 (kind 1:
// Jup = coef_in dtotal, added to the diagonal block).  diag [nlocal][naq*naq] column-major blocks, updated in place.
int orc_coupler_jacobian(const View *v, const uint8_t *active, int kind, int naq, int64_t nconn, const int32_t *id_dn,
                         const int32_t *g2l, const double *c_cell, int64_t nlocal, double *diag) {
  const double *D = v->f[RXN_F_DTOTAL];
  const int nn = naq * naq;
  std::vector<double> J(nn);
  for (int64_t c = 0; c < nconn; ++c) {
    const int64_t g = id_dn[c];
    if (active && !active[g]) continue;
    const int64_t l = g2l ? g2l[g] : g;
    if (l < 0 || l >= nlocal) return 1;
    for (int j = 0; j < naq; ++j)
      for (int i = 0; i < naq; ++i) {
        if (kind == 0) {
          J[j * naq + i] = 0.0 + D[(int64_t)(j * naq + i) * v->ld + g] * c_cell[c * naq + i];   // transport.F90:572-582
          J[j * naq + i] = -J[j * naq + i];                                                      // :3215
        } else {
          J[j * naq + i] = 0.0 + c_cell[c * naq + i] * D[(int64_t)(j * naq + i) * v->ld + g];   // :3425-3429
        }
      }
    double *dg = diag + l * nn;
    for (int e = 0; e < nn; ++e) dg[e] = dg[e] + J[e];                  // MatSetValuesBlockedLocal(ADD_VALUES) :3217, :3430
  }
  return 0;
}

int orc_desc_size(void) { return (int)sizeof(RxnTablesDesc); }

}  // extern "C"
