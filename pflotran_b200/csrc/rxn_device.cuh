// rxn_device.cuh — per-cell device routines of the reaction path (one thread = one cell).
//
// Each routine is the batched restatement of one reference routine (file:line cited),
// written for the SoA state in HBM (rxn_tab.h) and the shared-memory table blob.
// Loop / operation order follows the reference so results agree with the reference's
// arithmetic to rounding; deliberate re-associations are marked "REASSOC".
#pragma once
#include <math.h>
#include "rxn_tab.h"

namespace rxn {

#define RXN_LOG_TO_LN 2.30258509299           /* pflotran_constants.F90:48 (truncated on purpose) */
#define RXN_IDEAL_GAS_CONSTANT 8.31446        /* pflotran_constants.F90:53 */

struct Tab {               // shared-memory resident tables
  const double *d;
  const int *i;
  const DevTab *h;
};

// per-thread cell context.  N = compile-time bound on naqcomp.
template <int N>
struct Cell {
  double m[N];        // pri_molal
  double gam[N];      // pri_act_coef
  double lnc[N];      // ln(pri_molal)
  double lna[N];      // ln(pri_molal) + ln(pri_act_coef)
  double total[N];    // total(:,1) [mol/L]
  double tsorb[N];    // total_sorb_eq
  double ln_act_h2o, den_kg, sat, temp, pres, volume, porosity, soil_density;
  double hpt[6];      // tr, pr, log10(tr), sqrt(tr), 1/tr, 1/pr of this cell (hpt logK fit; filled by load_cell for RXN_LOGK_HPT tables)
  double im[RXN_MAX_IMMOBILE];   // rt_auxvar%immobile [mol/m^3 bulk]: dofs naq .. ncomp-1
  long long cell;     // state index
  int flags;
};

__device__ __forceinline__ double &G(const DevState &S, int field, long long row, long long cell) {
  return S.f[field][row * S.ld + cell];
}

// A cell of a global-implicit entry point raised a flag the reference would stop on (activity-coefficient Newton diverged,
// reaction.F90:3864; free-site / ion-exchange iteration that never ends; non-finite result): OR it into the launch's word, the
// entry point then returns RXN_ERR_CELL_FAILED.  (RReact reports per cell through flags_out instead.)
__device__ __forceinline__ void report_cell_flags(const DevState &S, int flags) {
  if (flags != 0 && S.fail) {
#ifdef __CUDA_ARCH__
    atomicOr(S.fail, (unsigned int)flags);
#else
    __atomic_fetch_or(S.fail, (unsigned int)flags, __ATOMIC_RELAXED);
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// logK at the cell's T (and P): reaction_aux.F90:1461-1488 (5-term fit), :1529-1571 (hpt).
// The reference overwrites the shared tables per cell (reaction.F90:5433-5524); here it is a
// pure function of the cell's temp/pres evaluated where it is needed.
__device__ __forceinline__ double logK_fit5(const double *c, double temp) {
  double tk = temp + 273.15;
  return c[0] * log(tk) + c[1] + c[2] * tk + c[3] / tk + c[4] / (tk * tk);
}
// x / d with r = RN(1/d): q = RN(x r) is within 1 ulp, the remainder fma(-d, q, x) is exact, and RN(q + rem r) is the correctly
// rounded quotient (Markstein) - the same bits as the division, at 3 FMA-class instructions instead of a ~25-instruction DDIV.
__device__ __forceinline__ double div_by(double x, double d, double r) {
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
}
// hpt form, reaction_aux.F90:1529-1571.  The reference's expression term by term; the 9 divisions by tr / pr go through
// div_by (identical results: the fit's terms cancel, so no term may change even in its last bit - a precomputed-basis dot
// product moved TOTAL by 3.8e-10 and was rejected).
__device__ __forceinline__ void hpt_terms(double temp, double pres, double *t) {
  const double tk = temp + 273.15, tr = tk / 273.15, pr = pres / 1.0e7;
  t[0] = tr; t[1] = pr; t[2] = log(tr) / log(10.0); t[3] = sqrt(tr); t[4] = 1.0 / tr; t[5] = 1.0 / pr;
}
__device__ __forceinline__ double logK_hpt(const double *c, const double *t) {
  const double tr = t[0], pr = t[1], logtr = t[2], sqtr = t[3], itr = t[4], ipr = t[5];
  return c[0] + c[1] * tr + div_by(c[2], tr, itr) + c[3] * logtr + c[4] * tr * tr + div_by(div_by(c[5], tr, itr), tr, itr) +
         c[6] * sqtr + c[7] * pr + c[8] * pr * tr + div_by(c[9] * pr, tr, itr) + c[10] * pr * logtr +
         div_by(c[11], pr, ipr) + div_by(c[12], pr, ipr) * tr + div_by(div_by(c[13], pr, ipr), tr, itr) + c[14] * pr * pr + c[15] * pr * pr * tr +
         div_by(c[16] * pr * pr, tr, itr);
}
// srf_list: surface-complex logK is temperature-updated only for the 5-term form
// (reaction.F90:5517-5521: hpt not implemented there).
template <int N>
__device__ __forceinline__ double logK_of(const Tab &T, const DSpec &s, int r, const Cell<N> &c, bool srf_list) {
  const DevTab &h = *T.h;
  if (h.logK_mode == RXN_LOGK_FIXED || s.o_coef < 0) return T.d[s.o_logK + r];
  if (h.logK_mode == RXN_LOGK_HPT) {
    if (srf_list) return T.d[s.o_logK + r];
    return logK_hpt(T.d + s.o_coef + r * h.ncoef, c.hpt);
  }
  return logK_fit5(T.d + s.o_coef + r * h.ncoef, c.temp);
}

__device__ __forceinline__ double dh_gamma(double Z, double a0, double sqrt_I, double I, double A, double B, double Bdot) {
  return exp((-Z * Z * sqrt_I * A / (1.0 + a0 * B * sqrt_I) + Bdot * I) * RXN_LOG_TO_LN);
}

// ---------------------------------------------------------------------------------------------
// RActivityCoefficients — reaction.F90:3812-4053
template <int N>
__device__ void activity_coefficients(const Tab &T, const DevState &S, Cell<N> &c) {
  const DevTab &h = *T.h;
  const int naq = h.naq, ncplx = h.ncplx;
  const double *Z = T.d + h.o_Z, *a0 = T.d + h.o_a0, *cZ = T.d + h.o_cplxZ, *ca0 = T.d + h.o_cplxa0;
  const double A = h.debyeA, B = h.debyeB, Bdot = h.debyeBdot;
  double sum_pri_molal = 0.0;
  if (h.use_act_h2o) {
    for (int j = 0; j < naq; ++j)
      if (j + 1 != h.h2o_aq_id) sum_pri_molal = sum_pri_molal + c.m[j];
  }
  if (h.act_alg == RXN_ACT_COEF_ALGORITHM_NEWTON) {  // :3846-3990
    for (int j = 0; j < naq; ++j) { c.lnc[j] = log(c.m[j]); c.lna[j] = c.lnc[j] + log(c.gam[j]); }
    double fpri = 0.0;
    for (int j = 0; j < naq; ++j) fpri = fpri + c.m[j] * Z[j] * Z[j];
    int it = 0;
    double II = 0.0, I = 0.0, f = 0.0;
    for (;;) {
      it = it + 1;
      if (it > 50) {  // reference poisons the state with NaN and spins (:3864-3875)
        double nan_ = nan("");
        for (int j = 0; j < naq; ++j) { c.m[j] = nan_; c.gam[j] = nan_; }
        for (int k = 0; k < ncplx; ++k) G(S, RXN_F_SEC_ACT_COEF, k, c.cell) = nan_;
        c.flags |= RXN_FLAG_ACT_DIVERGED;
        return;
      }
      I = fpri;
      for (int k = 0; k < ncplx; ++k) I = I + G(S, RXN_F_SEC_MOLAL, k, c.cell) * cZ[k] * cZ[k];
      I = 0.5 * I;
      f = I;
      if (fabs(I - II) < 1.0e-6 * I) break;
      if (ncplx > 0) {
        double didi = 0.0;
        double sqrt_I = sqrt(I);
        for (int k = 0; k < ncplx; ++k) {
          if (fabs(cZ[k]) > 0.0) {
            double tmp = 1.0 + B * ca0[k] * sqrt_I;
            double sum = 0.5 * A * cZ[k] * cZ[k] / (sqrt_I * (tmp * tmp)) - Bdot;
            for (int p = T.i[h.cplx.o_ptr + k]; p < T.i[h.cplx.o_ptr + k + 1]; ++p) {
              int j = T.i[h.cplx.o_id + p];
              if (fabs(Z[j]) > 0.0) {
                double tp = 1.0 + B * a0[j] * sqrt_I;
                double dgamdi = -0.5 * A * (Z[j] * Z[j]) / (sqrt_I * (tp * tp)) + Bdot;
                sum = sum + T.d[h.cplx.o_st + p] * dgamdi;
              }
            }
            double dcdi = G(S, RXN_F_SEC_MOLAL, k, c.cell) * RXN_LOG_TO_LN * sum;
            didi = didi + 0.5 * cZ[k] * cZ[k] * dcdi;
          }
        }
        double den = 1.0 - didi;
        if (fabs(den) > 0.0) II = (f - I * didi) / den; else II = f;
      } else {
        II = f;
      }
      I = II;
      double sqrt_I = sqrt(I);
      for (int i = 0; i < naq; ++i)
        c.gam[i] = (fabs(Z[i]) > 0.0) ? dh_gamma(Z[i], a0[i], sqrt_I, I, A, B, Bdot) : 1.0;
      double sum_sec_molal = 0.0;
      for (int k = 0; k < ncplx; ++k) {
        double g = (fabs(cZ[k]) > 0.0) ? dh_gamma(cZ[k], ca0[k], sqrt_I, I, A, B, Bdot) : 1.0;
        G(S, RXN_F_SEC_ACT_COEF, k, c.cell) = g;
        double lnQK = -logK_of(T, h.cplx, k, c, false) * RXN_LOG_TO_LN;
        double h2ost = T.d[h.cplx.o_h2ost + k];
        if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
        for (int p = T.i[h.cplx.o_ptr + k]; p < T.i[h.cplx.o_ptr + k + 1]; ++p)
          lnQK = lnQK + T.d[h.cplx.o_st + p] * c.lna[T.i[h.cplx.o_id + p]];
        double sm = exp(lnQK) / g;
        G(S, RXN_F_SEC_MOLAL, k, c.cell) = sm;
        sum_sec_molal = sum_sec_molal + sm;
      }
      if (h.use_act_h2o) {
        c.ln_act_h2o = 1.0 - 0.017 * (sum_pri_molal + sum_sec_molal);
        c.ln_act_h2o = (c.ln_act_h2o > 0.0) ? log(c.ln_act_h2o) : 0.0;
      }
    }
  } else {  // LAG (default) :3994-4050
    double I = 0.0;
    for (int i = 0; i < naq; ++i) I = I + c.m[i] * Z[i] * Z[i];
    for (int k = 0; k < ncplx; ++k) I = I + G(S, RXN_F_SEC_MOLAL, k, c.cell) * cZ[k] * cZ[k];
    I = 0.5 * I;
    double sqrt_I = sqrt(I);
    for (int i = 0; i < naq; ++i)
      c.gam[i] = (fabs(Z[i]) > 1.0e-10) ? dh_gamma(Z[i], a0[i], sqrt_I, I, A, B, Bdot) : 1.0;
    double sum_sec_molal = 0.0;
    for (int k = 0; k < ncplx; ++k) {
      double g = (fabs(cZ[k]) > 1.0e-10) ? dh_gamma(cZ[k], ca0[k], sqrt_I, I, A, B, Bdot) : 1.0;
      G(S, RXN_F_SEC_ACT_COEF, k, c.cell) = g;
      if (h.use_act_h2o) sum_sec_molal = sum_sec_molal + G(S, RXN_F_SEC_MOLAL, k, c.cell);
    }
    if (h.use_act_h2o) {
      c.ln_act_h2o = 1.0 - 0.017 * (sum_pri_molal + sum_sec_molal);
      c.ln_act_h2o = (c.ln_act_h2o > 0.0) ? log(c.ln_act_h2o) : 0.0;
    }
  }
}

template <int N>
__device__ __forceinline__ void compute_ln(const Tab &T, Cell<N> &c) {
  const int naq = T.h->naq;
  for (int i = 0; i < naq; ++i) { c.lnc[i] = log(c.m[i]); c.lna[i] = c.lnc[i] + log(c.gam[i]); }
}

// ---------------------------------------------------------------------------------------------
// RTotal — reaction.F90:4057-4158.  dtot: naq x naq column-major (ld = naq).
// Requires c.lnc / c.lna current (compute_ln).
template <int N>
__device__ void rtotal(const Tab &T, const DevState &S, Cell<N> &c, double *dtot) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  const double den_kg_per_L = c.den_kg * 1.0 * 1.0e-3;  // xmass = 1 (global_auxvar%xmass not associated)
  for (int i = 0; i < naq; ++i) c.total[i] = c.m[i];
  for (int e = 0; e < naq * naq; ++e) dtot[e] = 0.0;
  for (int i = 0; i < naq; ++i) dtot[i + i * naq] = 1.0;
  const int *ptr = T.i + h.cplx.o_ptr, *id = T.i + h.cplx.o_id;
  const double *st = T.d + h.cplx.o_st;
  for (int k = 0; k < h.ncplx; ++k) {
    double lnQK = -logK_of(T, h.cplx, k, c, false) * RXN_LOG_TO_LN;
    double h2ost = T.d[h.cplx.o_h2ost + k];
    if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
    const int p0 = ptr[k], p1 = ptr[k + 1];
    for (int p = p0; p < p1; ++p) lnQK = lnQK + st[p] * c.lna[id[p]];
    const double g = G(S, RXN_F_SEC_ACT_COEF, k, c.cell);
    const double sm = exp(lnQK) / g;
    G(S, RXN_F_SEC_MOLAL, k, c.cell) = sm;
    for (int p = p0; p < p1; ++p) c.total[id[p]] = c.total[id[p]] + st[p] * sm;
    for (int q = p0; q < p1; ++q) {
      const int jcomp = id[q];
      const double tempreal = st[q] * exp(lnQK - c.lnc[jcomp]) / g;
      for (int p = p0; p < p1; ++p) dtot[id[p] + jcomp * naq] = dtot[id[p] + jcomp * naq] + st[p] * tempreal;
    }
  }
  for (int i = 0; i < naq; ++i) c.total[i] = c.total[i] * den_kg_per_L;
  for (int e = 0; e < naq * naq; ++e) dtot[e] = dtot[e] * den_kg_per_L;
}

// ---------------------------------------------------------------------------------------------
// RTotalSorbEqSurfCplx1 — reaction_surf_complex.F90:658-934
// eq_conc_out: add srfcplx_conc into the EQSRFCPLX_CONC state rows (equilibrium rxns only).
template <int N>
__device__ void sorb_eq_surfcplx1(const Tab &T, const DevState &S, Cell<N> &c, int irxn, bool eq_conc_out,
                                  double *total_sorb, double *dtotal_sorb) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  const double tol = 1.0e-12;
  double srfcplx_conc[RXN_MAX_SRFCPLX_PER_RXN];
  double dSx_dmi[N];
  const int c0 = T.i[h.o_rxn_cptr + irxn], c1 = T.i[h.o_rxn_cptr + irxn + 1];
  const int *cid = T.i + h.o_rxn_cid;
  const int *sptr = T.i + h.srf.o_ptr, *sid = T.i + h.srf.o_id;
  const double *sst = T.d + h.srf.o_st, *site_st = T.d + h.o_srf_site_st;
  double free_site_conc = G(S, RXN_F_FREE_SITE_CONC, irxn, c.cell);
  double site_density;
  const int surf_type = T.i[h.o_rxn_surf_type + irxn];
  const double dens = T.d[h.o_rxn_density + irxn];
  if (surf_type == RXN_MINERAL_SURFACE)
    site_density = dens * G(S, RXN_F_MNRL_VOLFRAC, T.i[h.o_rxn_to_surf + irxn] - 1, c.cell);
  else if (surf_type == RXN_ROCK_SURFACE)
    site_density = dens * c.soil_density * (1.0 - c.porosity);
  else
    site_density = dens;
  if (site_density < 1.0e-40) return;
  const int stoich_flag = T.i[h.o_rxn_flag + irxn];
  bool one_more = false;
  int num_iterations = 0;
  double damping_factor = 1.0;
  double total;
  for (;;) {
    num_iterations = num_iterations + 1;
    total = free_site_conc;
    const double ln_free_site = log(free_site_conc);
    for (int j = c0; j < c1; ++j) {
      const int icplx = cid[j];
      double lnQK = -logK_of(T, h.srf, icplx, c, true) * RXN_LOG_TO_LN;
      const double h2ost = T.d[h.srf.o_h2ost + icplx];
      if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
      lnQK = lnQK + site_st[icplx] * ln_free_site;
      for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p) lnQK = lnQK + sst[p] * c.lna[sid[p]];
      const double sc = exp(lnQK);
      srfcplx_conc[j - c0] = sc;
      total = total + site_st[icplx] * sc;
    }
    if (one_more) break;
    if (stoich_flag) {
      double res = site_density - total;
      double dres_dfree_site = 1.0;
      for (int j = c0; j < c1; ++j)
        dres_dfree_site = dres_dfree_site + site_st[cid[j]] * srfcplx_conc[j - c0] / free_site_conc;
      double dfree_site_conc = res / dres_dfree_site;
      if (num_iterations > 1000) damping_factor = 0.5;
      free_site_conc = free_site_conc + damping_factor * dfree_site_conc;
      double rel_change = fabs(dfree_site_conc / free_site_conc);
      if (rel_change < tol) one_more = true;
      if (num_iterations > 100000) { c.flags |= RXN_FLAG_CAPPED; one_more = true; }  // reference would spin
    } else {
      total = total / free_site_conc;
      free_site_conc = site_density / total;
      one_more = true;
    }
  }
  G(S, RXN_F_FREE_SITE_CONC, irxn, c.cell) = free_site_conc;

  for (int i = 0; i < naq; ++i) dSx_dmi[i] = 0.0;
  double tempreal = 0.0;
  for (int j = c0; j < c1; ++j) {
    const int icplx = cid[j];
    const double sc = srfcplx_conc[j - c0];
    for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p)
      dSx_dmi[sid[p]] = dSx_dmi[sid[p]] + sst[p] * site_st[icplx] * sc;
    tempreal = tempreal + site_st[icplx] * site_st[icplx] * sc;
  }
  tempreal = tempreal / free_site_conc;
  tempreal = tempreal + 1.0;
  for (int i = 0; i < naq; ++i) dSx_dmi[i] = -dSx_dmi[i] / tempreal;
  for (int i = 0; i < naq; ++i) dSx_dmi[i] = dSx_dmi[i] / c.m[i];

  if (eq_conc_out)
    for (int j = c0; j < c1; ++j) G(S, RXN_F_EQSRFCPLX_CONC, cid[j], c.cell) += srfcplx_conc[j - c0];

  for (int k = c0; k < c1; ++k) {
    const int icplx = cid[k];
    const double sc = srfcplx_conc[k - c0];
    const int p0 = sptr[icplx], p1 = sptr[icplx + 1];
    for (int p = p0; p < p1; ++p) total_sorb[sid[p]] = total_sorb[sid[p]] + sst[p] * sc;
    const double nui_Si_over_Sx = site_st[icplx] * sc / free_site_conc;
    for (int q = p0; q < p1; ++q) {
      const int jcomp = sid[q];
      const double tr = sst[q] * sc / c.m[jcomp] + nui_Si_over_Sx * dSx_dmi[jcomp];
      for (int p = p0; p < p1; ++p)
        dtotal_sorb[sid[p] + jcomp * naq] = dtotal_sorb[sid[p] + jcomp * naq] + sst[p] * tr;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RTotalSorbEqIonx — reaction.F90:4305-4535
template <int N>
__device__ void sorb_eq_ionx(const Tab &T, const DevState &S, Cell<N> &c, double *dsorb) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  const double tol = 1.0e-12;
  const double *Z = T.d + h.o_Z;
  double cation_X[N];
  for (int r = 0; r < h.nionx * h.ionx_ld; ++r) G(S, RXN_F_EQIONX_CONC, r, c.cell) = 0.0;
  for (int irxn = 0; irxn < h.nionx; ++irxn) {
    const int p0 = T.i[h.o_ionx_ptr + irxn], ncomp = T.i[h.o_ionx_ptr + irxn + 1] - p0;
    const int *cat = T.i + h.o_ionx_cat + p0;
    const double *kk = T.d + h.o_ionx_k + p0;
    const int to_surf = T.i[h.o_ionx_to_surf + irxn];
    const double CEC = T.d[h.o_ionx_CEC + irxn];
    double omega;
    if (to_surf > 0) omega = fmax(CEC * G(S, RXN_F_MNRL_VOLFRAC, to_surf - 1, c.cell), 1.0e-40);
    else omega = CEC;
    for (int j = 0; j < naq; ++j) cation_X[j] = 0.0;
    if (T.i[h.o_ionx_Zflag + irxn]) {
      const int icomp = cat[0];
      const double ref_cation_conc = c.m[icomp] * c.gam[icomp];
      const double ref_cation_Z = Z[icomp];
      const double ref_cation_k = kk[0];
      double ref_cation_X = ref_cation_Z * G(S, RXN_F_EQIONX_REF_CATION_SORBED_CONC, irxn, c.cell) / omega;
      bool one_more = false;
      double KDj = ref_cation_X / (ref_cation_k * ref_cation_conc);
      int it = 0;
      for (;;) {
        it = it + 1;
        if (it > 20000) { c.flags |= RXN_FLAG_CAPPED; break; }  // reference: fatal error (:4382-4385)
        ref_cation_X = KDj * (ref_cation_k * ref_cation_conc);
        cation_X[0] = ref_cation_X;
        double total = ref_cation_X;
        double dres_dKDj = 0.0;
        for (int j = 1; j < ncomp; ++j) {
          const int ic = cat[j];
          cation_X[j] = kk[j] * c.m[ic] * c.gam[ic] * pow(KDj, Z[ic] / ref_cation_Z);
          total = total + cation_X[j];
          dres_dKDj = dres_dKDj + cation_X[j] / KDj * Z[ic];
        }
        dres_dKDj = dres_dKDj / ref_cation_Z + (ref_cation_k * ref_cation_conc);
        const double res = 1.0 - total;
        if (one_more) break;
        const double delta_KDj = res / dres_dKDj;
        KDj = KDj + delta_KDj;
        KDj = fmax(KDj, 1.0e-40);
        if (fabs(delta_KDj / KDj) < tol) one_more = true;
      }
      G(S, RXN_F_EQIONX_REF_CATION_SORBED_CONC, irxn, c.cell) = ref_cation_X * omega / ref_cation_Z;
    } else {
      double sumkm = 0.0;
      for (int j = 0; j < ncomp; ++j) {
        const int ic = cat[j];
        cation_X[j] = c.m[ic] * c.gam[ic] * kk[j];
        sumkm = sumkm + cation_X[j];
      }
      for (int j = 0; j < naq; ++j) cation_X[j] = cation_X[j] / sumkm;
    }
    double sumZX = 0.0;
    for (int i = 0; i < ncomp; ++i) sumZX = sumZX + Z[cat[i]] * cation_X[i];
    for (int i = 0; i < ncomp; ++i) {
      const int icomp = cat[i];
      const double tempreal1 = cation_X[i] * omega / Z[icomp];
      G(S, RXN_F_EQIONX_CONC, (long long)irxn * h.ionx_ld + i, c.cell) += tempreal1;
      c.tsorb[icomp] = c.tsorb[icomp] + tempreal1;
      const double tempreal2 = Z[icomp] / sumZX;
      for (int j = 0; j < ncomp; ++j) {
        const int jcomp = cat[j];
        const int e = icomp + jcomp * naq;
        if (i == j) dsorb[e] = dsorb[e] + tempreal1 * (1.0 - (tempreal2 * cation_X[j])) / c.m[jcomp];
        else dsorb[e] = dsorb[e] + (-tempreal1) * tempreal2 * cation_X[j] / c.m[jcomp];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RTotalSorbKD — reaction.F90:4220-4301
template <int N>
__device__ void sorb_kd(const Tab &T, const DevState &S, Cell<N> &c, double *dsorb) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  for (int irxn = 0; irxn < h.nkd; ++irxn) {
    const int icomp = T.i[h.o_kd_spec + irxn] - 1;
    const double molality = c.m[icomp];
    const int mn = T.i[h.o_kd_mnrl + irxn];
    const double coef = T.d[h.o_kd_coef + irxn];
    double kd_kgw_m3b;
    if (mn > 0)
      kd_kgw_m3b = coef * c.den_kg * (1.0 - c.porosity) * c.soil_density * 1.0e-3 * G(S, RXN_F_MNRL_VOLFRAC, mn - 1, c.cell);
    else
      kd_kgw_m3b = coef;
    double res, dres_dc;
    const int ty = T.i[h.o_kd_type + irxn];
    if (ty == RXN_SORPTION_LINEAR) {
      res = kd_kgw_m3b * molality;
      dres_dc = kd_kgw_m3b;
    } else if (ty == RXN_SORPTION_LANGMUIR) {
      const double tempreal = kd_kgw_m3b * molality;
      res = tempreal * T.d[h.o_kd_b + irxn] / (1.0 + tempreal);
      dres_dc = res / molality - res / (1.0 + tempreal) * tempreal / molality;
    } else if (ty == RXN_SORPTION_FREUNDLICH) {
      const double one_over_n = 1.0 / T.d[h.o_kd_n + irxn];
      res = kd_kgw_m3b * pow(molality, one_over_n);
      dres_dc = res / molality * one_over_n;
    } else {
      res = 0.0; dres_dc = 0.0;
    }
    c.tsorb[icomp] = c.tsorb[icomp] + res;
    dsorb[icomp + icomp * naq] = dsorb[icomp + icomp * naq] + dres_dc;
  }
}

// RZeroSorb + RTotalSorb — reaction.F90:4162-4216; RTotalSorbEqSurfCplx reaction_surf_complex.F90:441-502
template <int N>
__device__ void rtotal_sorb(const Tab &T, const DevState &S, Cell<N> &c, double *dsorb) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  for (int i = 0; i < naq; ++i) c.tsorb[i] = 0.0;
  for (int e = 0; e < naq * naq; ++e) dsorb[e] = 0.0;
  if (h.neq > 0)
    for (int k = 0; k < h.nsrf; ++k) G(S, RXN_F_EQSRFCPLX_CONC, k, c.cell) = 0.0;
  for (int ieq = 0; ieq < h.neq; ++ieq)
    sorb_eq_surfcplx1<N>(T, S, c, T.i[h.o_eq_rxn + ieq], true, c.tsorb, dsorb);
  if (h.nionx > 0) sorb_eq_ionx<N>(T, S, c, dsorb);
  if (h.nkd > 0) sorb_kd<N>(T, S, c, dsorb);
}

// RTAuxVarCompute — reaction.F90:4969-5006
template <int N>
__device__ __forceinline__ void auxvar_compute(const Tab &T, const DevState &S, Cell<N> &c, double *dtot, double *dsorb) {
  compute_ln<N>(T, c);
  rtotal<N>(T, S, c, dtot);
  if (T.h->neqsorb > 0) rtotal_sorb<N>(T, S, c, dsorb);
}

// ---------------------------------------------------------------------------------------------
// RKineticMineral — reaction_mineral.F90:564-1000.  Jac: ncomp x ncomp column-major.
// Requires c.lnc / c.lna current.  Prefactor species on secondary complexes are rejected at
// table creation (the reference branch clobbers its own loop variables, :977-979).
template <int N>
__device__ void kinetic_mineral(const Tab &T, const DevState &S, Cell<N> &c, double *Res, double *Jac, bool compute_derivative) {
  const DevTab &h = *T.h;
  const int n = h.naq;
  const int *ptr = T.i + h.kin.o_ptr, *id = T.i + h.kin.o_id;
  const double *st = T.d + h.kin.o_st;
  for (int im = 0; im < h.nkin; ++im) G(S, RXN_F_MNRL_RATE, im, c.cell) = 0.0;
  const int mp = h.maxpref > 1 ? h.maxpref : 1, mps = h.maxprefspec > 1 ? h.maxprefspec : 1;
  for (int imnrl = 0; imnrl < h.nkin; ++imnrl) {
    double lnQK = -logK_of(T, h.kin, imnrl, c, false) * RXN_LOG_TO_LN;
    const double h2ost = T.d[h.kin.o_h2ost + imnrl];
    if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
    const int p0 = ptr[imnrl], p1 = ptr[imnrl + 1];
    for (int p = p0; p < p1; ++p) lnQK = lnQK + st[p] * c.lna[id[p]];
    double QK;
    if (lnQK <= 6.90776) QK = exp(lnQK); else QK = 1.0e3;
    const double k_scale = h.has_scale ? T.d[h.o_k_scale + imnrl] : 1.0;
    const double k_Temkin = h.has_Temkin ? T.d[h.o_k_Temkin + imnrl] : 1.0;
    const double k_power = h.has_power ? T.d[h.o_k_power + imnrl] : 1.0;
    const double k_lim = T.d[h.o_k_lim + imnrl];
    const double k_aff = T.d[h.o_k_aff + imnrl];
    const int npref = T.i[h.o_k_npref + imnrl];
    double affinity_factor;
    if (h.has_Temkin) {
      if (h.has_scale) affinity_factor = 1.0 - pow(QK, 1.0 / (k_scale * k_Temkin));
      else affinity_factor = 1.0 - pow(QK, 1.0 / k_Temkin);
    } else if (h.has_scale) {
      affinity_factor = 1.0 - pow(QK, 1.0 / k_scale);
    } else {
      affinity_factor = 1.0 - QK;
    }
    const double sign_ = copysign(1.0, affinity_factor);
    double Im, Im_const, sum_prefactor_rate;
    double prefactor[RXN_MAX_PREF];
    double ln_prefactor_spec[RXN_MAX_PREF][RXN_MAX_PREF_SPEC];
    const double volfrac = G(S, RXN_F_MNRL_VOLFRAC, imnrl, c.cell);
    if (!(volfrac > 0 || sign_ < 0.0)) continue;
    if (k_aff > 0.0) {
      if (sign_ < 0.0 && QK < k_aff) continue;
    }
    if (k_lim > 0.0) affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / k_lim);
    if (npref > 0) {
      sum_prefactor_rate = 0.0;
      for (int ipref = 0; ipref < npref; ++ipref) {
        double ln_prefactor = 0.0;
        const int pb = imnrl * mp + ipref;
        const int nps = T.i[h.o_pref_id + pb * (h.maxprefspec + 1)];
        for (int ips = 0; ips < nps; ++ips) {
          const int icomp = T.i[h.o_pref_id + pb * (h.maxprefspec + 1) + ips + 1];
          const double ln_spec_act = c.lna[icomp - 1];
          const double ln_numerator = T.d[h.o_pref_alpha + pb * mps + ips] * ln_spec_act;
          const double ln_denominator =
              log(1.0 + exp(log(T.d[h.o_pref_atten + pb * mps + ips]) + T.d[h.o_pref_beta + pb * mps + ips] * ln_spec_act));
          ln_prefactor = ln_prefactor + ln_numerator;
          ln_prefactor = ln_prefactor - ln_denominator;
          ln_prefactor_spec[ipref][ips] = ln_numerator - ln_denominator;
        }
        prefactor[ipref] = exp(ln_prefactor);
        double arrhenius_factor = 1.0;
        const double Ea = T.d[h.o_pref_Ea + pb];
        if (Ea > 0.0) arrhenius_factor = exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
        sum_prefactor_rate = sum_prefactor_rate + prefactor[ipref] * T.d[h.o_pref_rate + pb] * arrhenius_factor;
      }
    } else {
      double arrhenius_factor = 1.0;
      const double Ea = T.d[h.o_k_Ea + imnrl];
      if (Ea > 0.0) arrhenius_factor = exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
      sum_prefactor_rate = T.d[h.o_k_rate + imnrl] * arrhenius_factor;
    }
    Im_const = -G(S, RXN_F_MNRL_AREA, imnrl, c.cell);
    if (h.has_scale) Im_const = Im_const / k_scale;
    if (h.has_power) Im = Im_const * sign_ * pow(fabs(affinity_factor), k_power) * sum_prefactor_rate;
    else Im = Im_const * sign_ * fabs(affinity_factor) * sum_prefactor_rate;
    G(S, RXN_F_MNRL_RATE, imnrl, c.cell) = Im;

    Im_const = Im_const * c.volume;
    Im = Im * c.volume;
    for (int p = p0; p < p1; ++p) Res[id[p]] = Res[id[p]] + st[p] * Im;
    if (!compute_derivative) continue;

    double dIm_dQK;
    if (h.has_power) dIm_dQK = -Im * k_power / fabs(affinity_factor);
    else dIm_dQK = -Im_const * sum_prefactor_rate;
    if (h.has_Temkin) {
      if (h.has_scale) dIm_dQK = dIm_dQK * (1.0 / (k_scale * k_Temkin)) / QK * (1.0 - affinity_factor);
      else dIm_dQK = dIm_dQK * (1.0 / k_Temkin) / QK * (1.0 - affinity_factor);
    } else if (h.has_scale) {
      dIm_dQK = dIm_dQK * (1.0 / k_scale) / QK * (1.0 - affinity_factor);
    }
    if (k_lim <= 0.0) {
      for (int q = p0; q < p1; ++q) {
        const int jcomp = id[q];
        const double dQK_dCj = st[q] * QK * exp(-c.lnc[jcomp]);
        const double dQK_dmj = dQK_dCj * c.den_kg * 1.0e-3;
        for (int p = p0; p < p1; ++p)
          Jac[id[p] + jcomp * n] = Jac[id[p] + jcomp * n] + st[p] * dIm_dQK * dQK_dmj;
      }
    } else {
      const double den = 1.0 + (1.0 - affinity_factor) / k_lim;
      for (int q = p0; q < p1; ++q) {
        const int jcomp = id[q];
        const double dQK_dCj = st[q] * QK * exp(-c.lnc[jcomp]);
        const double dQK_dmj = dQK_dCj * c.den_kg * 1.0e-3;
        for (int p = p0; p < p1; ++p)
          Jac[id[p] + jcomp * n] = Jac[id[p] + jcomp * n] + st[p] * dIm_dQK * (1.0 + QK / k_lim / den) * dQK_dmj / den;
      }
    }
    if (npref > 0) {
      const double dIm_dsum_prefactor_rate = Im / sum_prefactor_rate;
      for (int ipref = 0; ipref < npref; ++ipref) {
        const int pb = imnrl * mp + ipref;
        double arrhenius_factor = 1.0;
        const double Ea = T.d[h.o_pref_Ea + pb];
        if (Ea > 0.0) arrhenius_factor = exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
        const double ln_prefactor = log(prefactor[ipref]);
        const int nps = T.i[h.o_pref_id + pb * (h.maxprefspec + 1)];
        for (int ips = 0; ips < nps; ++ips) {
          const double dprefactor_dprefactor_spec = exp(ln_prefactor - ln_prefactor_spec[ipref][ips]);
          const int icomp = T.i[h.o_pref_id + pb * (h.maxprefspec + 1) + ips + 1];
          const double ln_spec_act = c.lna[icomp - 1], spec_act_coef = c.gam[icomp - 1];
          const double alpha = T.d[h.o_pref_alpha + pb * mps + ips], beta = T.d[h.o_pref_beta + pb * mps + ips],
                       atten = T.d[h.o_pref_atten + pb * mps + ips];
          const double dnum = alpha * exp(ln_prefactor_spec[ipref][ips] - ln_spec_act);
          const double ln_gam_m_beta = beta * ln_spec_act;
          const double denominator = 1.0 + exp(log(atten) + ln_gam_m_beta);
          const double dden = -1.0 * exp(ln_prefactor_spec[ipref][ips]) / denominator * atten * beta * exp(ln_gam_m_beta - ln_spec_act);
          double dprefactor_spec_dspec = dnum + dden;
          dprefactor_spec_dspec = dprefactor_spec_dspec * spec_act_coef;
          const double dIm_dspec = dIm_dsum_prefactor_rate * dprefactor_dprefactor_spec * dprefactor_spec_dspec *
                                   T.d[h.o_pref_rate + pb] * arrhenius_factor;
          for (int p = p0; p < p1; ++p)
            Jac[id[p] + (icomp - 1) * n] = Jac[id[p] + (icomp - 1) * n] + st[p] * dIm_dspec;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RMultiRateSorption — reaction_surf_complex.F90:566-654.
// REASSOC: the reference adds, per rate r, V*k_r/(1+k_r dt)*(f_r*S_eq - S_r) to Res and
// V*k_r/(1+k_r dt)*f_r*dS_eq to Jac (nrate x naq^2 FMAs per Newton iteration).  S_r does not
// change inside a Newton solve, so the per-rate sums  R0[i] = sum_r kk_r*S_r[i]  and
// K1 = sum_r kk_r*f_r  are formed once per call (mr_pre) and each iteration applies
// Res += V*(K1*S_eq - R0), Jac += V*K1*dS_eq.  Differs from the reference by re-association only.
template <int N>
__device__ void multirate_prepare(const Tab &T, const DevState &S, const Cell<N> &c, double dt, double *mrK1, double *mrR0) {
  // mrK1[nmr], mrR0[nmr*naq]
  const DevTab &h = *T.h;
  const int naq = h.naq;
  for (int ikr = 0; ikr < h.nmr; ++ikr) {
    double K1 = 0.0;
    for (int i = 0; i < naq; ++i) mrR0[ikr * naq + i] = 0.0;
    const int nrate = T.i[h.o_mr_nrate + ikr];
    for (int irate = 0; irate < nrate; ++irate) {
      const double rate = T.d[h.o_mr_rate + ikr * h.mr_ld + irate], frac = T.d[h.o_mr_frac + ikr * h.mr_ld + irate];
      const double kdt = rate * dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
      const long long row0 = ((long long)ikr * (h.mr_ld + 1) + (irate + 1)) * naq;
      for (int i = 0; i < naq; ++i) mrR0[ikr * naq + i] = mrR0[ikr * naq + i] + kk * G(S, RXN_F_KINMR_TOTAL_SORB, row0 + i, c.cell);
    }
    mrK1[ikr] = K1;
  }
}

template <int N>
__device__ void multirate_sorption(const Tab &T, const DevState &S, Cell<N> &c, const double *mrK1, const double *mrR0,
                                   double *Res, double *Jac, bool compute_derivative, double *scratch_dsorb) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  double total_sorb_eq[N];
  for (int ikr = 0; ikr < h.nmr; ++ikr) {
    const int irxn = T.i[h.o_mr_rxn + ikr];
    for (int i = 0; i < naq; ++i) total_sorb_eq[i] = 0.0;
    for (int e = 0; e < naq * naq; ++e) scratch_dsorb[e] = 0.0;
    sorb_eq_surfcplx1<N>(T, S, c, irxn, false, total_sorb_eq, scratch_dsorb);
    const double K1 = mrK1[ikr];
    for (int i = 0; i < naq; ++i) Res[i] = Res[i] + c.volume * (K1 * total_sorb_eq[i] - mrR0[ikr * naq + i]);
    if (compute_derivative) {
      const double cc = c.volume * K1;
      for (int e = 0; e < naq * naq; ++e) Jac[e] = Jac[e] + cc * scratch_dsorb[e];
    }
    const long long row0 = (long long)ikr * (h.mr_ld + 1) * naq;
    for (int i = 0; i < naq; ++i) G(S, RXN_F_KINMR_TOTAL_SORB, row0 + i, c.cell) = total_sorb_eq[i];
  }
}

// ---------------------------------------------------------------------------------------------
// ludcmp / lubksb — utility.F90:393-476, 480-523 (Crout, implicit scaling, partial pivoting,
// `>=` tie-break = last maximum wins, tiny = 1e-20).  A column-major n x n.
// ---------------------------------------------------------------------------------------------
// RKineticSurfCplx — reaction_surf_complex.F90:938-1137.  One kinetic reaction = surface complexation reaction 1 on
// mineral surface 1 (rxn_pack.h: the only configuration in which the reference's site / complex indices are in bounds);
// S^k, S^{k+1} and the free-site concentration live in the state (KINSRFCPLX_*), S^k -> S^{k+1} in RUpdateKineticState.
template <int N>
__device__ void kinetic_surfcplx(const Tab &T, const DevState &S, Cell<N> &c, double dt, double *Res, double *Jac, bool compute_derivative) {
  const DevTab &h = *T.h;
  const int n = h.naq;
  const int c0 = T.i[h.o_rxn_cptr + h.kin_rxn], c1 = T.i[h.o_rxn_cptr + h.kin_rxn + 1];
  const int *sptr = T.i + h.srf.o_ptr, *sid = T.i + h.srf.o_id;
  const double *sst = T.d + h.srf.o_st, *sh2o = T.d + h.srf.o_h2ost, *kf = T.d + h.o_kin_kf, *kb = T.d + h.o_kin_kb;
  double Q[RXN_MAX_SRFCPLX_PER_RXN];
  for (int j = c0; j < c1; ++j) {
    const int icplx = T.i[h.o_rxn_cid + j];
    double lnQ = 0.0;
    if (sh2o[icplx] != 0.0) lnQ = lnQ + sh2o[icplx] * c.ln_act_h2o;
    for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p) lnQ = lnQ + sst[p] * c.lna[sid[p]];
    Q[j - c0] = exp(lnQ);
  }
  double numerator_sum = 0.0;
  for (int j = c0; j < c1; ++j) {
    const int icplx = T.i[h.o_rxn_cid + j];
    numerator_sum = numerator_sum + G(S, RXN_F_KINSRFCPLX_CONC, icplx, c.cell) / (1.0 + kb[icplx] * dt);
  }
  numerator_sum = T.d[h.o_rxn_density + 0] - numerator_sum;
  double denominator_sum = 1.0;
  for (int j = c0; j < c1; ++j) {
    const int icplx = T.i[h.o_rxn_cid + j];
    denominator_sum = denominator_sum + (kf[icplx] * dt) / (1.0 + kb[icplx] * dt) * Q[j - c0];
  }
  for (int j = c0; j < c1; ++j) {
    const int icplx = T.i[h.o_rxn_cid + j];
    const double conc_k = G(S, RXN_F_KINSRFCPLX_CONC, icplx, c.cell);
    const double denominator = 1.0 + kb[icplx] * dt;
    const double conc_kp1 = (conc_k + kf[icplx] * dt * numerator_sum / denominator_sum * Q[j - c0]) / denominator;
    G(S, RXN_F_KINSRFCPLX_CONC_KP1, icplx, c.cell) = conc_kp1;
    for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p)
      Res[sid[p]] = Res[sid[p]] + sst[p] * (conc_kp1 - conc_k) / dt * c.volume;
  }
  G(S, RXN_F_KINSRFCPLX_FREE_SITE_CONC, 0, c.cell) = numerator_sum / denominator_sum;
  if (compute_derivative) {
    double fac_sum[N];
    for (int i = 0; i < n; ++i) fac_sum[i] = 0.0;
    for (int j = c0; j < c1; ++j) {
      const int icplx = T.i[h.o_rxn_cid + j];
      const double denominator = 1.0 + kb[icplx] * dt;
      const double fac = kf[icplx] / denominator;
      for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p) fac_sum[sid[p]] = fac_sum[sid[p]] + sst[p] * fac * Q[j - c0];
    }
    for (int j = c0; j < c1; ++j) {
      const int icplx = T.i[h.o_rxn_cid + j];
      const double denominator = 1.0 + kb[icplx] * dt;
      const double fac = kf[icplx] / denominator;
      for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p) {
        const int jcomp = sid[p];
        for (int q = sptr[icplx]; q < sptr[icplx + 1]; ++q) {
          const int lcomp = sid[q];
          Jac[jcomp + lcomp * n] = Jac[jcomp + lcomp * n] +
              (sst[p] * fac * numerator_sum * Q[j - c0] * (sst[q] - dt * fac_sum[lcomp] / denominator_sum)) / denominator_sum *
                  exp(-c.lnc[lcomp]) * c.volume;
        }
      }
    }
  }
}

// RRadioactiveDecay — reaction.F90:4607-4690.  dtot / dsorb: d total / d free-ion and d total_sorb_eq / d free-ion
// (column-major) of this iterate.
template <int N>
__device__ void radioactive_decay(const Tab &T, const Cell<N> &c, const double *dtot, const double *dsorb, double *Res, double *Jac,
                                  bool compute_derivative) {
  const DevTab &h = *T.h;
  const int n = h.naq;
  const double L_water = c.porosity * c.sat * c.volume * 1.0e3;
  for (int irxn = 0; irxn < h.ndecay; ++irxn) {
    const int jcomp = T.i[h.o_dec_fwd + irxn];
    double sum = c.total[jcomp] * L_water;
    sum = sum + (h.neqsorb > 0 ? c.tsorb[jcomp] : 0.0) * c.volume;
    const double kf = T.d[h.o_dec_kf + irxn];
    const double rate = sum * kf;
    const int p0 = T.i[h.o_dec_ptr + irxn], p1 = T.i[h.o_dec_ptr + irxn + 1];
    for (int p = p0; p < p1; ++p) Res[T.i[h.o_dec_id + p]] = Res[T.i[h.o_dec_id + p]] - T.d[h.o_dec_st + p] * rate;
    if (!compute_derivative) continue;
    const double tempreal = -1.0 * kf;
    for (int p = p0; p < p1; ++p) {
      const int icomp = T.i[h.o_dec_id + p];
      for (int j = 0; j < n; ++j)
        Jac[icomp + j * n] = Jac[icomp + j * n] + tempreal * T.d[h.o_dec_st + p] *
                                                      (dtot[jcomp + j * n] * L_water + (h.neqsorb > 0 ? dsorb[jcomp + j * n] : 0.0) * c.volume);
    }
  }
}

// RGeneral — reaction.F90:4694-4831
template <int N>
__device__ void general_reaction(const Tab &T, const Cell<N> &c, double *Res, double *Jac, bool compute_derivative) {
  const DevTab &h = *T.h;
  const int n = h.naq;
  for (int irxn = 0; irxn < h.ngen; ++irxn) {
    const double kf = T.d[h.o_gen_kf + irxn], kr = T.d[h.o_gen_kr + irxn];
    const int f0 = T.i[h.o_genf_ptr + irxn], f1 = T.i[h.o_genf_ptr + irxn + 1];
    const int b0 = T.i[h.o_genb_ptr + irxn], b1 = T.i[h.o_genb_ptr + irxn + 1];
    const int g0 = T.i[h.o_gen_ptr + irxn], g1 = T.i[h.o_gen_ptr + irxn + 1];
    double Qkf, lnQkf = 0.0, Qkr, lnQkr = 0.0;
    if (kf > 0.0) {
      lnQkf = log(kf);
      for (int p = f0; p < f1; ++p) lnQkf = lnQkf + T.d[h.o_genf_st + p] * c.lna[T.i[h.o_genf_id + p]];
      Qkf = exp(lnQkf);
    } else {
      Qkf = 0.0;
    }
    if (kr > 0.0) {
      lnQkr = log(kr);
      for (int p = b0; p < b1; ++p) lnQkr = lnQkr + T.d[h.o_genb_st + p] * c.lna[T.i[h.o_genb_id + p]];
      Qkr = exp(lnQkr);
    } else {
      Qkr = 0.0;
    }
    const double por_den_sat_vol = c.porosity * c.den_kg * c.sat * c.volume;
    for (int p = g0; p < g1; ++p)
      Res[T.i[h.o_gen_id + p]] = Res[T.i[h.o_gen_id + p]] - T.d[h.o_gen_st + p] * (Qkf - Qkr) * por_den_sat_vol;
    if (!compute_derivative) continue;
    if (kf > 0.0) {
      for (int q = f0; q < f1; ++q) {
        const int jcomp = T.i[h.o_genf_id + q];
        const double tempreal = -1.0 * T.d[h.o_genf_st + q] * exp(lnQkf - c.lnc[jcomp]) * por_den_sat_vol;
        for (int p = g0; p < g1; ++p) Jac[T.i[h.o_gen_id + p] + jcomp * n] = Jac[T.i[h.o_gen_id + p] + jcomp * n] + T.d[h.o_gen_st + p] * tempreal;
      }
    }
    if (kr > 0.0) {
      for (int q = b0; q < b1; ++q) {
        const int jcomp = T.i[h.o_genb_id + q];
        const double tempreal = T.d[h.o_genb_st + q] * exp(lnQkr - c.lnc[jcomp]) * por_den_sat_vol;
        for (int p = g0; p < g1; ++p) Jac[T.i[h.o_gen_id + p] + jcomp * n] = Jac[T.i[h.o_gen_id + p] + jcomp * n] + T.d[h.o_gen_st + p] * tempreal;
      }
    }
  }
}

template <int N>
__device__ int ludcmp(double *A, int n, int *indx) {
  const double tiny = 1.0e-20;
  double vv[N];
#define AA(i, j) A[(i) + (j) * n]
  for (int i = 0; i < n; ++i) {
    double aamax = 0.0;
    for (int j = 0; j < n; ++j) if (fabs(AA(i, j)) > aamax) aamax = fabs(AA(i, j));
    if (aamax <= 0.0) return 1;
    vv[i] = 1.0 / aamax;
  }
  int imax = 0;
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < j; ++i) {
      double sum = AA(i, j);
      for (int k = 0; k < i; ++k) sum = sum - AA(i, k) * AA(k, j);
      AA(i, j) = sum;
    }
    double aamax = 0.0;
    for (int i = j; i < n; ++i) {
      double sum = AA(i, j);
      for (int k = 0; k < j; ++k) sum = sum - AA(i, k) * AA(k, j);
      AA(i, j) = sum;
      const double dum = vv[i] * fabs(sum);
      if (dum >= aamax) { imax = i; aamax = dum; }
    }
    if (j != imax) {
      for (int k = 0; k < n; ++k) { const double dum = AA(imax, k); AA(imax, k) = AA(j, k); AA(j, k) = dum; }
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (AA(j, j) == 0.0) AA(j, j) = tiny;
    if (j != n - 1) {
      const double dum = 1.0 / AA(j, j);
      for (int i = j + 1; i < n; ++i) AA(i, j) = AA(i, j) * dum;
    }
  }
  return 0;
}

__device__ inline void lubksb(const double *A, int n, const int *indx, double *B) {
  int ii = -1;
  for (int i = 0; i < n; ++i) {
    const int ll = indx[i];
    double sum = B[ll];
    B[ll] = B[i];
    if (ii != -1) {
      for (int j = ii; j < i; ++j) sum = sum - AA(i, j) * B[j];
    } else if (sum != 0.0) {
      ii = i;
    }
    B[i] = sum;
  }
  for (int i = n - 1; i >= 0; --i) {
    double sum = B[i];
    for (int j = i + 1; j < n; ++j) sum = sum - AA(i, j) * B[j];
    B[i] = sum / AA(i, i);
  }
#undef AA
}

// RSolve — reaction.F90:4835-4880.  On return Res holds the update.
template <int N>
__device__ int rsolve(double *Res, double *Jac, const double *conc, int n, bool use_log) {
  int indices[N];
  for (int i = 0; i < n; ++i) {
    double mx = 0.0;
    for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(Jac[i + j * n]));
    double norm = fmax(1.0, mx);
    norm = 1.0 / norm;
    Res[i] = Res[i] * norm;
    for (int j = 0; j < n; ++j) Jac[i + j * n] = Jac[i + j * n] * norm;
  }
  if (use_log) {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) Jac[i + j * n] = Jac[i + j * n] * conc[j];
  }
  if (ludcmp<N>(Jac, n, indices)) return 1;
  lubksb(Jac, n, indices, Res);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Immobile dofs (reaction%ncomp = naqcomp + nimmobile).  The routines above assemble the naq x naq aqueous block with leading
// dimension naq; before the routines that touch immobile rows / columns run, the block is re-strided in place to ncomp x ncomp,
// the new rows and columns zeroed.  RMicrobial and RImmobileDecay come last in RReaction's dispatch order (reaction.F90:3563-3571),
// so every entry's terms are still summed in the reference's order.
__device__ __forceinline__ void expand_block(double *J, int naq, int n) {
  if (n == naq) return;
  for (int j = n - 1; j >= 0; --j)
    for (int i = n - 1; i >= 0; --i) J[i + j * n] = (i < naq && j < naq) ? J[i + j * naq] : 0.0;
}

// RMicrobial - reaction_microbial.F90:236-450.  Jac: ncomp x ncomp column-major.
template <int N>
__device__ void microbial_reaction(const Tab &T, const Cell<N> &c, double *Res, double *Jac, bool compute_derivative) {
  const DevTab &h = *T.h;
  const int n = h.ncomp, naq = h.naq;
  const double PI = 3.14159265359;               // pflotran_constants.F90:55
  double monod[RXN_MAX_MONOD], inhibition[RXN_MAX_MONOD];
  for (int irxn = 0; irxn < h.nmic; ++irxn) {
    const int s0 = T.i[h.o_mic_ptr + irxn], s1 = T.i[h.o_mic_ptr + irxn + 1];
    const int m0 = T.i[h.o_mic_mptr + irxn], m1 = T.i[h.o_mic_mptr + irxn + 1];
    const int i0 = T.i[h.o_mic_iptr + irxn], i1 = T.i[h.o_mic_iptr + irxn + 1];
    double Im = T.d[h.o_mic_k + irxn];
    if (h.mic_has_Ea) Im = Im * exp(T.d[h.o_mic_Ea + irxn] / RXN_IDEAL_GAS_CONSTANT * (1.0 / 298.15 - 1.0 / (c.temp + 273.15)));
    double yield = 0.0, biomass_conc = 0.0;
    for (int ii = m0; ii < m1; ++ii) {
      const int imonod = T.i[h.o_mic_mid + ii], icomp = T.i[h.o_mon_spec + imonod];
      const double activity = c.m[icomp] * c.gam[icomp], Cth = T.d[h.o_mon_Cth + imonod];
      monod[ii - m0] = (activity - Cth) / (T.d[h.o_mon_K + imonod] + activity - Cth);
      Im = Im * monod[ii - m0];
    }
    for (int ii = i0; ii < i1; ++ii) {
      const int iinh = T.i[h.o_mic_iid + ii], icomp = T.i[h.o_inh_spec + iinh];
      const double activity = c.m[icomp] * c.gam[icomp], C1 = T.d[h.o_inh_C + iinh];
      double v;
      switch (T.i[h.o_inh_type + iinh]) {
        case RXN_INHIBITION_MONOD: v = C1 / (C1 + activity); break;
        case RXN_INHIBITION_INVERSE_MONOD: v = activity / (C1 + activity); break;
        default: v = 0.5 + atan((activity - C1) * T.d[h.o_inh_C2 + iinh]) / PI; break;   // RXN_INHIBITION_THRESHOLD (other types rejected)
      }
      inhibition[ii - i0] = v;
      Im = Im * v;
    }
    const int ibiomass = T.i[h.o_mic_bio + irxn];
    const int immobile_id = naq + ibiomass;
    if (ibiomass >= 0) {
      biomass_conc = c.im[ibiomass];
      yield = T.d[h.o_mic_yield + irxn];
      Im = Im * biomass_conc;
    }
    const double por_sat_vol = c.porosity * c.sat * c.volume;
    Im = Im * 1.0e3 * por_sat_vol;
    for (int p = s0; p < s1; ++p) Res[T.i[h.o_mic_id + p]] = Res[T.i[h.o_mic_id + p]] - T.d[h.o_mic_st + p] * Im;
    if (ibiomass >= 0) Res[immobile_id] = Res[immobile_id] - yield * Im;
    if (!compute_derivative) continue;
    for (int ii = m0; ii < m1; ++ii) {
      const int imonod = T.i[h.o_mic_mid + ii], jcomp = T.i[h.o_mon_spec + imonod];
      const double act_coef = c.gam[jcomp], activity = c.m[jcomp] * act_coef, Cth = T.d[h.o_mon_Cth + imonod];
      const double dR_dX = Im / monod[ii - m0];
      const double denominator = T.d[h.o_mon_K + imonod] + activity - Cth;
      const double dX_dc = act_coef / denominator - act_coef * (activity - Cth) / (denominator * denominator);
      const double dR_dc = -1.0 * dR_dX * dX_dc;
      for (int p = s0; p < s1; ++p) Jac[T.i[h.o_mic_id + p] + jcomp * n] = Jac[T.i[h.o_mic_id + p] + jcomp * n] + T.d[h.o_mic_st + p] * dR_dc;
      if (ibiomass >= 0) Jac[immobile_id + jcomp * n] = Jac[immobile_id + jcomp * n] + yield * dR_dc;
    }
    for (int ii = i0; ii < i1; ++ii) {
      const int iinh = T.i[h.o_mic_iid + ii], jcomp = T.i[h.o_inh_spec + iinh];
      const double act_coef = c.gam[jcomp], activity = c.m[jcomp] * act_coef, C1 = T.d[h.o_inh_C + iinh];
      const double dR_dX = Im / inhibition[ii - i0];
      double dX_dc;
      switch (T.i[h.o_inh_type + iinh]) {
        case RXN_INHIBITION_MONOD: {
          const double denominator = C1 + activity;
          dX_dc = -1.0 * act_coef * C1 / (denominator * denominator);
        } break;
        case RXN_INHIBITION_INVERSE_MONOD: {
          const double denominator = C1 + activity;
          dX_dc = act_coef / denominator - act_coef * activity / (denominator * denominator);
        } break;
        default: {
          const double C2 = T.d[h.o_inh_C2 + iinh], tempreal = (activity - C1) * C2;
          dX_dc = (C2 * act_coef / (1.0 + tempreal * tempreal)) / PI;
        } break;
      }
      const double dR_dc = -1.0 * dR_dX * dX_dc;
      for (int p = s0; p < s1; ++p) Jac[T.i[h.o_mic_id + p] + jcomp * n] = Jac[T.i[h.o_mic_id + p] + jcomp * n] + T.d[h.o_mic_st + p] * dR_dc;
      if (ibiomass >= 0) Jac[immobile_id + jcomp * n] = Jac[immobile_id + jcomp * n] + yield * dR_dc;
    }
    if (ibiomass >= 0) {
      const double dR_dbiomass = -1.0 * Im / biomass_conc;
      for (int p = s0; p < s1; ++p)
        Jac[T.i[h.o_mic_id + p] + immobile_id * n] = Jac[T.i[h.o_mic_id + p] + immobile_id * n] + T.d[h.o_mic_st + p] * dR_dbiomass;
      Jac[immobile_id + immobile_id * n] = Jac[immobile_id + immobile_id * n] + yield * dR_dbiomass;
    }
  }
}

// RImmobileDecay - reaction_immobile.F90:240-293
template <int N>
__device__ void immobile_decay(const Tab &T, const Cell<N> &c, double *Res, double *Jac, bool compute_derivative) {
  const DevTab &h = *T.h;
  const int n = h.ncomp;
  for (int irxn = 0; irxn < h.nimdecay; ++irxn) {
    const int icomp = T.i[h.o_imdec_id + irxn];
    const double rate_constant = T.d[h.o_imdec_k + irxn] * c.volume;
    const double rate = rate_constant * c.im[icomp];
    const int immobile_id = h.naq + icomp;
    Res[immobile_id] = Res[immobile_id] + rate;
    if (compute_derivative) Jac[immobile_id + immobile_id * n] = Jac[immobile_id + immobile_id * n] + rate_constant;
  }
}

// ---------------------------------------------------------------------------------------------
// load / store of the per-thread part of the state
template <int N>
__device__ void load_cell(const Tab &T, const DevState &S, long long cell, Cell<N> &c) {
  const int naq = T.h->naq;
  c.cell = cell;
  c.flags = 0;
  for (int i = 0; i < naq; ++i) {
    c.m[i] = G(S, RXN_F_PRI_MOLAL, i, cell);
    c.gam[i] = G(S, RXN_F_PRI_ACT_COEF, i, cell);
    c.total[i] = G(S, RXN_F_TOTAL, i, cell);
    c.tsorb[i] = G(S, RXN_F_TOTAL_SORB_EQ, i, cell);
  }
  for (int i = 0; i < T.h->nim; ++i) c.im[i] = G(S, RXN_F_IMMOBILE, i, cell);
  c.ln_act_h2o = G(S, RXN_F_LN_ACT_H2O, 0, cell);
  c.den_kg = G(S, RXN_F_DEN_KG, 0, cell);
  c.sat = G(S, RXN_F_SAT, 0, cell);
  c.temp = G(S, RXN_F_TEMP, 0, cell);
  c.pres = G(S, RXN_F_PRES, 0, cell);
  if (T.h->logK_mode == RXN_LOGK_HPT) hpt_terms(c.temp, c.pres, c.hpt);
  c.volume = G(S, RXN_F_VOLUME, 0, cell);
  c.porosity = G(S, RXN_F_POROSITY, 0, cell);
  c.soil_density = G(S, RXN_F_SOIL_PARTICLE_DENSITY, 0, cell);
}

template <int N>
__device__ void store_cell(const Tab &T, const DevState &S, const Cell<N> &c, const double *dtot, const double *dsorb) {
  const int naq = T.h->naq;
  for (int i = 0; i < naq; ++i) {
    G(S, RXN_F_PRI_MOLAL, i, c.cell) = c.m[i];
    G(S, RXN_F_PRI_ACT_COEF, i, c.cell) = c.gam[i];
    G(S, RXN_F_TOTAL, i, c.cell) = c.total[i];
    G(S, RXN_F_TOTAL_SORB_EQ, i, c.cell) = c.tsorb[i];
  }
  for (int i = 0; i < T.h->nim; ++i) G(S, RXN_F_IMMOBILE, i, c.cell) = c.im[i];
  G(S, RXN_F_LN_ACT_H2O, 0, c.cell) = c.ln_act_h2o;
  if (S.f[RXN_F_DTOTAL] && dtot)
    for (int e = 0; e < naq * naq; ++e) G(S, RXN_F_DTOTAL, e, c.cell) = dtot[e];
  if (S.f[RXN_F_DTOTAL_SORB_EQ] && dsorb && T.h->neqsorb > 0)
    for (int e = 0; e < naq * naq; ++e) G(S, RXN_F_DTOTAL_SORB_EQ, e, c.cell) = dsorb[e];
}

// ---------------------------------------------------------------------------------------------
// RReact — reaction.F90:3322-3511 (control flow: SURVEY.md 3.3).  tran_xx: this cell's AoS row.
// Returns num_iterations; *exit_reason as RXN_EXIT_*.
template <int N>
__device__ int rreact(const Tab &T, const DevState &S, Cell<N> &c, double *tran_xx, double tran_dt, int dt_mode,
                      double *J, double *dsorb, int *exit_reason) {
  const DevTab &h = *T.h;
  const int n = h.ncomp, naq = h.naq, nim = h.nim;       // dofs: naq aqueous species, then nim immobile species
  double residual[N], fixed_accum[N], prev_solution[N], new_solution[N];
  double mrK1[2], mrR0[2 * N];   // nmr <= 2 enforced at table creation
  double dtot[N * N];            // d total / d free-ion of the iterate (RRadioactiveDecay reads it after J has been scaled)
  int num_iterations = 0;
  *exit_reason = 0;
  for (int i = 0; i < naq; ++i) c.total[i] = tran_xx[i];                       // :3370
  for (int i = 0; i < nim; ++i) c.im[i] = tran_xx[naq + i];                    // :3386-3392
  // fixed accumulation: RTAccumulation (:5072-5148) + RAccumulationSorb (:4539-4568)
  const double psv_t = c.porosity * c.sat * 1000.0 * c.volume;
  for (int i = 0; i < naq; ++i) fixed_accum[i] = psv_t * c.total[i];
  if (h.neqsorb > 0) for (int i = 0; i < naq; ++i) fixed_accum[i] = fixed_accum[i] + c.tsorb[i] * c.volume;
  for (int i = 0; i < nim; ++i) fixed_accum[naq + i] = c.im[i] * c.volume;      // :5117-5125
  if (h.nmr > 0) multirate_prepare<N>(T, S, c, tran_dt, mrK1, mrR0);
  if (h.act_freq != RXN_ACT_COEF_FREQUENCY_OFF) activity_coefficients<N>(T, S, c);
  const double psvd_t = c.porosity * c.sat * 1000.0 * c.volume / tran_dt;    // :5189
  const double v_t = c.volume / tran_dt;                                       // :4590
  for (;;) {
    num_iterations = num_iterations + 1;
    if (h.act_freq == RXN_ACT_COEF_FREQUENCY_NEWTON_ITER) activity_coefficients<N>(T, S, c);
    auxvar_compute<N>(T, S, c, J, dsorb);                                      // J <- dtotal
    if (h.ndecay > 0) for (int e = 0; e < naq * naq; ++e) dtot[e] = J[e];
    for (int i = 0; i < naq; ++i) residual[i] = psv_t * c.total[i];
    for (int i = 0; i < nim; ++i) residual[naq + i] = c.im[i] * c.volume;
    for (int i = 0; i < n; ++i) residual[i] = residual[i] - fixed_accum[i];
    for (int e = 0; e < naq * naq; ++e) J[e] = J[e] * psvd_t;                  // RTAccumulationDerivative
    if (h.neqsorb > 0) {
      for (int i = 0; i < naq; ++i) residual[i] = residual[i] + c.tsorb[i] * c.volume;
      for (int e = 0; e < naq * naq; ++e) J[e] = J[e] + dsorb[e] * v_t;
    }
    if (dt_mode == RXN_DT_CONSISTENT)
      for (int i = 0; i < n; ++i) residual[i] = residual[i] / tran_dt;
    // RReaction (:3515-3584): minerals, then multirate
    if (h.nkin > 0) kinetic_mineral<N>(T, S, c, residual, J, true);
    if (h.nmr > 0) multirate_sorption<N>(T, S, c, mrK1, mrR0, residual, J, true, dsorb);
    if (h.nkinrxn > 0) kinetic_surfcplx<N>(T, S, c, tran_dt, residual, J, true);
    if (h.ndecay > 0) radioactive_decay<N>(T, c, dtot, dsorb, residual, J, true);
    if (h.ngen > 0) general_reaction<N>(T, c, residual, J, true);
    if (nim > 0) {                                                             // immobile rows / columns: :5201-5206, then RMicrobial, RImmobileDecay
      expand_block(J, naq, n);
      for (int i = naq; i < n; ++i) J[i + i * n] = c.volume / tran_dt;
    }
    if (h.nmic > 0) microbial_reaction<N>(T, c, residual, J, true);
    if (h.nimdecay > 0) immobile_decay<N>(T, c, residual, J, true);
    double mx = 0.0;
    bool nonfinite = false;
    for (int i = 0; i < n; ++i) { mx = fmax(mx, fabs(residual[i])); if (!isfinite(residual[i])) nonfinite = true; }
    if (nonfinite) { c.flags |= RXN_FLAG_NONFINITE; break; }
    if (mx < h.res_tol) { *exit_reason = RXN_EXIT_RESIDUAL; break; }           // :3443
    for (int i = 0; i < naq; ++i) prev_solution[i] = c.m[i];
    for (int i = 0; i < nim; ++i) prev_solution[naq + i] = c.im[i];            // :3448-3452
    // log formulation with immobile dofs: the reference hands RSolve pri_molal (naqcomp values) as conc(ncomp) (:3445); the
    // immobile concentrations scale their own columns here (include/rxn_b200.h, rxn_react_batch)
    if (rsolve<N>(residual, J, prev_solution, n, h.use_log != 0)) { c.flags |= RXN_FLAG_LU_ZERO_ROW; break; }
    double *update = residual;
    if (h.use_log) {                                                           // :3454-3458
      for (int i = 0; i < n; ++i) update[i] = copysign(1.0, update[i]) * fmin(fabs(update[i]), h.max_dlnC);
      for (int i = 0; i < n; ++i) new_solution[i] = prev_solution[i] * exp(-update[i]);
    } else {                                                                   // :3459-3471
      double min_ratio = 1.0e20;
      for (int i = 0; i < n; ++i) {
        if (prev_solution[i] <= update[i]) {
          const double ratio = fabs(prev_solution[i] / update[i]);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      if (min_ratio < 1.0) for (int i = 0; i < n; ++i) update[i] = update[i] * min_ratio * 0.99;
      for (int i = 0; i < n; ++i) new_solution[i] = prev_solution[i] - update[i];
    }
    double maximum_relative_change = 0.0;
    for (int i = 0; i < n; ++i) {
      const double r = fabs((new_solution[i] - prev_solution[i]) / prev_solution[i]);
      if (!(r <= maximum_relative_change)) maximum_relative_change = r;
    }
    if (!isfinite(maximum_relative_change)) { c.flags |= RXN_FLAG_NONFINITE; break; }
    if (maximum_relative_change < h.rel_tol) { *exit_reason = RXN_EXIT_REL_CHANGE; break; }  // :3476 (update discarded)
    if (num_iterations > 50) {                                                 // :3478-3496
      const double scale = 0.1;
      for (int i = 0; i < n; ++i) new_solution[i] = scale * (new_solution[i] - prev_solution[i]) + prev_solution[i];
    }
    for (int i = 0; i < naq; ++i) c.m[i] = new_solution[i];                    // :3498
    for (int i = 0; i < nim; ++i) c.im[i] = new_solution[naq + i];             // :3499-3502
    if (num_iterations >= h.maxit) { c.flags |= RXN_FLAG_CAPPED; break; }      // GPU-only guard (reference spins)
  }
  auxvar_compute<N>(T, S, c, J, dsorb);                                        // :3507
  for (int i = 0; i < naq; ++i) tran_xx[i] = c.m[i];                           // reactive_transport.F90:1711
  for (int i = 0; i < nim; ++i) tran_xx[naq + i] = c.im[i];                    // the cell's own slots (:1712-1716 drops the cell offset)
  return num_iterations;
}

// =============================================================================================
// per-cell bodies of the batched entry points (one call = one cell of one kernel thread)

// RTReact loop body (reactive_transport.F90:1697-1724)
template <int N>
__device__ void cell_react(const Tab &T, const DevState &S, long long i, double *tran_xx, const int *l2g, double dt,
                           int dt_mode, int *iters, int *flags) {
  const long long cell = l2g ? l2g[i] : i;
  if (S.active && !S.active[cell]) {           // imat <= 0 (reactive_transport.F90:1699)
    if (iters) iters[i] = 0;
    if (flags) flags[i] = RXN_FLAG_INACTIVE;
    return;
  }
  Cell<N> c;
  double J[N * N], dsorb[N * N], xx[N];
  load_cell<N>(T, S, cell, c);
  const int n = T.h->ncomp;
  for (int k = 0; k < n; ++k) xx[k] = tran_xx[i * n + k];
  int reason = 0;
  const int it = rreact<N>(T, S, c, xx, dt, dt_mode, J, dsorb, &reason);
  store_cell<N>(T, S, c, J, dsorb);
  for (int k = 0; k < n; ++k) tran_xx[i * n + k] = xx[k];
  if (iters) iters[i] = it;
  if (flags) flags[i] = reason | c.flags;
}

// cells part of RTUpdateAuxVars (reactive_transport.F90:3790-3846) + RTUpdateActivityCoefficients (:3620-3700)
template <int N>
__device__ void cell_update_auxvars(const Tab &T, const DevState &S, long long cell, const double *xx_loc, int update_act_coefs) {
  if (S.active && !S.active[cell]) return;
  Cell<N> c;
  double dtot[N * N], dsorb[N * N];
  load_cell<N>(T, S, cell, c);
  const int n = T.h->naq, nc = T.h->ncomp;
  if (xx_loc) for (int k = 0; k < n; ++k) c.m[k] = xx_loc[cell * nc + k];
  if (xx_loc) for (int k = n; k < nc; ++k) c.im[k - n] = xx_loc[cell * nc + k];             // :3801-3805
  if (update_act_coefs) activity_coefficients<N>(T, S, c);
  auxvar_compute<N>(T, S, c, dtot, dsorb);
  store_cell<N>(T, S, c, dtot, dsorb);
  for (int k = 0; k < n; ++k) if (!isfinite(c.total[k])) c.flags |= RXN_FLAG_NONFINITE;
  report_cell_flags(S, c.flags);
}

// RTUpdateFixedAccumulation (reactive_transport.F90:786-843)
template <int N>
__device__ void cell_fixed_accum(const Tab &T, const DevState &S, long long i, const double *xx, const int *l2g, double *accum_out) {
  const long long cell = l2g ? l2g[i] : i;
  if (S.active && !S.active[cell]) return;
  Cell<N> c;
  double dtot[N * N], dsorb[N * N];
  load_cell<N>(T, S, cell, c);
  const int n = T.h->naq, nc = T.h->ncomp;
  if (xx) for (int k = 0; k < n; ++k) c.m[k] = xx[i * nc + k];
  if (xx) for (int k = n; k < nc; ++k) c.im[k - n] = xx[i * nc + k];                        // :809-813
  auxvar_compute<N>(T, S, c, dtot, dsorb);
  const double psv_t = c.porosity * c.sat * 1000.0 * c.volume;
  for (int k = 0; k < n; ++k) {
    double r = psv_t * c.total[k];
    if (T.h->neqsorb > 0) r = r + c.tsorb[k] * c.volume;
    accum_out[i * nc + k] = r;
    if (!isfinite(r)) c.flags |= RXN_FLAG_NONFINITE;
  }
  for (int k = n; k < nc; ++k) accum_out[i * nc + k] = c.im[k - n] * c.volume;              // reaction.F90:5117-5125
  store_cell<N>(T, S, c, dtot, dsorb);
  report_cell_flags(S, c.flags);
}

// accumulation + reaction loops of RTResidualNonFlux (reactive_transport.F90:2545-2586, 2735-2758)
// and RTJacobianNonFlux (:3342-3389, 3445-3465).
// res_out[i] = (RTAccumulation + RAccumulationSorb)/dt + RReaction ; jac_out[i] = accumulation
// derivative block + reaction derivative block (each accumulated from zero, then added, as the
// two MatSetValuesBlockedLocal(ADD_VALUES) calls of the reference do).
// total / dtotal are re-evaluated from pri_molal and the current activity coefficients - the
// values RTUpdateAuxVars produced for this Newton iterate - so no naq^2 block is kept in HBM.
template <int N>
__device__ void cell_residual_jacobian(const Tab &T, const DevState &S, long long i, const int *l2g, double dt,
                                       double *res_out, double *jac_out) {
  const long long cell = l2g ? l2g[i] : i;
  if (S.active && !S.active[cell]) return;
  const DevTab &tab = *T.h;
  Cell<N> c;
  double J[N * N], dsorb[N * N], J2[N * N], Res[N], Res2[N];
  double mrK1[2], mrR0[2 * N];
  load_cell<N>(T, S, cell, c);
  const int n = tab.naq;
  auxvar_compute<N>(T, S, c, J, dsorb);
  if (tab.ndecay > 0) for (int e = 0; e < n * n; ++e) J2[e] = J[e];          // d total / d free-ion for RRadioactiveDecay
  const double psv_t = c.porosity * c.sat * 1000.0 * c.volume;
  const double psvd_t = c.porosity * c.sat * 1000.0 * c.volume / dt;
  const double v_t = c.volume / dt;
  for (int k = 0; k < n; ++k) Res[k] = psv_t * c.total[k];
  for (int e = 0; e < n * n; ++e) J[e] = J[e] * psvd_t;
  if (tab.neqsorb > 0) {
    for (int k = 0; k < n; ++k) Res[k] = Res[k] + c.tsorb[k] * c.volume;
    for (int e = 0; e < n * n; ++e) J[e] = J[e] + dsorb[e] * v_t;
  }
  const int nc = tab.ncomp;
  for (int k = n; k < nc; ++k) Res[k] = c.im[k - n] * c.volume;               // reaction.F90:5117-5125
  for (int k = 0; k < nc; ++k) { Res[k] = Res[k] / dt; Res2[k] = 0.0; }
  const bool deriv = jac_out != nullptr;
  if (tab.ndecay > 0) {
    // the decay terms read d total / d free-ion (kept in J2 so far): add them into the accumulation block first -
    // REASSOC: the reference adds the two blocks entry by entry in MatSetValuesBlockedLocal(ADD_VALUES); the sum of an
    // entry's terms is taken in the order accumulation, decay, minerals, ... instead of accumulation + (minerals + ... + decay)
    radioactive_decay<N>(T, c, J2, dsorb, Res2, J, deriv);
  }
  for (int e = 0; e < n * n; ++e) J2[e] = 0.0;
  if (tab.nkin > 0) kinetic_mineral<N>(T, S, c, Res2, J2, deriv);
  if (tab.nmr > 0) {
    multirate_prepare<N>(T, S, c, dt, mrK1, mrR0);
    multirate_sorption<N>(T, S, c, mrK1, mrR0, Res2, J2, deriv, dsorb);
  }
  if (tab.nkinrxn > 0) kinetic_surfcplx<N>(T, S, c, dt, Res2, J2, deriv);
  if (tab.ngen > 0) general_reaction<N>(T, c, Res2, J2, deriv);
  if (nc > n) {                                                                // immobile rows / columns (see expand_block)
    expand_block(J, n, nc);
    expand_block(J2, n, nc);
    for (int k = n; k < nc; ++k) J[k + k * nc] = c.volume / dt;               // reaction.F90:5201-5206
  }
  if (tab.nmic > 0) microbial_reaction<N>(T, c, Res2, J2, deriv);
  if (tab.nimdecay > 0) immobile_decay<N>(T, c, Res2, J2, deriv);
  for (int k = 0; k < nc; ++k) if (!isfinite(Res[k] + Res2[k])) c.flags |= RXN_FLAG_NONFINITE;
  if (res_out) for (int k = 0; k < nc; ++k) res_out[i * nc + k] = Res[k] + Res2[k];
  if (jac_out) for (int e = 0; e < nc * nc; ++e) jac_out[i * (long long)(nc * nc) + e] = J[e] + J2[e];
  store_cell<N>(T, S, c, nullptr, nullptr);   // totals + warm-start free sites stay consistent
  report_cell_flags(S, c.flags);
}

// RTUpdateKineticState loop (reactive_transport.F90:692-705) = RUpdateKineticState (reaction.F90:5320-5429)
// skip_mr: the multirate sorbed totals are advanced by the streaming kernel k_kinmr_update (rxn_b200.cu) instead
template <int N>
__device__ void cell_update_kinetic_state(const Tab &T, const DevState &S, long long cell, double dt, bool skip_mr = false) {
  if (S.active && !S.active[cell]) return;
  const DevTab &tab = *T.h;
  Cell<N> c;
  load_cell<N>(T, S, cell, c);
  const int n = tab.naq;
  if (tab.nkin > 0) {
    double res[N];
    for (int k = 0; k < n; ++k) res[k] = 0.0;
    compute_ln<N>(T, c);
    kinetic_mineral<N>(T, S, c, res, nullptr, false);
    for (int im = 0; im < tab.nkin; ++im) {                                   // :5354-5364
      const double delta_volfrac = G(S, RXN_F_MNRL_RATE, im, cell) * T.d[tab.o_k_molar_vol + im] * dt;
      double vf = G(S, RXN_F_MNRL_VOLFRAC, im, cell) + delta_volfrac;
      if (vf < 0.0) vf = 0.0;
      G(S, RXN_F_MNRL_VOLFRAC, im, cell) = vf;
    }
  }
  for (int ikr = 0; ikr < (skip_mr ? 0 : tab.nmr); ++ikr) {                    // :5394-5408
    const int nrate = T.i[tab.o_mr_nrate + ikr];
    const long long blk = (long long)ikr * (tab.mr_ld + 1) * n;
    for (int irate = 0; irate < nrate; ++irate) {
      const double rate = T.d[tab.o_mr_rate + ikr * tab.mr_ld + irate], frac = T.d[tab.o_mr_frac + ikr * tab.mr_ld + irate];
      const double kdt = rate * dt;
      const double one_plus_kdt = 1.0 + kdt;
      for (int k = 0; k < n; ++k) {
        double &Sr = G(S, RXN_F_KINMR_TOTAL_SORB, blk + (long long)(irate + 1) * n + k, cell);
        const double S0 = G(S, RXN_F_KINMR_TOTAL_SORB, blk + k, cell);
        Sr = (Sr + kdt * frac * S0) / one_plus_kdt;
      }
    }
  }
  if (tab.nkinrxn > 0) {                                                     // :5411-5419
    for (int j = T.i[tab.o_rxn_cptr + tab.kin_rxn]; j < T.i[tab.o_rxn_cptr + tab.kin_rxn + 1]; ++j) {
      const int icplx = T.i[tab.o_rxn_cid + j];
      G(S, RXN_F_KINSRFCPLX_CONC, icplx, cell) = G(S, RXN_F_KINSRFCPLX_CONC_KP1, icplx, cell);
    }
  }
  for (int im = 0; im < tab.nkin; ++im) if (!isfinite(G(S, RXN_F_MNRL_VOLFRAC, im, cell))) c.flags |= RXN_FLAG_NONFINITE;
  report_cell_flags(S, c.flags);
}

// RTotalSorbMultiRateAsEQ — reaction_surf_complex.F90:506-562
template <int N>
__device__ void rtotal_sorb_multirate_as_eq(const Tab &T, const DevState &S, Cell<N> &c, double *scratch_dsorb) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  double total_sorb_eq[N];
  for (int ikr = 0; ikr < h.nmr; ++ikr) {
    const int irxn = T.i[h.o_mr_rxn + ikr];
    for (int i = 0; i < naq; ++i) total_sorb_eq[i] = 0.0;
    for (int e = 0; e < naq * naq; ++e) scratch_dsorb[e] = 0.0;
    sorb_eq_surfcplx1<N>(T, S, c, irxn, false, total_sorb_eq, scratch_dsorb);
    const long long row0 = (long long)ikr * (h.mr_ld + 1) * naq;
    for (int i = 0; i < naq; ++i) G(S, RXN_F_KINMR_TOTAL_SORB, row0 + i, c.cell) = total_sorb_eq[i];
  }
}

// ln(Q/K) of a mineral / gas of the constraint lists at the stored logK (reaction.F90:1730-1790)
template <int N>
__device__ double constraint_lnQK(const Tab &T, const DSpec &sp, int r, const Cell<N> &c, double *Jrow, int naq, bool fill) {
  double lnQK = -T.d[sp.o_logK + r] * RXN_LOG_TO_LN;
  const double h2ost = T.d[sp.o_h2ost + r];
  if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
  for (int p = T.i[sp.o_ptr + r]; p < T.i[sp.o_ptr + r + 1]; ++p) {
    const int j = T.i[sp.o_id + p];
    lnQK = lnQK + T.d[sp.o_st + p] * log(c.m[j] * c.gam[j]);
  }
  if (fill)
    for (int p = T.i[sp.o_ptr + r]; p < T.i[sp.o_ptr + r + 1]; ++p) {
      const int j = T.i[sp.o_id + p];
      Jrow[(size_t)j * naq] = T.d[sp.o_st + p] / c.m[j];
    }
  return lnQK;
}

// ReactionEquilibrateConstraint — reaction.F90:1308-2046, one cell.  Returns RXN_EQ_*.
template <int N>
__device__ int cell_equilibrate(const Tab &T, const DevState &S, long long cell, const int *constraint_type, const double *conc_in,
                                const int *constraint_id, const double *free_ion_guess, int use_prev, int init_molal,
                                double *basis_molarity, int *num_iterations_out) {
  const DevTab &h = *T.h;
  const int naq = h.naq;
  const double *Z = T.d + h.o_Z;
  Cell<N> c;
  load_cell<N>(T, S, cell, c);
  double conc[N], Res[N], total_conc[N], free_conc[N], prev_molal[N];
  double Jac[N * N], dsorb[N * N], dtot[N * N];
  double convert_molal_to_molar, convert_molar_to_molal;
  const double xmass = 1.0;
  if (init_molal) { convert_molal_to_molar = c.den_kg * xmass / 1000.0; convert_molar_to_molal = 1.0; }
  else { convert_molal_to_molar = 1.0; convert_molar_to_molal = 1000.0 / c.den_kg / xmass; }
  for (int i = 0; i < naq; ++i) {
    conc[i] = conc_in[i];
    total_conc[i] = 0.0;
    free_conc[i] = use_prev ? c.m[i] : (free_ion_guess ? free_ion_guess[i] : 1.0e-9);
  }
  for (int i = 0; i < naq; ++i) {                                     // :1409-1470
    switch (constraint_type[i]) {
      case RXN_CONSTRAINT_NULL: case RXN_CONSTRAINT_TOTAL: total_conc[i] = conc[i] * convert_molal_to_molar; break;
      case RXN_CONSTRAINT_TOTAL_SORB: total_conc[i] = conc[i]; break;
      case RXN_CONSTRAINT_FREE: free_conc[i] = conc[i] * convert_molar_to_molal; break;
      case RXN_CONSTRAINT_LOG: free_conc[i] = pow(10.0, conc[i]) * convert_molar_to_molal; break;
      case RXN_CONSTRAINT_CHARGE_BAL: if (!use_prev) free_conc[i] = conc[i] * convert_molar_to_molal; break;
      case RXN_CONSTRAINT_PH:
        if (h.h_ion_id == 0) return RXN_EQ_NO_H_ION;
        free_conc[i] = pow(10.0, -conc[i]);
        break;
      case RXN_CONSTRAINT_MINERAL: if (!use_prev) free_conc[i] = conc[i] * convert_molar_to_molal; break;
      case RXN_CONSTRAINT_GAS: if (conc[i] <= 0.0) conc[i] = pow(10.0, conc[i]); break;
      default: return RXN_EQ_BAD_CONSTRAINT;
    }
  }
  for (int i = 0; i < naq; ++i) c.m[i] = free_conc[i];
  int num_iterations = 0, num_it_act_coef_turned_on = 0;
  bool compute_activity_coefs = use_prev != 0;
  bool charge_balance_warning_flag = false;
  for (;;) {                                                            // :1553-1990
    for (int i = 0; i < naq; ++i)
      if (constraint_type[i] == RXN_CONSTRAINT_FREE || constraint_type[i] == RXN_CONSTRAINT_LOG) c.m[i] = free_conc[i];
    if (h.act_freq != RXN_ACT_COEF_FREQUENCY_OFF && compute_activity_coefs) activity_coefficients<N>(T, S, c);
    compute_ln<N>(T, c);
    rtotal<N>(T, S, c, dtot);
    if (h.neqsorb + h.nmr > 0) {
      if (h.neqsorb > 0) rtotal_sorb<N>(T, S, c, dsorb);
      if (h.nmr > 0) rtotal_sorb_multirate_as_eq<N>(T, S, c, Jac);      // Jac is free here: scratch
    }
    for (int e = 0; e < naq * naq; ++e) Jac[e] = 0.0;
#define JAC(i, j) Jac[(i) + (size_t)(j) * naq]
    for (int icomp = 0; icomp < naq; ++icomp) {
      switch (constraint_type[icomp]) {
        case RXN_CONSTRAINT_NULL: case RXN_CONSTRAINT_TOTAL:
          Res[icomp] = c.total[icomp] - total_conc[icomp];
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = dtot[icomp + (size_t)j * naq];
          break;
        case RXN_CONSTRAINT_TOTAL_SORB:
          Res[icomp] = c.tsorb[icomp] - total_conc[icomp];
          for (int j = 0; j < naq; ++j) JAC(icomp, j) = dsorb[icomp + (size_t)j * naq];
          break;
        case RXN_CONSTRAINT_FREE: case RXN_CONSTRAINT_LOG:
          Res[icomp] = 0.0;
          JAC(icomp, icomp) = 1.0;
          break;
        case RXN_CONSTRAINT_CHARGE_BAL:
          Res[icomp] = 0.0;
          for (int jcomp = 0; jcomp < naq; ++jcomp) {
            Res[icomp] = Res[icomp] + Z[jcomp] * c.total[jcomp];
            for (int kcomp = 0; kcomp < naq; ++kcomp)
              JAC(icomp, jcomp) = JAC(icomp, jcomp) + Z[kcomp] * dtot[kcomp + (size_t)jcomp * naq];
          }
          if (c.m[icomp] < 1.0e-20 && !charge_balance_warning_flag) {
            if ((Res[icomp] > 0.0 && Z[icomp] > 0.0) || (Res[icomp] < 0.0 && Z[icomp] < 0.0)) {
              charge_balance_warning_flag = true;
              c.m[icomp] = (double)1.e-3f;                              // the reference literal is single precision
            }
          }
          break;
        case RXN_CONSTRAINT_PH:
          Res[icomp] = 0.0;
          if (h.h_ion_id > 0) {
            c.m[icomp] = pow(10.0, -conc[icomp]) / c.gam[icomp];
            JAC(icomp, icomp) = 1.0;
          } else {
            const int icplx = -h.h_ion_id - 1;
            double lnQK = -logK_of(T, h.cplx, icplx, c, false) * RXN_LOG_TO_LN;
            const double h2ost = T.d[h.cplx.o_h2ost + icplx];
            if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
            for (int p = T.i[h.cplx.o_ptr + icplx]; p < T.i[h.cplx.o_ptr + icplx + 1]; ++p) {
              const int j = T.i[h.cplx.o_id + p];
              lnQK = lnQK + T.d[h.cplx.o_st + p] * log(c.m[j] * c.gam[j]);
            }
            lnQK = lnQK + conc[icomp] * RXN_LOG_TO_LN;
            const double QK = exp(lnQK);
            Res[icomp] = 1.0 - QK;
            for (int p = T.i[h.cplx.o_ptr + icplx]; p < T.i[h.cplx.o_ptr + icplx + 1]; ++p) {
              const int j = T.i[h.cplx.o_id + p];
              JAC(icomp, j) = -QK / c.m[j] * T.d[h.cplx.o_st + p];
            }
          }
          break;
        case RXN_CONSTRAINT_MINERAL:
          Res[icomp] = constraint_lnQK<N>(T, h.mnrl, constraint_id[icomp] - 1, c, &JAC(icomp, 0), naq, true);
          break;
        case RXN_CONSTRAINT_GAS:
          Res[icomp] = constraint_lnQK<N>(T, h.gas, constraint_id[icomp] - 1, c, &JAC(icomp, 0), naq, true) - log(conc[icomp]);
          break;
      }
    }
#undef JAC
    double maximum_residual = 0.0;
    for (int i = 0; i < naq; ++i) maximum_residual = fmax(maximum_residual, fabs(Res[i]));
    bool use_log_formulation;
    if (h.use_log) {                                                    // :1900-1912
      if (num_iterations > 3 && num_iterations < 9) use_log_formulation = (num_iterations % 2 == 0);
      else use_log_formulation = true;
    } else {
      use_log_formulation = false;
    }
    if (rsolve<N>(Res, Jac, c.m, naq, use_log_formulation)) return RXN_EQ_LU_ZERO_ROW;
    double *update = Res;
    for (int i = 0; i < naq; ++i) prev_molal[i] = c.m[i];
    if (use_log_formulation) {
      for (int i = 0; i < naq; ++i) update[i] = copysign(1.0, update[i]) * fmin(fabs(update[i]), h.max_dlnC);
      for (int i = 0; i < naq; ++i) c.m[i] = c.m[i] * exp(-update[i]);
    } else {
      double min_ratio = 1.0e20;
      for (int i = 0; i < naq; ++i) {
        if (prev_molal[i] <= update[i]) {
          const double ratio = fabs(prev_molal[i] / update[i]);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      if (min_ratio <= 1.0) for (int i = 0; i < naq; ++i) update[i] = update[i] * min_ratio * 0.99;
      for (int i = 0; i < naq; ++i) c.m[i] = prev_molal[i] - update[i];
    }
    double mn = c.m[0];
    for (int i = 1; i < naq; ++i) mn = fmin(mn, c.m[i]);
    if (!(mn > 0.0)) return RXN_EQ_ZERO_CONCENTRATION;                  // "Zero concentrations found in constraint"
    double maximum_relative_change = 0.0;
    for (int i = 0; i < naq; ++i) maximum_relative_change = fmax(maximum_relative_change, fabs((c.m[i] - prev_molal[i]) / prev_molal[i]));
    num_iterations = num_iterations + 1;
    if (num_iterations >= 10000) return RXN_EQ_NOT_CONVERGED;
    if (maximum_residual < h.res_tol && maximum_relative_change < h.rel_tol) {
      if (compute_activity_coefs && num_iterations - num_it_act_coef_turned_on > 1) break;
      if (!compute_activity_coefs) num_it_act_coef_turned_on = num_iterations;
      compute_activity_coefs = true;
    }
  }
  if (h.neqsorb + h.nmr > 0) {                                          // :1995-2012
    if (h.neqsorb > 0) rtotal_sorb<N>(T, S, c, dsorb);
    if (h.nmr > 0) rtotal_sorb_multirate_as_eq<N>(T, S, c, Jac);
  }
  for (int ikr = 0; ikr < h.nmr; ++ikr) {                               // :2014-2026
    const long long blk = (long long)ikr * (h.mr_ld + 1) * naq;
    const int nrate = T.i[h.o_mr_nrate + ikr];
    for (int irate = 0; irate < nrate; ++irate) {
      const double frac = T.d[h.o_mr_frac + ikr * h.mr_ld + irate];
      for (int i = 0; i < naq; ++i)
        G(S, RXN_F_KINMR_TOTAL_SORB, blk + (long long)(irate + 1) * naq + i, cell) = frac * G(S, RXN_F_KINMR_TOTAL_SORB, blk + i, cell);
    }
  }
  if (basis_molarity) for (int i = 0; i < naq; ++i) basis_molarity[i] = c.m[i] * c.den_kg / 1000.0;
  *num_iterations_out = num_iterations;
  store_cell<N>(T, S, c, nullptr, nullptr);
  return RXN_EQ_OK;
}

}  // namespace rxn
