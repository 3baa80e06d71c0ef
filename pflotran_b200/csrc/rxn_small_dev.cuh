// rxn_small_dev.cuh — device code of the register RReact kernel for small chemistries (design: rxn_small.h).
//
// Reference routines restated (file:line at each site): RReact (reaction.F90:3322-3511), RActivityCoefficients LAG
// (:3994-4050), RTotal (:4057-4158), RKineticMineral (reaction_mineral.F90:564-1000), RSolve (reaction.F90:4835-4880) with
// ludcmp / lubksb (utility.F90:393-523), RUpdateTempDependentCoefs (reaction.F90:5433-5524; fits reaction_aux.F90:1461-1488,
// :1529-1571).  Deviations from the reference's operation order are those of the resident-lane kernel (REASSOC, <= 1e-14):
// ln-m Jacobian, exp(lnQK - ln gamma), sums over species in ascending order.
//
// The same source is compiled for the host by the CPU-only test harness (RXN_SMALL_HOST) and checked against the oracle.
#pragma once
#include "rxn_small.h"

#ifndef RXN_SMALL_HOST
#include <cuda_runtime.h>
#define SM_DEV static __device__ __forceinline__
#define SM_COLD static __device__ __noinline__
#else
#define SM_DEV static inline
#define SM_COLD static inline
#endif

namespace rxn {
namespace small {

#ifndef RXN_LOG_TO_LN
#define RXN_LOG_TO_LN 2.30258509299           /* pflotran_constants.F90:48 (truncated on purpose) */
#define RXN_IDEAL_GAS_CONSTANT 8.31446        /* pflotran_constants.F90:53 */
#endif
#define SGS(S, field, row, cell) ((S).f[field][(long long)(row) * (S).ld + (cell)])

SM_DEV double s_pow(double x, double y) { return y == 1.0 ? x : pow(x, y); }   // pow(x, 1) == x exactly
// x / d with r = RN(1/d): the correctly rounded quotient (Markstein), 3 FMA-class instructions
SM_DEV double s_div(double x, double d, double r) {
#ifndef RXN_SMALL_HOST
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
#else
  (void)r;
  return x / d;
#endif
}

// -logK * LOG_TO_LN of one reaction at the cell's T (and P); hp = {tr, pr, log10 tr, sqrt tr, 1/tr, 1/pr}
SM_DEV double s_nlk(const double *cf, int logK_mode, double tk, const double *hp) {
  double lk;
  if (logK_mode == RXN_LOGK_HPT) {                             // reaction_aux.F90:1529-1571 (divisions: s_div, the same quotients)
    const double tr = hp[0], pr = hp[1], logtr = hp[2], sqtr = hp[3], itr = hp[4], ipr = hp[5];
    lk = cf[0] + cf[1] * tr + s_div(cf[2], tr, itr) + cf[3] * logtr + cf[4] * tr * tr + s_div(s_div(cf[5], tr, itr), tr, itr) + cf[6] * sqtr +
         cf[7] * pr + cf[8] * pr * tr + s_div(cf[9] * pr, tr, itr) + cf[10] * pr * logtr + s_div(cf[11], pr, ipr) + s_div(cf[12], pr, ipr) * tr +
         s_div(s_div(cf[13], pr, ipr), tr, itr) + cf[14] * pr * pr + cf[15] * pr * pr * tr + s_div(cf[16] * pr * pr, tr, itr);
  } else {                                                     // reaction_aux.F90:1461-1488
    lk = cf[0] * log(tk) + cf[1] + cf[2] * tk + cf[3] / tk + cf[4] / (tk * tk);
  }
  return -lk * RXN_LOG_TO_LN;
}

// One cell, start to finish.  N = SMALL_N.
template <int N>
SM_DEV void small_react_cell(const SmallTab &T, const DevState &S, long long item, long long cell, double *tran_xx, double tran_dt,
                             int dt_mode, int32_t *iters, int32_t *flags) {
  constexpr int MC = SMALL_MAXC, MK = SMALL_MAXK;
  const int n = T.n, ncplx = T.ncplx, nkin = T.nkin;
  // ---- load (RTAuxVarInit values + flow coupling scalars)
  double ln_act_h2o = SGS(S, RXN_F_LN_ACT_H2O, 0, cell);
  const double den_kg = SGS(S, RXN_F_DEN_KG, 0, cell), temp = SGS(S, RXN_F_TEMP, 0, cell), volume = SGS(S, RXN_F_VOLUME, 0, cell);
  const double porosity = SGS(S, RXN_F_POROSITY, 0, cell), sat = SGS(S, RXN_F_SAT, 0, cell);
  const double psv = porosity * sat * 1000.0 * volume;
  const double psvd = porosity * sat * 1000.0 * volume / tran_dt;             // :5189
  const double den_kg_per_L = den_kg * 1.0 * 1.0e-3;
  const double inv_dt = 1.0 / tran_dt;
  double m[N], fix[N], lngp[N], lngc[MC], sm[MC], nlkc[MC], nlkk[MK], volfrac[MK], area[MK], rate_out[MK];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < n) {
      m[i] = SGS(S, RXN_F_PRI_MOLAL, i, cell);
      fix[i] = psv * tran_xx[item * n + i];                    // :3370, RTAccumulation :5072-5148
      lngp[i] = T.act_off ? log(SGS(S, RXN_F_PRI_ACT_COEF, i, cell)) : 0.0;
    } else {                                                   // padding row: m = 1, no complexes -> residual 0, decoupled
      m[i] = 1.0; fix[i] = psv * ((1.0 + 0.0) * den_kg_per_L); lngp[i] = 0.0;
    }
  }
#pragma unroll
  for (int k = 0; k < MC; ++k) {
    sm[k] = 0.0; lngc[k] = 0.0; nlkc[k] = 0.0;
    if (k < ncplx) {
      sm[k] = SGS(S, RXN_F_SEC_MOLAL, k, cell);                // lagged, for the ionic strength
      if (T.act_off) lngc[k] = log(SGS(S, RXN_F_SEC_ACT_COEF, k, cell));
    }
  }
#pragma unroll
  for (int q = 0; q < MK; ++q) {
    volfrac[q] = 0.0; area[q] = 0.0; rate_out[q] = 0.0; nlkk[q] = 0.0;
    if (q < nkin) { volfrac[q] = SGS(S, RXN_F_MNRL_VOLFRAC, q, cell); area[q] = SGS(S, RXN_F_MNRL_AREA, q, cell); }
  }
  {                                                            // RUpdateTempDependentCoefs :5433-5524
    double hp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double tk = temp + 273.15;
    if (T.logK_mode == RXN_LOGK_HPT) {
      const double pres = SGS(S, RXN_F_PRES, 0, cell);
      hp[0] = tk / 273.15; hp[1] = pres / 1.0e7; hp[2] = log(hp[0]) / log(10.0); hp[3] = sqrt(hp[0]); hp[4] = 1.0 / hp[0]; hp[5] = 1.0 / hp[1];
    }
#pragma unroll
    for (int k = 0; k < MC; ++k)
      if (k < ncplx) nlkc[k] = T.cplx_fit[k] ? s_nlk(T.ccoef[k], T.logK_mode, tk, hp) : T.cnlk[k];
#pragma unroll
    for (int q = 0; q < MK; ++q)
      if (q < nkin) nlkk[q] = T.kin_fit[q] ? s_nlk(T.kcoef[q], T.logK_mode, tk, hp) : T.knlk[q];
  }
  const bool consistent = dt_mode == RXN_DT_CONSISTENT;
  int iter = 0, status = 0, cflags = 0;
  bool closing = false;
  double tot[N], lna[N];
#pragma unroll 1
  for (;;) {
    if (!closing) {
      iter = iter + 1;
      // RActivityCoefficients, LAG :3994-4050 (:3407-3409 before the loop and :3413-3418 in iteration 1 see the same inputs)
      if (!T.act_off && (iter == 1 || T.act_newton_iter)) {
        double I = 0.0, psum = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i)
          if (i < n) { I = fma(m[i], T.pz2[i], I); if (T.use_act_h2o && i + 1 != T.h2o_aq_id) psum += m[i]; }
#pragma unroll
        for (int k = 0; k < MC; ++k)
          if (k < ncplx) { I = fma(sm[k], T.cz2[k], I); if (T.use_act_h2o) psum += sm[k]; }
        I = 0.5 * I;
        const double sqrt_I = sqrt(I);
#pragma unroll
        for (int i = 0; i < N; ++i)
          lngp[i] = (i < n && T.pcharged[i]) ? (-T.pz2[i] * sqrt_I * T.debyeA / (1.0 + T.pa0[i] * T.debyeB * sqrt_I) + T.debyeBdot * I) * RXN_LOG_TO_LN : 0.0;
#pragma unroll
        for (int k = 0; k < MC; ++k)
          lngc[k] = (k < ncplx && T.ccharged[k]) ? (-T.cz2[k] * sqrt_I * T.debyeA / (1.0 + T.ca0[k] * T.debyeB * sqrt_I) + T.debyeBdot * I) * RXN_LOG_TO_LN : 0.0;
        if (T.use_act_h2o) {                                   // :4043-4050
          const double a = 1.0 - 0.017 * psum;
          ln_act_h2o = (a > 0.0) ? log(a) : 0.0;
        }
      }
    }
    // RTotal :4057-4158: ln a_i, sec_molal_k = exp(lnQK_k - ln gamma_k), total_i
#pragma unroll
    for (int i = 0; i < N; ++i) lna[i] = log(m[i]) + lngp[i];
#pragma unroll
    for (int i = 0; i < N; ++i) tot[i] = m[i];
#pragma unroll
    for (int k = 0; k < MC; ++k)
      if (k < ncplx) {
        double lnQK = nlkc[k];
        if (T.ch2o[k] != 0.0) lnQK = lnQK + T.ch2o[k] * ln_act_h2o;
#pragma unroll
        for (int j = 0; j < N; ++j) lnQK = fma(T.nu[k][j], lna[j], lnQK);
        sm[k] = exp(lnQK - lngc[k]);
#pragma unroll
        for (int i = 0; i < N; ++i) tot[i] = fma(T.nu[k][i], sm[k], tot[i]);
      }
#pragma unroll
    for (int i = 0; i < N; ++i) tot[i] = tot[i] * den_kg_per_L;                 // :4095, 4124, 4148
    if (closing) break;
    // residual (:3424-3426) and ln-m Jacobian: Jln_ij = (sum_k nu_ki nu_kj sm_k) dp, diagonal + m_i dp (RTAccumulationDerivative :5189-5204)
    const double dp = den_kg_per_L * psvd;
    double J[N][N], b[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < MC; ++k)
          if (k < ncplx) a = fma(T.nu[k][i] * T.nu[k][j], sm[k], a);
        a = a * dp;
        if (i == j) a = fma(m[i], dp, a);
        J[i][j] = a;
      }
      double res = psv * tot[i];
      res = res - fix[i];
      if (consistent) res = s_div(res, tran_dt, inv_dt);
      b[i] = res;
    }
    // RReaction :3440 -> RKineticMineral, reaction_mineral.F90:564-1000
#pragma unroll
    for (int q = 0; q < MK; ++q)
      if (q < nkin) {
        double lnQK = nlkk[q];
        if (T.kh2o[q] != 0.0) lnQK = lnQK + T.kh2o[q] * ln_act_h2o;
#pragma unroll
        for (int j = 0; j < N; ++j) lnQK = fma(T.nuk[q][j], lna[j], lnQK);
        double QK;
        if (lnQK <= 6.90776) QK = exp(lnQK); else QK = 1.0e3;
        const double k_scale = T.k_scale[q], k_Temkin = T.k_Temkin[q], k_power = T.k_power[q], k_lim = T.k_lim[q], k_aff = T.k_aff[q];
        double affinity_factor;
        if (T.has_Temkin) {
          if (T.has_scale) affinity_factor = 1.0 - s_pow(QK, 1.0 / (k_scale * k_Temkin));
          else affinity_factor = 1.0 - s_pow(QK, 1.0 / k_Temkin);
        } else if (T.has_scale) {
          affinity_factor = 1.0 - s_pow(QK, 1.0 / k_scale);
        } else {
          affinity_factor = 1.0 - QK;
        }
        const double sign_ = copysign(1.0, affinity_factor);
        bool act = (volfrac[q] > 0 || sign_ < 0.0);              // :723
        if (k_aff > 0.0 && sign_ < 0.0 && QK < k_aff) act = false;   // :730
        if (k_lim > 0.0) affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / k_lim);
        double arrhenius_factor = 1.0;
        if (T.k_Ea[q] > 0.0) arrhenius_factor = exp(T.k_Ea[q] / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));
        const double sum_prefactor_rate = T.k_rate[q] * arrhenius_factor;
        double Im_const = -area[q], Im;
        if (T.has_scale) Im_const = Im_const / k_scale;
        if (T.has_power) Im = Im_const * sign_ * s_pow(fabs(affinity_factor), k_power) * sum_prefactor_rate;
        else Im = Im_const * sign_ * fabs(affinity_factor) * sum_prefactor_rate;
        rate_out[q] = act ? Im : 0.0;                           // :575 (zeroed) / :816
        Im_const = Im_const * volume;
        Im = Im * volume;
        double dIm_dQK;
        if (T.has_power) dIm_dQK = -Im * k_power / fabs(affinity_factor);
        else dIm_dQK = -Im_const * sum_prefactor_rate;
        if (T.has_Temkin) {
          if (T.has_scale) dIm_dQK = dIm_dQK * (1.0 / (k_scale * k_Temkin)) / QK * (1.0 - affinity_factor);
          else dIm_dQK = dIm_dQK * (1.0 / k_Temkin) / QK * (1.0 - affinity_factor);
        } else if (T.has_scale) {
          dIm_dQK = dIm_dQK * (1.0 / k_scale) / QK * (1.0 - affinity_factor);
        }
        const double den = (k_lim <= 0.0) ? 1.0 : 1.0 + (1.0 - affinity_factor) / k_lim;
        if (!act) { Im = 0.0; dIm_dQK = 0.0; }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double stp = T.nuk[q][i];
          if (stp != 0.0) {                                    // the species of the mineral (warp uniform: a table value)
            b[i] = b[i] + stp * Im;
#pragma unroll
            for (int j = 0; j < N; ++j) {
              if (T.nuk[q][j] != 0.0) {
                const double dQK_dCj = T.nuk[q][j] * QK;        // d/d ln m_j: the reference's exp(-ln m_j) factor is not applied
                const double dQK_dmj = dQK_dCj * den_kg * 1.0e-3;
                double add;
                if (k_lim <= 0.0) add = stp * dIm_dQK * dQK_dmj;
                else add = stp * dIm_dQK * (1.0 + QK / k_lim / den) * dQK_dmj / den;
                J[i][j] = J[i][j] + (act ? add : 0.0);
              }
            }
          }
        }
      }
    // convergence on the residual (:3443)
    double mx = 0.0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < N; ++i) { mx = fmax(mx, fabs(b[i])); if (!isfinite(b[i])) bad = true; }
    if (bad) { status = RXN_FLAG_NONFINITE; closing = true; continue; }
    if (mx < T.res_tol) { status = RXN_EXIT_RESIDUAL; break; }
    // RSolve :4835-4880: rows scaled by 1/max(1, max_j |J_ij|), J_ij = Jln_ij / m_j; log form: times m_j (:4866-4870)
    double vv[N];
    bool zero = false;
    {
      double invm[N];
#pragma unroll
      for (int j = 0; j < N; ++j) invm[j] = 1.0 / m[j];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double rmx = 0.0, mraw = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double av = fabs(J[i][j]), v = av * invm[j];
          if (v > rmx) rmx = v;
          if (av > mraw) mraw = av;
        }
        const double norm = 1.0 / ((rmx > 1.0) ? rmx : 1.0);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double v = J[i][j];
          if (!T.use_log) v = v * invm[j];
          J[i][j] = v * norm;
        }
        b[i] = b[i] * norm;
        const double aamax = T.use_log ? mraw * norm : rmx * norm;   // ludcmp :413-425
        if (aamax <= 0.0) zero = true;
        vv[i] = 1.0 / aamax;
      }
    }
    if (zero) { status = RXN_FLAG_LU_ZERO_ROW; closing = true; continue; }
    // ludcmp / lubksb (utility.F90:393-523) as a right-looking elimination with b carried along: per element the same
    // a(i,j) -= a(i,k) a(k,j), k ascending, as Crout; pivot = last maximum of vv(i) |a(i,k)|, i >= k (:440-449)
#pragma unroll
    for (int K = 0; K < N; ++K) {
      double best = -1.0;
      int imax = K;
#pragma unroll
      for (int i = K; i < N; ++i) {
        const double cand = vv[i] * fabs(J[i][K]);
        if (cand >= best) { best = cand; imax = i; }
      }
#pragma unroll
      for (int i = K + 1; i < N; ++i) {                         // swap rows K and imax (selects: every index is a constant)
        const bool sw = imax == i;
#pragma unroll
        for (int j = 0; j < N; ++j) { const double a = J[K][j], c2 = J[i][j]; J[K][j] = sw ? c2 : a; J[i][j] = sw ? a : c2; }
        { const double a = b[K], c2 = b[i]; b[K] = sw ? c2 : a; b[i] = sw ? a : c2; }
        if (sw) vv[i] = vv[K];                                  // :453
      }
      if (J[K][K] == 0.0) J[K][K] = 1.0e-20;
      const double dum = 1.0 / J[K][K];
#pragma unroll
      for (int i = K + 1; i < N; ++i) {
        const double lik = J[i][K] * dum;
#pragma unroll
        for (int j = K + 1; j < N; ++j) J[i][j] = J[i][j] - lik * J[K][j];
        b[i] = b[i] - lik * b[K];
      }
      vv[K] = dum;                                              // vv(K) is dead: keep 1/a(K,K) for the back substitution
    }
    double x[N];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {                          // lubksb :511-520 (REASSOC: times 1/a(i,i))
      double sum = b[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) sum = sum - J[i][j] * x[j];
      x[i] = sum * vv[i];
    }
    // update (:3454-3498)
    double min_ratio = 1.0e20;
    if (!T.use_log) {                                          // :3459-3471
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < n && m[i] <= x[i]) { const double ratio = fabs(m[i] / x[i]); if (ratio < min_ratio) min_ratio = ratio; }
    }
    double maxrel = 0.0, nw[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      nw[i] = m[i];
      if (i < n) {
        double u = x[i];
        const double prev = m[i];
        if (T.use_log) {                                        // :3454-3458
          u = copysign(1.0, u) * fmin(fabs(u), T.max_dlnC);
          nw[i] = prev * exp(-u);
        } else {
          if (min_ratio < 1.0) u = u * min_ratio * 0.99;
          nw[i] = prev - u;
        }
        const double rc = fabs((nw[i] - prev) / prev);
        if (!isfinite(rc)) bad = true;
        maxrel = fmax(maxrel, rc);
        if (iter > 50) nw[i] = 0.1 * (nw[i] - prev) + prev;     // :3478-3496
      }
    }
    if (bad) { status = RXN_FLAG_NONFINITE; closing = true; continue; }
    if (maxrel < T.rel_tol) { status = RXN_EXIT_REL_CHANGE; break; }   // :3476 (update discarded)
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = nw[i];                   // :3498
    if (iter >= T.maxit) { status = RXN_FLAG_CAPPED; closing = true; continue; }   // GPU-only guard (reference spins)
  }
  // closing RTAuxVarCompute (:3507): after a normal exit m and gamma are those of the last RTotal, so sec_molal / total are final
  // (an abnormal exit went through the closing pass above); write back (reactive_transport.F90:1711)
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < n) {
      tran_xx[item * n + i] = m[i];
      SGS(S, RXN_F_PRI_MOLAL, i, cell) = m[i];
      SGS(S, RXN_F_TOTAL, i, cell) = tot[i];
      if (!T.act_off) SGS(S, RXN_F_PRI_ACT_COEF, i, cell) = exp(lngp[i]);
    }
#pragma unroll
  for (int k = 0; k < MC; ++k)
    if (k < ncplx) {
      SGS(S, RXN_F_SEC_MOLAL, k, cell) = sm[k];
      if (!T.act_off) SGS(S, RXN_F_SEC_ACT_COEF, k, cell) = exp(lngc[k]);
    }
#pragma unroll
  for (int q = 0; q < MK; ++q)
    if (q < nkin) SGS(S, RXN_F_MNRL_RATE, q, cell) = rate_out[q];
  SGS(S, RXN_F_LN_ACT_H2O, 0, cell) = ln_act_h2o;
  if (iters) iters[item] = iter;
  if (flags) flags[item] = status | cflags;
}

}  // namespace small
}  // namespace rxn
