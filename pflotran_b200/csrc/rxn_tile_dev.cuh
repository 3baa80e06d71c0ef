// rxn_tile_dev.cuh — device code of the cooperative (lane-group per cell) RReact kernel.
// See rxn_tile.cuh for the design.  Reference routines restated here (file:line at each site):
// RReact reaction.F90:3322-3511, RTotal :4057-4158, RActivityCoefficients (LAG) :3994-4050,
// RTotalSorbEqSurfCplx1 reaction_surf_complex.F90:658-934, RMultiRateSorption :566-654,
// RKineticMineral reaction_mineral.F90:564-1000, RSolve reaction.F90:4835-4880,
// ludcmp/lubksb utility.F90:393-523.
//
// Code-size discipline: the first version of this kernel was 22 k SASS instructions and spent most
// of its issue slots waiting for instruction fetch (ncu: stall_no_instruction 4.5 per issue).  The
// Newton loop is therefore written so that every routine has ONE call site (the closing
// RTAuxVarCompute is a last, shortened trip through the same loop body), row loops are runtime
// loops over shared-memory vectors instead of unrolled register arrays, exp/log are shared
// non-inlined copies, and rarely used branches (mineral prefactors, per-cell logK) are cold calls.
#pragma once
#include <algorithm>

#include "rxn_device.cuh"
#include "rxn_tile.cuh"

namespace rxn {

namespace {

constexpr int CODE_LAST = 1 << 16;

// All shared-memory accesses are written as offsets from this symbol so that the compiler keeps
// them in the shared address space (LDS/STS with 32-bit addresses).  Pointers carried through a
// struct degrade to generic LD/ST with 64-bit address arithmetic (ncu, first compact version:
// 343 LD / 136 ST, no LDS, long-scoreboard stalls on every table read).
extern __shared__ __align__(16) double tsm[];

struct TCtx {              // per-group context (all lanes of a group hold identical copies)
  const DevTab *h;
  const TileTab *tt;
  const DevState *S;
  int d_off, i_off;        // main table blob: doubles (double index) / ints (int index)
  int pd_off, pi_off;      // plan blob
  int cs_off;              // this cell's shared-memory area (double index)
  int nlk_off;             // -logK*LOG_TO_LN for [complexes | kinetic minerals | surface complexes]
  long long cell;
  unsigned gm;             // lane mask of the group
  int l;                   // lane within the group
  int flags;
  double ln_act_h2o, den_kg, temp, pres, volume, porosity, soil_density;
  __device__ __forceinline__ double *cs() const { return tsm + cs_off; }
  __device__ __forceinline__ const double *d() const { return tsm + d_off; }
  __device__ __forceinline__ const double *pd() const { return tsm + pd_off; }
  __device__ __forceinline__ const double *nlk() const { return tsm + nlk_off; }
  __device__ __forceinline__ const int *i() const { return reinterpret_cast<const int *>(tsm) + i_off; }
  __device__ __forceinline__ const int *pi() const { return reinterpret_cast<const int *>(tsm) + pi_off; }
};

#define GS(c, field, row) ((c).S->f[field][(long long)(row) * (c).S->ld + (c).cell])

__device__ __noinline__ double c_exp(double x) { return exp(x); }
__device__ __noinline__ double c_log(double x) { return log(x); }

template <int G>
__device__ __forceinline__ double grp_sum(double v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o, G);
  return v;
}
template <int G>
__device__ __forceinline__ double grp_max(double v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gm, v, o, G));
  return v;
}
template <int G>
__device__ __forceinline__ double grp_min(double v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(gm, v, o, G));
  return v;
}
__device__ __forceinline__ bool grp_any(bool p, unsigned gm) { return (__ballot_sync(gm, p) & gm) != 0u; }

// ---------------------------------------------------------------------------------------------
// RActivityCoefficients, LAG algorithm — reaction.F90:3994-4050.  One Debye-Hueckel exponent per
// distinct (Z^2, a0) class (species of a class have bit-identical coefficients); class 0 = neutral
// (|Z| <= 1e-10, gamma = 1).  ln gamma of the primaries -> shared vector lgp.
template <int G>
__device__ __forceinline__ void t_act_coefs(TCtx &c) {
  const DevTab &h = *c.h;
  const TileTab &tt = *c.tt;
  const int n = h.naq, ncplx = h.ncplx;
  const double *pz2 = c.pd() + tt.o_pz2, *cz2 = c.pd() + tt.o_cz2;
  const double *sm = c.cs() + tt.c_sm, *vm = c.cs() + tt.c_m;
  double part = 0.0, psum = 0.0;
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    const double mm = vm[row];
    part += mm * pz2[row];
    if (row + 1 != h.h2o_aq_id) psum += mm;
  }
#pragma unroll 1
  for (int k = c.l; k < ncplx; k += G) {
    part += sm[k] * cz2[k];
    psum += sm[k];
  }
  const double I = 0.5 * grp_sum<G>(part, c.gm);            // REASSOC: tree sum, Z^2 premultiplied
  const double sqrt_I = sqrt(I);
  const double *z2 = c.pd() + tt.o_cls_z2, *a0 = c.pd() + tt.o_cls_a0;
  double *lngc = c.cs() + tt.c_lng;
#pragma unroll 1
  for (int q = c.l; q < tt.ncls; q += G)
    lngc[q] = (q == 0) ? 0.0 : (-z2[q] * sqrt_I * h.debyeA / (1.0 + a0[q] * h.debyeB * sqrt_I) + h.debyeBdot * I) * RXN_LOG_TO_LN;
  if (h.use_act_h2o) {                                         // :4043-4050
    const double s = grp_sum<G>(psum, c.gm);
    const double a = 1.0 - 0.017 * s;
    c.ln_act_h2o = (a > 0.0) ? c_log(a) : 0.0;
  }
  __syncwarp(c.gm);
  const int *pcls = c.pi() + tt.o_pri_cls;
  double *lgp = c.cs() + tt.c_lgp;
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    lgp[row] = lngc[pcls[row]];
    if (tt.need_gam) c.cs()[tt.c_gam + row] = c_exp(lgp[row]);
  }
}

// ln a = ln m + ln gamma and 1/m of the owned primaries -> shared; then sec_molal of the
// complexes k = l, l+G, ... (RTotal, reaction.F90:4090-4122)
template <int G>
__device__ __forceinline__ void t_speciate(TCtx &c, bool act_off) {
  const DevTab &h = *c.h;
  const TileTab &tt = *c.tt;
  const int n = h.naq;
  double *cs = c.cs();
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    const double mm = cs[tt.c_m + row];
    cs[tt.c_invm + row] = 1.0 / mm;
    cs[tt.c_lna + row] = c_log(mm) + cs[tt.c_lgp + row];
  }
  __syncwarp(c.gm);
  const int *ptr = c.i() + h.cplx.o_ptr, *id = c.i() + h.cplx.o_id;
  const double *st = c.d() + h.cplx.o_st, *h2ost = c.d() + h.cplx.o_h2ost;
  const double *lna = cs + tt.c_lna, *lngc = cs + tt.c_lng;
  const int *ccls = c.pi() + tt.o_cplx_cls;
  double *sm = cs + tt.c_sm;
#pragma unroll 1
  for (int k = c.l; k < h.ncplx; k += G) {
    double lnQK = c.nlk()[k];
    if (h2ost[k] != 0.0) lnQK = lnQK + h2ost[k] * c.ln_act_h2o;
    const int p1 = ptr[k + 1];
#pragma unroll 1
    for (int p = ptr[k]; p < p1; ++p) lnQK = lnQK + st[p] * lna[id[p]];
    // REASSOC: exp(lnQK)/gamma_k -> exp(lnQK - ln gamma_k)
    const double lg = act_off ? c_log(GS(c, RXN_F_SEC_ACT_COEF, k)) : lngc[ccls[k]];
    sm[k] = c_exp(lnQK - lg);
  }
  __syncwarp(c.gm);
}

// ---------------------------------------------------------------------------------------------
// sparse accumulation plans (host-built, rxn_tile.cu).  16-byte records {coef, code, key}; lane l
// walks records l, l+G, l+2G, ...; a record with CODE_LAST closes the current entry `key`.
//   plan A: total_i = (m_i + sum_k nu_ik sm_k) * den                       (reaction.F90:4095,4124,4148)
//   plan B: J_ij    = ((delta_ij + (sum_k nu_ik nu_jk sm_k)/m_j) * den) * psvd, i <= j and mirrored
//           (:4098-4101, 4126-4146, 4149; RTAccumulationDerivative :5189-5204)
// Per entry the k order is ascending, as in the reference's loop over complexes.
template <int G, bool JAC>
__device__ __forceinline__ void t_plan(const TCtx &c, double den, double psvd) {
  const TileTab &tt = *c.tt;
  const int4 *__restrict__ rec = reinterpret_cast<const int4 *>(c.pd() + (JAC ? tt.o_B_rec : tt.o_A_rec)) + c.l;
  const int T = JAC ? tt.TB : tt.TA;
  const double *__restrict__ sm = c.cs() + tt.c_sm;
  const double *__restrict__ invm = c.cs() + tt.c_invm;
  const double *__restrict__ vm = c.cs() + tt.c_m;
  double *__restrict__ J = c.cs();
  double *__restrict__ tot = c.cs() + tt.c_tot;
  const int LDJ = tt.LDJ;
  double acc = 0.0;
  // the table / sec_molal loads of the next records do not depend on the stores of a closing entry
  // (__restrict__): unrolled by 4 so that their shared-memory latencies overlap
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const int4 v = rec[t * G];
    acc = fma(__hiloint2double(v.y, v.x), sm[v.z & 0xffff], acc);
    if (v.z & CODE_LAST) {
      if (!JAC) {
        tot[v.w] = (vm[v.w] + acc) * den;
      } else {
        const int i = v.w >> 8, j = v.w & 0xff;
        const double a = acc * den;
        if (i == j) {
          J[i * LDJ + i] = ((1.0 + acc * invm[i]) * den) * psvd;
        } else {                                             // REASSOC: 1/m_j factored out of the k sum
          J[i * LDJ + j] = (a * invm[j]) * psvd;
          J[j * LDJ + i] = (a * invm[i]) * psvd;
        }
      }
      acc = 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RTotalSorbEqSurfCplx1 — reaction_surf_complex.F90:658-934, for one surface complexation
// reaction.  Complexes are spread over the lanes; totals / Jacobian rows go to the owner lane.
//   tvec[row] += total sorbed of the owned primaries
//   addJ: J(row, jc) += fac * d(total_sorb_row)/d(m_jc)   (REASSOC: term by term, no dense temp)
template <int G>
__device__ __noinline__ void t_srf_rxn(TCtx &c, int irxn, double fac, bool addJ, double *tvec, bool store_conc) {
  const DevTab &h = *c.h;
  const TileTab &tt = *c.tt;
  const int n = h.naq, LDJ = tt.LDJ;
  const double tol = 1.0e-12;
  double *cs = c.cs(), *J = c.cs();
  const int c0 = c.i()[h.o_rxn_cptr + irxn], c1 = c.i()[h.o_rxn_cptr + irxn + 1];
  const int *cid = c.i() + h.o_rxn_cid;
  const int *sptr = c.i() + h.srf.o_ptr, *sid = c.i() + h.srf.o_id;
  const double *sst = c.d() + h.srf.o_st, *site_st = c.d() + h.o_srf_site_st, *sh2o = c.d() + h.srf.o_h2ost;
  const double *nlk = c.nlk() + h.ncplx + h.nkin;
  const double *lna = cs + tt.c_lna, *invm = cs + tt.c_invm;
  double *sc = cs + tt.c_sc, *dsx = cs + tt.c_dsx;
  double free_site_conc = cs[tt.c_free + irxn];
  double site_density;
  const int surf_type = c.i()[h.o_rxn_surf_type + irxn];
  const double dens = c.d()[h.o_rxn_density + irxn];
  if (surf_type == RXN_MINERAL_SURFACE)
    site_density = dens * cs[tt.c_mnrl + c.i()[h.o_rxn_to_surf + irxn] - 1];
  else if (surf_type == RXN_ROCK_SURFACE)
    site_density = dens * c.soil_density * (1.0 - c.porosity);
  else
    site_density = dens;
  if (site_density < 1.0e-40) return;                        // :749
  const int stoich_flag = c.i()[h.o_rxn_flag + irxn];
  bool one_more = false;
  int num_iterations = 0;
  double damping_factor = 1.0;
  __syncwarp(c.gm);
#pragma unroll 1
  for (;;) {                                                  // :760-829
    num_iterations = num_iterations + 1;
    const double ln_free_site = c_log(free_site_conc);
    double part = 0.0, part2 = 0.0;
#pragma unroll 1
    for (int j = c0 + c.l; j < c1; j += G) {
      const int icplx = cid[j];
      double lnQK = nlk[icplx];
      if (sh2o[icplx] != 0.0) lnQK = lnQK + sh2o[icplx] * c.ln_act_h2o;
      lnQK = lnQK + site_st[icplx] * ln_free_site;
#pragma unroll 1
      for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p) lnQK = lnQK + sst[p] * lna[sid[p]];
      const double s = c_exp(lnQK);
      sc[j - c0] = s;
      part += site_st[icplx] * s;
      part2 += site_st[icplx] * s / free_site_conc;
    }
    double total = free_site_conc + grp_sum<G>(part, c.gm);   // REASSOC: tree sum
    if (one_more) break;
    if (stoich_flag) {
      const double res = site_density - total;
      const double dres_dfree_site = 1.0 + grp_sum<G>(part2, c.gm);
      const double dfree_site_conc = res / dres_dfree_site;
      if (num_iterations > 1000) damping_factor = 0.5;
      free_site_conc = free_site_conc + damping_factor * dfree_site_conc;
      const double rel_change = fabs(dfree_site_conc / free_site_conc);
      if (rel_change < tol) one_more = true;
      if (num_iterations > 100000) { c.flags |= RXN_FLAG_CAPPED; one_more = true; }  // reference would spin
    } else {
      total = total / free_site_conc;
      free_site_conc = site_density / total;
      one_more = true;
    }
  }
  __syncwarp(c.gm);
  cs[tt.c_free + irxn] = free_site_conc;                      // all lanes write the same value

  double tempreal = 0.0;                                      // :838-866 (redundant per lane: few complexes)
#pragma unroll 1
  for (int j = c0; j < c1; ++j) tempreal = tempreal + site_st[cid[j]] * site_st[cid[j]] * sc[j - c0];
  tempreal = tempreal / free_site_conc;
  tempreal = tempreal + 1.0;
  const double inv_free = 1.0 / free_site_conc;
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    double b = 0.0;
#pragma unroll 1
    for (int j = c0; j < c1; ++j) {
      const int icplx = cid[j];
#pragma unroll 1
      for (int p = sptr[icplx]; p < sptr[icplx + 1]; ++p)
        if (sid[p] == row) b = b + sst[p] * site_st[icplx] * sc[j - c0];
    }
    b = -b / tempreal;
    dsx[row] = b * invm[row];                                 // REASSOC: /m -> *(1/m)
  }
  if (store_conc) {
#pragma unroll 1
    for (int j = c0 + c.l; j < c1; j += G) GS(c, RXN_F_EQSRFCPLX_CONC, cid[j]) += sc[j - c0];
  }
  __syncwarp(c.gm);
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {                     // :872-931
    double tsum = tvec[row];
#pragma unroll 1
    for (int j = c0; j < c1; ++j) {
      const int icplx = cid[j];
      const int p0 = sptr[icplx], p1 = sptr[icplx + 1];
#pragma unroll 1
      for (int p = p0; p < p1; ++p) {
        if (sid[p] != row) continue;
        const double s = sc[j - c0];
        tsum = tsum + sst[p] * s;
        if (addJ) {
          const double nui_Si_over_Sx = site_st[icplx] * s * inv_free;
#pragma unroll 1
          for (int q = p0; q < p1; ++q) {
            const int jc = sid[q];
            const double tr = sst[q] * s * invm[jc] + nui_Si_over_Sx * dsx[jc];
            J[row * LDJ + jc] += (sst[p] * tr) * fac;
          }
        }
      }
    }
    tvec[row] = tsum;
  }
}

// ---------------------------------------------------------------------------------------------
// RKineticMineral — reaction_mineral.F90:564-1000.  Every lane evaluates the (few) rate laws
// redundantly (uniform, no exchange); the owner lane of a primary adds its residual entry
// (column NP of J) and its Jacobian row.  Prefactor terms (:743-790, :905-997) are cold calls.
struct PrefScratch {
  double prefactor[RXN_MAX_PREF];
  double ln_prefactor_spec[RXN_MAX_PREF][RXN_MAX_PREF_SPEC];
};

__device__ __noinline__ double t_mnrl_prefactor_rate(const TCtx &c, int imnrl, int npref, PrefScratch &ps) {
  const DevTab &h = *c.h;
  const double *lna = c.cs() + c.tt->c_lna;
  const int mp = h.maxpref > 1 ? h.maxpref : 1, mps = h.maxprefspec > 1 ? h.maxprefspec : 1;
  double sum_prefactor_rate = 0.0;
  for (int ipref = 0; ipref < npref; ++ipref) {
    double ln_prefactor = 0.0;
    const int pb = imnrl * mp + ipref;
    const int nps = c.i()[h.o_pref_id + pb * (h.maxprefspec + 1)];
    for (int ips = 0; ips < nps; ++ips) {
      const int icomp = c.i()[h.o_pref_id + pb * (h.maxprefspec + 1) + ips + 1];
      const double ln_spec_act = lna[icomp - 1];
      const double ln_numerator = c.d()[h.o_pref_alpha + pb * mps + ips] * ln_spec_act;
      const double ln_denominator =
          log(1.0 + exp(log(c.d()[h.o_pref_atten + pb * mps + ips]) + c.d()[h.o_pref_beta + pb * mps + ips] * ln_spec_act));
      ln_prefactor = ln_prefactor + ln_numerator;
      ln_prefactor = ln_prefactor - ln_denominator;
      ps.ln_prefactor_spec[ipref][ips] = ln_numerator - ln_denominator;
    }
    ps.prefactor[ipref] = exp(ln_prefactor);
    double arrhenius_factor = 1.0;
    const double Ea = c.d()[h.o_pref_Ea + pb];
    if (Ea > 0.0) arrhenius_factor = exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
    sum_prefactor_rate = sum_prefactor_rate + ps.prefactor[ipref] * c.d()[h.o_pref_rate + pb] * arrhenius_factor;
  }
  return sum_prefactor_rate;
}

template <int G>
__device__ __noinline__ void t_mnrl_prefactor_jac(const TCtx &c, int imnrl, int npref, const PrefScratch &ps, double Im,
                                                  double sum_prefactor_rate) {
  const DevTab &h = *c.h;
  const TileTab &tt = *c.tt;
  const int LDJ = tt.LDJ;
  double *J = c.cs();
  const double *lna = c.cs() + tt.c_lna, *gam = c.cs() + tt.c_gam;
  const int *ptr = c.i() + h.kin.o_ptr, *id = c.i() + h.kin.o_id;
  const double *st = c.d() + h.kin.o_st;
  const int p0 = ptr[imnrl], p1 = ptr[imnrl + 1];
  const int mp = h.maxpref > 1 ? h.maxpref : 1, mps = h.maxprefspec > 1 ? h.maxprefspec : 1;
  const double dIm_dsum_prefactor_rate = Im / sum_prefactor_rate;
  for (int ipref = 0; ipref < npref; ++ipref) {
    const int pb = imnrl * mp + ipref;
    double arrhenius_factor = 1.0;
    const double Ea = c.d()[h.o_pref_Ea + pb];
    if (Ea > 0.0) arrhenius_factor = exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
    const double ln_prefactor = log(ps.prefactor[ipref]);
    const int nps = c.i()[h.o_pref_id + pb * (h.maxprefspec + 1)];
    for (int ips = 0; ips < nps; ++ips) {
      const double dprefactor_dprefactor_spec = exp(ln_prefactor - ps.ln_prefactor_spec[ipref][ips]);
      const int icomp = c.i()[h.o_pref_id + pb * (h.maxprefspec + 1) + ips + 1];
      const double ln_spec_act = lna[icomp - 1], spec_act_coef = gam[icomp - 1];
      const double alpha = c.d()[h.o_pref_alpha + pb * mps + ips], beta = c.d()[h.o_pref_beta + pb * mps + ips],
                   atten = c.d()[h.o_pref_atten + pb * mps + ips];
      const double dnum = alpha * exp(ps.ln_prefactor_spec[ipref][ips] - ln_spec_act);
      const double ln_gam_m_beta = beta * ln_spec_act;
      const double denominator = 1.0 + exp(log(atten) + ln_gam_m_beta);
      const double dden = -1.0 * exp(ps.ln_prefactor_spec[ipref][ips]) / denominator * atten * beta * exp(ln_gam_m_beta - ln_spec_act);
      double dprefactor_spec_dspec = dnum + dden;
      dprefactor_spec_dspec = dprefactor_spec_dspec * spec_act_coef;
      const double dIm_dspec = dIm_dsum_prefactor_rate * dprefactor_dprefactor_spec * dprefactor_spec_dspec *
                               c.d()[h.o_pref_rate + pb] * arrhenius_factor;
      for (int p = p0; p < p1; ++p)
        if ((id[p] & (G - 1)) == c.l) J[id[p] * LDJ + (icomp - 1)] += st[p] * dIm_dspec;
    }
  }
}

__device__ __noinline__ double c_pow(double x, double y) { return pow(x, y); }

template <int G>
__device__ __forceinline__ void t_kinetic_mineral(TCtx &c) {
  const DevTab &h = *c.h;
  const TileTab &tt = *c.tt;
  const int LDJ = tt.LDJ, NP = tt.NP;
  double *J = c.cs();
  const int *ptr = c.i() + h.kin.o_ptr, *id = c.i() + h.kin.o_id;
  const double *st = c.d() + h.kin.o_st;
  const double *lna = c.cs() + tt.c_lna, *invm = c.cs() + tt.c_invm;
  const double *nlk = c.nlk() + h.ncplx;
#pragma unroll 1
  for (int imnrl = 0; imnrl < h.nkin; ++imnrl) {
    double rate_out = 0.0;
    do {
      double lnQK = nlk[imnrl];
      const double h2ost = c.d()[h.kin.o_h2ost + imnrl];
      if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
      const int p0 = ptr[imnrl], p1 = ptr[imnrl + 1];
#pragma unroll 1
      for (int p = p0; p < p1; ++p) lnQK = lnQK + st[p] * lna[id[p]];
      double QK;
      if (lnQK <= 6.90776) QK = c_exp(lnQK); else QK = 1.0e3;
      const double k_scale = h.has_scale ? c.d()[h.o_k_scale + imnrl] : 1.0;
      const double k_Temkin = h.has_Temkin ? c.d()[h.o_k_Temkin + imnrl] : 1.0;
      const double k_power = h.has_power ? c.d()[h.o_k_power + imnrl] : 1.0;
      const double k_lim = c.d()[h.o_k_lim + imnrl];
      const double k_aff = c.d()[h.o_k_aff + imnrl];
      const int npref = c.i()[h.o_k_npref + imnrl];
      double affinity_factor;
      if (h.has_Temkin) {
        if (h.has_scale) affinity_factor = 1.0 - c_pow(QK, 1.0 / (k_scale * k_Temkin));
        else affinity_factor = 1.0 - c_pow(QK, 1.0 / k_Temkin);
      } else if (h.has_scale) {
        affinity_factor = 1.0 - c_pow(QK, 1.0 / k_scale);
      } else {
        affinity_factor = 1.0 - QK;
      }
      const double sign_ = copysign(1.0, affinity_factor);
      double Im, Im_const, sum_prefactor_rate;
      PrefScratch ps;
      const double volfrac = c.cs()[tt.c_mnrl + imnrl];
      if (!(volfrac > 0 || sign_ < 0.0)) break;
      if (k_aff > 0.0) {
        if (sign_ < 0.0 && QK < k_aff) break;
      }
      if (k_lim > 0.0) affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / k_lim);
      if (npref > 0) {
        sum_prefactor_rate = t_mnrl_prefactor_rate(c, imnrl, npref, ps);
      } else {
        double arrhenius_factor = 1.0;
        const double Ea = c.d()[h.o_k_Ea + imnrl];
        if (Ea > 0.0) arrhenius_factor = c_exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
        sum_prefactor_rate = c.d()[h.o_k_rate + imnrl] * arrhenius_factor;
      }
      Im_const = -c.cs()[tt.c_mnrl + h.nkin + imnrl];
      if (h.has_scale) Im_const = Im_const / k_scale;
      if (h.has_power) Im = Im_const * sign_ * c_pow(fabs(affinity_factor), k_power) * sum_prefactor_rate;
      else Im = Im_const * sign_ * fabs(affinity_factor) * sum_prefactor_rate;
      rate_out = Im;

      Im_const = Im_const * c.volume;
      Im = Im * c.volume;
      double dIm_dQK;
      if (h.has_power) dIm_dQK = -Im * k_power / fabs(affinity_factor);
      else dIm_dQK = -Im_const * sum_prefactor_rate;
      if (h.has_Temkin) {
        if (h.has_scale) dIm_dQK = dIm_dQK * (1.0 / (k_scale * k_Temkin)) / QK * (1.0 - affinity_factor);
        else dIm_dQK = dIm_dQK * (1.0 / k_Temkin) / QK * (1.0 - affinity_factor);
      } else if (h.has_scale) {
        dIm_dQK = dIm_dQK * (1.0 / k_scale) / QK * (1.0 - affinity_factor);
      }
      const double den = (k_lim <= 0.0) ? 1.0 : 1.0 + (1.0 - affinity_factor) / k_lim;
#pragma unroll 1
      for (int p = p0; p < p1; ++p) {
        const int ip = id[p];
        if ((ip & (G - 1)) != c.l) continue;                    // owner lane of primary ip
        double *a = J + ip * LDJ;
        a[NP] += st[p] * Im;
#pragma unroll 1
        for (int q = p0; q < p1; ++q) {
          const int jcomp = id[q];
          const double dQK_dCj = st[q] * QK * invm[jcomp];      // REASSOC: exp(-ln m_j) -> 1/m_j
          const double dQK_dmj = dQK_dCj * c.den_kg * 1.0e-3;
          if (k_lim <= 0.0) a[jcomp] += st[p] * dIm_dQK * dQK_dmj;
          else a[jcomp] += st[p] * dIm_dQK * (1.0 + QK / k_lim / den) * dQK_dmj / den;
        }
      }
      if (npref > 0) t_mnrl_prefactor_jac<G>(c, imnrl, npref, ps, Im, sum_prefactor_rate);
    } while (false);
    if (c.l == 0) GS(c, RXN_F_MNRL_RATE, imnrl) = rate_out;     // :575 (zeroed) / :816
  }
}

// ---------------------------------------------------------------------------------------------
// RSolve (reaction.F90:4835-4880) + ludcmp/lubksb (utility.F90:393-523) on the augmented system
// [J | b] in shared memory (b = column NP, rows 16-byte aligned, LDJ even).  Right-looking
// elimination: per element the same a(i,j) -= a(i,k)*a(k,j), k ascending, as Crout; pivot = last
// maximum of vv(i)*|a(i,k)|, i >= k.  The forward substitution of lubksb is the elimination
// applied to column NP.  Returns 1 if a row is all zero (reference: MPI_Abort).
template <int G>
__device__ __forceinline__ int t_rsolve(TCtx &c, bool use_log) {
  const TileTab &tt = *c.tt;
  const int n = c.h->naq, LDJ = tt.LDJ, NP = tt.NP;
  double *J = c.cs();
  const double *__restrict__ vm = c.cs() + tt.c_m;
  double *vv = c.cs() + tt.c_dsx;                              // scratch shared with the sorption routine
  const double tiny = 1.0e-20;
  bool zero = false;
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    double *__restrict__ a = J + row * LDJ;
    double mx = 0.0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(a[j]));
    const double norm = 1.0 / fmax(1.0, mx);
    a[NP] = a[NP] * norm;
    double aamax = 0.0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      double v = a[j] * norm;
      if (use_log) v = v * vm[j];
      a[j] = v;
      aamax = fmax(aamax, fabs(v));
    }
    if (aamax <= 0.0) zero = true;
    vv[row] = 1.0 / aamax;
  }
  if (grp_any(zero, c.gm)) return 1;
  __syncwarp(c.gm);
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    double best = -1.0;
    int bidx = -1;
#pragma unroll 1
    for (int row = c.l + ((k - c.l + G - 1) & ~(G - 1)); row < n; row += G) {   // first owned row >= k
      const double dum = vv[row] * fabs(J[row * LDJ + k]);
      if (dum >= best) { best = dum; bidx = row; }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(c.gm, best, o, G);
      const int oi = __shfl_xor_sync(c.gm, bidx, o, G);
      if (ob > best || (ob == best && oi > bidx)) { best = ob; bidx = oi; }
    }
    const int imax = bidx < 0 ? k : bidx;
    double *pk = J + k * LDJ;
    if (imax != k) {
      double *pi = J + imax * LDJ;
#pragma unroll 1
      for (int col = 2 * c.l; col < LDJ; col += 2 * G) {
        const double2 t = *reinterpret_cast<double2 *>(pi + col);
        *reinterpret_cast<double2 *>(pi + col) = *reinterpret_cast<double2 *>(pk + col);
        *reinterpret_cast<double2 *>(pk + col) = t;
      }
      if (c.l == 0) vv[imax] = vv[k];
    }
    __syncwarp(c.gm);
    double piv = pk[k];
    if (piv == 0.0) piv = tiny;
    const double dum = 1.0 / piv;
    __syncwarp(c.gm);
    if (c.l == 0 && pk[k] == 0.0) pk[k] = tiny;
#pragma unroll 1
    for (int row = c.l + ((k + 1 - c.l + G - 1) & ~(G - 1)); row < n; row += G) {   // first owned row > k
      double *__restrict__ a = J + row * LDJ;                  // own row: never the pivot row
      const double *__restrict__ pr = pk;
      const double lik = a[k] * dum;
      a[k] = lik;
      int col = k + 1;
      if (col & 1) { a[col] = a[col] - lik * pr[col]; ++col; }
#pragma unroll 4
      for (; col < LDJ; col += 2) {                           // includes b (column NP) and padding
        double2 av = *reinterpret_cast<double2 *>(a + col);
        const double2 pv = *reinterpret_cast<const double2 *>(pr + col);
        av.x = av.x - lik * pv.x;
        av.y = av.y - lik * pv.y;
        *reinterpret_cast<double2 *>(a + col) = av;
      }
    }
    __syncwarp(c.gm);
  }
#pragma unroll 1
  for (int k = n - 1; k >= 0; --k) {                           // REASSOC: column-oriented back substitution
    if ((k & (G - 1)) == c.l) J[k * LDJ + NP] = J[k * LDJ + NP] / J[k * LDJ + k];
    __syncwarp(c.gm);
    const double xk = J[k * LDJ + NP];
#pragma unroll 1
    for (int row = c.l; row < k; row += G) J[row * LDJ + NP] = J[row * LDJ + NP] - J[row * LDJ + k] * xk;
  }
  __syncwarp(c.gm);
  return 0;
}

// per-cell logK (non-isothermal tables): RUpdateTempDependentCoefs reaction.F90:5433-5524
template <int G>
__device__ __noinline__ void t_percell_logK(TCtx &c) {
  const DevTab &h = *c.h;
  Cell<1> tc;
  tc.temp = c.temp; tc.pres = c.pres;
  Tab T{c.d(), c.i(), c.h};
  double *lk = c.cs() + c.tt->c_lk;
  for (int k = c.l; k < h.ncplx; k += G) lk[k] = -logK_of(T, h.cplx, k, tc, false) * RXN_LOG_TO_LN;
  for (int k = c.l; k < h.nkin; k += G) lk[h.ncplx + k] = -logK_of(T, h.kin, k, tc, false) * RXN_LOG_TO_LN;
  for (int k = c.l; k < h.nsrf; k += G) lk[h.ncplx + h.nkin + k] = -logK_of(T, h.srf, k, tc, true) * RXN_LOG_TO_LN;
  c.nlk_off = c.cs_off + c.tt->c_lk;
}

// ---------------------------------------------------------------------------------------------
// RReact for one cell by one lane group — reaction.F90:3322-3511 (control flow: SURVEY.md 3.3)
template <int G>
__device__ __forceinline__ void t_react_cell(TCtx &c, long long i, double *tran_xx, const int32_t *l2g, double tran_dt, int dt_mode,
                                             int32_t *iters, int32_t *flags) {
  const DevTab &h = *c.h;
  const TileTab &tt = *c.tt;
  const DevState &S = *c.S;
  const int n = h.naq, LDJ = tt.LDJ, NP = tt.NP;
  double *cs = c.cs(), *J = c.cs();
  c.cell = l2g ? l2g[i] : i;
  c.flags = 0;
  if (S.active && !S.active[c.cell]) {                        // imat <= 0 (reactive_transport.F90:1699)
    if (c.l == 0) {
      if (iters) iters[i] = 0;
      if (flags) flags[i] = RXN_FLAG_INACTIVE;
    }
    return;
  }
  c.ln_act_h2o = GS(c, RXN_F_LN_ACT_H2O, 0);
  c.den_kg = GS(c, RXN_F_DEN_KG, 0);
  c.temp = GS(c, RXN_F_TEMP, 0);
  c.pres = GS(c, RXN_F_PRES, 0);
  c.volume = GS(c, RXN_F_VOLUME, 0);
  c.porosity = GS(c, RXN_F_POROSITY, 0);
  c.soil_density = GS(c, RXN_F_SOIL_PARTICLE_DENSITY, 0);
  const double sat = GS(c, RXN_F_SAT, 0);
  const bool act_off = h.act_freq == RXN_ACT_COEF_FREQUENCY_OFF;
  const double psv_t = c.porosity * sat * 1000.0 * c.volume;
  const double psvd_t = c.porosity * sat * 1000.0 * c.volume / tran_dt;      // :5189
  const double v_t = c.volume / tran_dt;                                       // :4590
  const double den_kg_per_L = c.den_kg * 1.0 * 1.0e-3;
  double *vm = cs + tt.c_m, *vfix = cs + tt.c_fix, *vts = cs + tt.c_tsorb, *vtot = cs + tt.c_tot;
  double mrK1[2] = {0.0, 0.0};

  __syncwarp(c.gm);
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    vm[row] = GS(c, RXN_F_PRI_MOLAL, row);
    const double g = GS(c, RXN_F_PRI_ACT_COEF, row);
    cs[tt.c_lgp + row] = c_log(g);
    if (tt.need_gam) cs[tt.c_gam + row] = g;
    double fx = psv_t * tran_xx[i * n + row];                 // :3370, RTAccumulation :5072-5148
    if (h.neqsorb > 0) fx = fx + GS(c, RXN_F_TOTAL_SORB_EQ, row) * c.volume;   // RAccumulationSorb :4539-4568
    vfix[row] = fx;
    double *a = J + row * LDJ;
    for (int col = n; col < LDJ; ++col) a[col] = 0.0;          // padding columns stay finite
  }
#pragma unroll 1
  for (int k = c.l; k < h.ncplx; k += G) cs[tt.c_sm + k] = GS(c, RXN_F_SEC_MOLAL, k);   // lagged, for I
  if (c.l == 0) cs[tt.c_sm + h.ncplx] = 0.0;                  // plan padding slot
  for (int q = c.l; q < h.nrxn; q += G) cs[tt.c_free + q] = GS(c, RXN_F_FREE_SITE_CONC, q);
  for (int q = c.l; q < h.nkin; q += G) {                     // read-only inside RReact
    cs[tt.c_mnrl + q] = GS(c, RXN_F_MNRL_VOLFRAC, q);
    cs[tt.c_mnrl + h.nkin + q] = GS(c, RXN_F_MNRL_AREA, q);
  }
  if (tt.percell_logK) t_percell_logK<G>(c);
  else c.nlk_off = c.pd_off + tt.o_nlk;
#pragma unroll
  for (int ikr = 0; ikr < 2; ++ikr) {                         // multirate_prepare (rxn_device.cuh; REASSOC)
    if (ikr >= h.nmr) break;
    double K1 = 0.0;
    const int nrate = c.i()[h.o_mr_nrate + ikr];
    double *r0 = cs + tt.c_r0 + ikr * NP;
    for (int row = c.l; row < n; row += G) r0[row] = 0.0;
#pragma unroll 1
    for (int irate = 0; irate < nrate; ++irate) {
      const double rate = c.d()[h.o_mr_rate + ikr * h.mr_ld + irate], frac = c.d()[h.o_mr_frac + ikr * h.mr_ld + irate];
      const double kdt = rate * tran_dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
      const long long row0 = ((long long)ikr * (h.mr_ld + 1) + (irate + 1)) * n;
      for (int row = c.l; row < n; row += G) r0[row] = r0[row] + kk * GS(c, RXN_F_KINMR_TOTAL_SORB, row0 + row);
    }
    mrK1[ikr] = K1;
  }
  __syncwarp(c.gm);

  int num_iterations = 0, reason = 0;
  bool closing = false;                                       // the one last RTAuxVarCompute (:3507)
#pragma unroll 1
  for (;;) {
    if (!closing) {
      num_iterations = num_iterations + 1;
      // :3407-3409 (once, before the loop) and :3413-3418 (every iteration): the call before the
      // loop and the call of iteration 1 see identical inputs, so one evaluation serves both
      if (!act_off && (num_iterations == 1 || h.act_freq == RXN_ACT_COEF_FREQUENCY_NEWTON_ITER)) t_act_coefs<G>(c);
    }
    // RTAuxVarCompute :3419 -> RTotal + RTotalSorb
    t_speciate<G>(c, act_off);
    t_plan<G, false>(c, den_kg_per_L, 0.0);
    if (!closing) t_plan<G, true>(c, den_kg_per_L, psvd_t);   // J <- dtotal * psvd_t  (:3429-3437)
    __syncwarp(c.gm);
    if (h.neqsorb > 0) {
      for (int row = c.l; row < n; row += G) vts[row] = 0.0;
      if (closing && h.neq > 0) {                             // RZeroSorb :4162-4178
        for (int k = c.l; k < h.nsrf; k += G) GS(c, RXN_F_EQSRFCPLX_CONC, k) = 0.0;
        __syncwarp(c.gm);
      }
#pragma unroll 1
      for (int ieq = 0; ieq < h.neq; ++ieq) t_srf_rxn<G>(c, c.i()[h.o_eq_rxn + ieq], v_t, !closing, vts, closing);
    }
    if (closing) break;
#pragma unroll 1
    for (int row = c.l; row < n; row += G) {
      double res = psv_t * vtot[row];
      res = res - vfix[row];                                  // :3424-3426
      if (h.neqsorb > 0) res = res + vts[row] * c.volume;
      if (dt_mode == RXN_DT_CONSISTENT) res = res / tran_dt;
      J[row * LDJ + NP] = res;
    }
    // RReaction :3440 (minerals, then multirate)
    if (h.nkin > 0) t_kinetic_mineral<G>(c);
#pragma unroll
    for (int ikr = 0; ikr < 2; ++ikr) {                       // RMultiRateSorption reaction_surf_complex.F90:566-654
      if (ikr >= h.nmr) break;
      double *seq = cs + tt.c_seq + ikr * NP;
      const double *r0 = cs + tt.c_r0 + ikr * NP;
      for (int row = c.l; row < n; row += G) seq[row] = 0.0;
      t_srf_rxn<G>(c, c.i()[h.o_mr_rxn + ikr], c.volume * mrK1[ikr], true, seq, false);
      for (int row = c.l; row < n; row += G) J[row * LDJ + NP] += c.volume * (mrK1[ikr] * seq[row] - r0[row]);
    }
    double mx = 0.0;
    bool bad = false;
#pragma unroll 1
    for (int row = c.l; row < n; row += G) {
      const double v = J[row * LDJ + NP];
      mx = fmax(mx, fabs(v));
      if (!isfinite(v)) bad = true;
    }
    if (grp_any(bad, c.gm)) { c.flags |= RXN_FLAG_NONFINITE; closing = true; continue; }
    mx = grp_max<G>(mx, c.gm);
    if (mx < h.res_tol) { reason = RXN_EXIT_RESIDUAL; closing = true; continue; }     // :3443
    if (t_rsolve<G>(c, h.use_log != 0)) { c.flags |= RXN_FLAG_LU_ZERO_ROW; closing = true; continue; }
    // update (new solution staged in column NP of J)
    double maxrel = 0.0;
    bad = false;
    double min_ratio = 1.0e20;
    if (!h.use_log) {                                                      // :3459-3471
#pragma unroll 1
      for (int row = c.l; row < n; row += G) {
        const double u = J[row * LDJ + NP];
        if (vm[row] <= u) {
          const double ratio = fabs(vm[row] / u);
          if (ratio < min_ratio) min_ratio = ratio;
        }
      }
      min_ratio = grp_min<G>(min_ratio, c.gm);
    }
#pragma unroll 1
    for (int row = c.l; row < n; row += G) {
      double u = J[row * LDJ + NP];
      const double prev = vm[row];
      double nw;
      if (h.use_log) {                                                     // :3454-3458
        u = copysign(1.0, u) * fmin(fabs(u), h.max_dlnC);
        nw = prev * c_exp(-u);
      } else {
        if (min_ratio < 1.0) u = u * min_ratio * 0.99;
        nw = prev - u;
      }
      const double rc = fabs((nw - prev) / prev);
      if (!isfinite(rc)) bad = true;
      maxrel = fmax(maxrel, rc);
      if (num_iterations > 50) nw = 0.1 * (nw - prev) + prev;              // :3478-3496
      J[row * LDJ + NP] = nw;
    }
    if (grp_any(bad, c.gm)) { c.flags |= RXN_FLAG_NONFINITE; closing = true; continue; }
    maxrel = grp_max<G>(maxrel, c.gm);
    if (maxrel < h.rel_tol) { reason = RXN_EXIT_REL_CHANGE; closing = true; continue; }  // :3476 (update discarded)
#pragma unroll 1
    for (int row = c.l; row < n; row += G) vm[row] = J[row * LDJ + NP];    // :3498
    if (num_iterations >= h.maxit) { c.flags |= RXN_FLAG_CAPPED; closing = true; }   // GPU-only guard (reference spins)
  }
  __syncwarp(c.gm);
  // write back (store_cell of the thread-per-cell path + reactive_transport.F90:1711)
#pragma unroll 1
  for (int row = c.l; row < n; row += G) {
    const double mm = vm[row];
    tran_xx[i * n + row] = mm;
    GS(c, RXN_F_PRI_MOLAL, row) = mm;
    if (!act_off) GS(c, RXN_F_PRI_ACT_COEF, row) = c_exp(cs[tt.c_lgp + row]);
    GS(c, RXN_F_TOTAL, row) = vtot[row];
    if (h.neqsorb > 0) GS(c, RXN_F_TOTAL_SORB_EQ, row) = vts[row];
    for (int ikr = 0; ikr < h.nmr; ++ikr)
      GS(c, RXN_F_KINMR_TOTAL_SORB, (long long)ikr * (h.mr_ld + 1) * n + row) = cs[tt.c_seq + ikr * NP + row];
  }
  if (!act_off) {
    double *lngc = cs + tt.c_lng;
    for (int q = c.l; q < tt.ncls; q += G) lngc[q] = c_exp(lngc[q]);       // gamma per class
    __syncwarp(c.gm);
    const int *ccls = c.pi() + tt.o_cplx_cls;
#pragma unroll 1
    for (int k = c.l; k < h.ncplx; k += G) GS(c, RXN_F_SEC_ACT_COEF, k) = lngc[ccls[k]];
  }
#pragma unroll 1
  for (int k = c.l; k < h.ncplx; k += G) GS(c, RXN_F_SEC_MOLAL, k) = cs[tt.c_sm + k];
  for (int q = c.l; q < h.nrxn; q += G) GS(c, RXN_F_FREE_SITE_CONC, q) = cs[tt.c_free + q];
  if (c.l == 0) {
    GS(c, RXN_F_LN_ACT_H2O, 0) = c.ln_act_h2o;
    if (iters) iters[i] = num_iterations;
    if (flags) flags[i] = reason | c.flags;
  }
  __syncwarp(c.gm);
}

template <int G>
__global__ void __launch_bounds__(tile_max_threads(G), 1)
k_react_tile(const __grid_constant__ DevTab tab, const __grid_constant__ TileTab tt, const double *__restrict__ blob,
             const double *__restrict__ pblob, DevState S, double *tran_xx, const int32_t *__restrict__ l2g, long long nlocal,
             double dt, int dt_mode, int32_t *iters, int32_t *flags) {
  // [plan blob (16-byte records first)][main table blob][per-cell areas]
  const int w1 = tt.ndbl + tt.nint / 2, w0 = tab.ndbl + tab.nint / 2;
  const int o_main = (w1 + 1) & ~1, o_cells = (o_main + w0 + 1) & ~1;
  for (int w = threadIdx.x; w < w1; w += blockDim.x) tsm[w] = pblob[w];
  for (int w = threadIdx.x; w < w0; w += blockDim.x) tsm[o_main + w] = blob[w];
  __syncthreads();
  TCtx c;
  c.h = &tab; c.tt = &tt; c.S = &S;
  c.pd_off = 0; c.pi_off = 2 * tt.ndbl;
  c.d_off = o_main; c.i_off = 2 * (o_main + tab.ndbl);
  const int lane = threadIdx.x & 31;
  const int giw = lane / G, grp = (threadIdx.x >> 5) * tt.gpw + giw;   // group within the warp / within the CTA
  if (giw >= tt.gpw || grp >= tt.cpb) return;                 // idle lanes (no CTA-wide barrier below)
  c.l = threadIdx.x & (G - 1);
  c.gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  c.cs_off = o_cells + grp * tt.pc_dbl;
  const int cpb = tt.cpb;
#pragma unroll 1
  for (long long tile = blockIdx.x; tile * cpb < nlocal; tile += gridDim.x) {
    const long long i = tile * cpb + grp;
    if (i < nlocal) t_react_cell<G>(c, i, tran_xx, l2g, dt, dt_mode, iters, flags);
  }
}

}  // namespace

template <int G>
void tile_launch_variant(const TilePlan &p, const DevTab &h, const double *blob, const DevState &S, double *tran_xx,
                         const int32_t *l2g, long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags,
                         cudaStream_t stream) {
  cudaFuncSetAttribute(k_react_tile<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
  int bps = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_react_tile<G>, p.tt.threads, p.smem_bytes) != cudaSuccess || bps < 1) bps = 1;
  const long long ntiles = (nlocal + p.tt.cpb - 1) / p.tt.cpb;
  const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)p.grid * bps);
  k_react_tile<G><<<grid, p.tt.threads, p.smem_bytes, stream>>>(h, p.tt, blob, p.d_blob, S, tran_xx, l2g, nlocal, dt, dt_mode,
                                                              iters, flags);
}

}  // namespace rxn
