// rxn_lane_dev.cuh — device code of the resident-lane RReact kernel (design: rxn_lane.h).
// A group of G lanes (G = 1, 2 or 4, adjacent lanes of one warp) solves one cell; all per-cell
// arrays live in shared memory in a cell-fastest layout; tables + term streams are staged once per
// persistent CTA.  Reference routines restated (file:line at each site):
// RReact reaction.F90:3322-3511, RTotal :4057-4158, RActivityCoefficients (LAG) :3994-4050,
// RTotalSorbEqSurfCplx1 reaction_surf_complex.F90:658-934, RMultiRateSorption :566-654,
// RKineticMineral reaction_mineral.F90:564-1000, RSolve reaction.F90:4835-4880,
// ludcmp/lubksb utility.F90:393-523.
//
// Work split inside a group: term-stream groups, complexes, activity classes and I/O elements are
// dealt round-robin to the lanes; row i of the Newton system (residual, Jacobian row, LU row
// operations, update) belongs to lane i mod G.  Group-wide values (ionic strength, free-site sums,
// pivot, convergence tests) are xor-butterflies, so every lane of a group holds bit-identical
// copies and the group's control flow is uniform.
//
// The same source is compiled for the host by the CPU-only test harness (RXN_LANE_HOST: one thread per lane,
// butterflies through a barrier) and checked against the oracle in the CPU-only suite.
//
// Deviations from the reference's operation order (REASSOC, all deterministic and <= 1e-14 relative; parity is
// measured in tests/):
//   - sec_molal = exp(lnQK - ln gamma) instead of exp(lnQK)/gamma; ln gamma of a species is the
//     Debye-Hueckel exponent itself instead of log(exp(exponent));
//   - d(total_i)/d(m_j) = (sum_k nu_ik nu_jk sec_molal_k) / m_j instead of one exp(lnQK - ln m_j)/gamma per
//     (complex, species) pair: removes S = sum nspec exps per Newton iteration (202 of ~430 for 300A);
//   - ionic strength and free-site sums are tree reductions over the group;
//   - sorption / multirate derivative blocks are added into J term by term instead of through a dense temporary;
//   - J is assembled as dR_i/d ln m_j (column j times m_j): in the log formulation the reference's
//     (.../m_j)*m_j pair is not executed; the row norms of RSolve use |Jln_ij|/m_j;
//   - long sums of RTotal are split over 4 accumulators (WIDE groups), combined (a0+a1)+(a2+a3);
//   - back substitution multiplies by the stored reciprocal pivot; residual/dt uses a Newton-corrected
//     reciprocal (bit-identical to the division barring double rounding).
#pragma once
#include "rxn_lane.h"

#ifndef RXN_LANE_HOST
#include <cuda_runtime.h>
#define LANE_DEV static __device__ __forceinline__
#define LANE_COLD static __device__ __noinline__
#else
#include <pthread.h>
#define LANE_DEV static inline
#define LANE_COLD static inline
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
#endif

namespace rxn {
namespace lane {

#define RXN_LOG_TO_LN 2.30258509299           /* pflotran_constants.F90:48 (truncated on purpose) */
#define RXN_IDEAL_GAS_CONSTANT 8.31446        /* pflotran_constants.F90:53 */

#ifndef RXN_LANE_HOST
extern __shared__ __align__(16) double tsm[];
#else
static double *tsm = nullptr;                 // one emulated CTA at a time
struct HostGroup {                            // the G host threads of the emulated lane group
  pthread_barrier_t bar;
  double d[8];
  int i[8];
};
static HostGroup *g_hg = nullptr;
static thread_local int g_hl = 0;
#endif

#define TSM2 (reinterpret_cast<double2 *>(tsm))
#define TSB(ob) (*reinterpret_cast<const double *>(reinterpret_cast<const char *>(tsm) + (ob)))   /* byte offset */
#define TD(lt, o) (tsm[(o)])
#define TI(lt, o) (reinterpret_cast<const int *>(tsm)[2 * (lt).blob_dbl + (o)])
#define TI4(lt, o4) (reinterpret_cast<const int4 *>(tsm)[((lt).blob_dbl >> 1) + (o4)])    /* o4 in units of 4 ints */
#define TD2(lt, o2) (reinterpret_cast<const double2 *>(tsm)[(o2)])                          /* o2 in units of 2 doubles */

#define GSL(S, field, row, cell) ((S).f[field][(long long)(row) * (S).ld + (cell)])

// ---------------------------------------------------------------------------------------------
// group primitives.  The G lanes of a cell are 32/G lanes apart in their warp (lane = l*(32/G) + column):
// every half-warp of an 8-byte access and every quarter-warp of a 16-byte access then touches DISTINCT
// cell columns of one element row, so the lanes of a cell never meet in a shared-memory bank (adjacent lanes
// collided on every access: 43 % of the wavefronts of round 1's G = 2 kernel were replays).
#ifndef LANE_ADJACENT
#define LANE_STRIDE(G) (32 / (G))
#else
#define LANE_STRIDE(G) 1
#endif
template <int G> LANE_DEV void grp_sync(unsigned gm) {
#ifndef RXN_LANE_HOST
  if (G > 1) __syncwarp(gm);
#else
  (void)gm;
  if (G > 1) pthread_barrier_wait(&g_hg->bar);
#endif
}
#ifdef RXN_LANE_HOST
// butterfly in the order of the device: round o = G/2 .. 1, v_l <- op(v_l, v_{l^o})
template <int G, class OP> static inline double host_butterfly(double v, OP op) {
  if (G == 1) return v;
  for (int o = G / 2; o > 0; o >>= 1) {
    g_hg->d[g_hl] = v;
    pthread_barrier_wait(&g_hg->bar);
    const double other = g_hg->d[g_hl ^ o];
    pthread_barrier_wait(&g_hg->bar);
    v = op(v, other);
  }
  return v;
}
#endif
template <int G> LANE_DEV double grp_sum(double v, unsigned gm) {
#ifndef RXN_LANE_HOST
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o * LANE_STRIDE(G));
  return v;
#else
  (void)gm;
  return host_butterfly<G>(v, [](double a, double b) { return a + b; });
#endif
}
template <int G> LANE_DEV double grp_max(double v, unsigned gm) {
#ifndef RXN_LANE_HOST
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gm, v, o * LANE_STRIDE(G)));
  return v;
#else
  (void)gm;
  return host_butterfly<G>(v, [](double a, double b) { return fmax(a, b); });
#endif
}
template <int G> LANE_DEV double grp_min(double v, unsigned gm) {
#ifndef RXN_LANE_HOST
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(gm, v, o * LANE_STRIDE(G)));
  return v;
#else
  (void)gm;
  return host_butterfly<G>(v, [](double a, double b) { return fmin(a, b); });
#endif
}
template <int G> LANE_DEV bool grp_any(bool p, unsigned gm) {
#ifndef RXN_LANE_HOST
  if (G == 1) return p;
  return (__ballot_sync(gm, p) & gm) != 0u;
#else
  (void)gm;
  return host_butterfly<G>(p ? 1.0 : 0.0, [](double a, double b) { return fmax(a, b); }) != 0.0;
#endif
}
// pivot of ludcmp over the group: maximum value, ties -> largest index ("last maximum")
template <int G> LANE_DEV void grp_argmax_last(double &best, int &bidx, unsigned gm) {
#ifndef RXN_LANE_HOST
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(gm, best, o * LANE_STRIDE(G));
    const int oi = __shfl_xor_sync(gm, bidx, o * LANE_STRIDE(G));
    if (ob > best || (ob == best && oi > bidx)) { best = ob; bidx = oi; }
  }
#else
  (void)gm;
  if (G == 1) return;
  for (int o = G / 2; o > 0; o >>= 1) {
    g_hg->d[g_hl] = best; g_hg->i[g_hl] = bidx;
    pthread_barrier_wait(&g_hg->bar);
    const double ob = g_hg->d[g_hl ^ o];
    const int oi = g_hg->i[g_hl ^ o];
    pthread_barrier_wait(&g_hg->bar);
    if (ob > best || (ob == best && oi > bidx)) { best = ob; bidx = oi; }
  }
#endif
}

// per-lane context: registers.  R = rows of the Newton system owned by a lane.
template <int N, int G>
struct Lane {
  static constexpr int R = (N + G - 1) / G;
  int l;                  // lane within the group
  unsigned gm;            // lane mask of the group
  int s;                  // cell column within the CTA
  int jb;                 // double2 index of this cell's J(0,0) pair
  int vm, vlna, vlng, vsm, vtot, vscr, vsc, vfree, vmnrl, vr0, vseq, vlk;   // double index of element 0 of each slot
  double fix[R];          // fixed accumulation of the owned rows l, l+G, ... (reaction.F90:3370-3400)
  double den_kg_per_L, psv, psvd, v_t, volume, porosity, soil_density, temp, ln_act_h2o, den_kg;
  long long item, cell;
  int iter, flags;
};

template <int N, int CPB, int G>
LANE_DEV void lane_bind(const LaneTab &lt, Lane<N, G> &c, int s, int l, unsigned gm) {
  c.s = s; c.l = l; c.gm = gm;
  c.jb = lt.o_J2 + s;
  const int v = lt.o_vec + s;
  c.vm = v + lt.s_m * CPB; c.vlna = v + lt.s_lna * CPB; c.vlng = v + lt.s_lng * CPB; c.vsm = v + lt.s_sm * CPB;
  c.vtot = v + lt.s_tot * CPB; c.vscr = v + lt.s_scr * CPB; c.vsc = v + lt.s_sc * CPB; c.vfree = v + lt.s_free * CPB;
  c.vmnrl = v + lt.s_mnrl * CPB; c.vr0 = v + lt.s_r0 * CPB; c.vseq = v + lt.s_seq * CPB; c.vlk = v + lt.s_lk * CPB;
}

// J pair (i, p) / element (i, j) of this cell; j = N is the right-hand side b
#define JP(c, i, p) TSM2[(c).jb + ((i) * LDJ2 + (p)) * CPB]
#define JE(c, i, j) tsm[2 * ((c).jb + ((i) * LDJ2 + ((j) >> 1)) * CPB) + ((j) & 1)]

LANE_COLD double c_exp(double x) { return exp(x); }
LANE_COLD double c_log(double x) { return log(x); }
LANE_COLD double c_pow_slow(double x, double y) { return pow(x, y); }
LANE_DEV double c_pow(double x, double y) { return y == 1.0 ? x : c_pow_slow(x, y); }   // pow(x, 1) == x exactly

// ---------------------------------------------------------------------------------------------
// RActivityCoefficients, LAG algorithm — reaction.F90:3994-4050.  One Debye-Hueckel exponent per
// (Z^2, a0) class; ln gamma is kept (never exponentiated inside the Newton loop).
template <int N, int CPB, int G>
LANE_DEV void lane_act_coefs(const LaneTab &lt, Lane<N, G> &c) {
  const int n = lt.n, ncplx = lt.ncplx;
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, psum = 0.0;
#pragma unroll 1
  for (int i = c.l; i < n; i += G) {
    const double mm = tsm[c.vm + i * CPB];
    p0 = fma(mm, TD(lt, lt.d_pz2 + i), p0);
    if (lt.use_act_h2o && i + 1 != lt.h2o_aq_id) psum += mm;
  }
  const int k4 = ncplx & ~3;
#pragma unroll 1
  for (int k = 4 * c.l; k < k4; k += 4 * G) {                  // REASSOC: partial sums, Z^2 premultiplied
    const double s0 = tsm[c.vsm + k * CPB], s1 = tsm[c.vsm + (k + 1) * CPB], s2 = tsm[c.vsm + (k + 2) * CPB],
                 s3 = tsm[c.vsm + (k + 3) * CPB];
    p0 = fma(s0, TD(lt, lt.d_cz2 + k), p0); p1 = fma(s1, TD(lt, lt.d_cz2 + k + 1), p1);
    p2 = fma(s2, TD(lt, lt.d_cz2 + k + 2), p2); p3 = fma(s3, TD(lt, lt.d_cz2 + k + 3), p3);
    if (lt.use_act_h2o) psum += (s0 + s1) + (s2 + s3);
  }
#pragma unroll 1
  for (int k = k4 + c.l; k < ncplx; k += G) {
    const double s0 = tsm[c.vsm + k * CPB];
    p1 = fma(s0, TD(lt, lt.d_cz2 + k), p1);
    if (lt.use_act_h2o) psum += s0;
  }
  const double I = 0.5 * grp_sum<G>((p0 + p1) + (p2 + p3), c.gm);
  const double sqrt_I = sqrt(I);
  if (c.l == 0) tsm[c.vlng] = 0.0;
#pragma unroll 1
  for (int q = 1 + c.l; q < lt.ncls; q += G)
    tsm[c.vlng + q * CPB] =
        (-TD(lt, lt.d_cls_z2 + q) * sqrt_I * lt.debyeA / (1.0 + TD(lt, lt.d_cls_a0 + q) * lt.debyeB * sqrt_I) + lt.debyeBdot * I) * RXN_LOG_TO_LN;
  if (lt.use_act_h2o) {                                        // :4043-4050
    const double a = 1.0 - 0.017 * grp_sum<G>(psum, c.gm);
    c.ln_act_h2o = (a > 0.0) ? c_log(a) : 0.0;
    if (c.l == 0) tsm[c.vlna + (lt.n + 1) * CPB] = c.ln_act_h2o;
  }
  grp_sync<G>(c.gm);
}

// ---------------------------------------------------------------------------------------------
// term streams: 4 accumulators advance together, one {coef[4], offset[4]} record per step
#ifndef LANE_UNROLL
#define LANE_UNROLL 2
#endif
LANE_DEV void lane_run_group(const LaneTab &lt, int s, int c0, int o0, int nsteps, double &a0, double &a1, double &a2, double &a3) {
  constexpr int U = LANE_UNROLL;
  const char *cell8 = reinterpret_cast<const char *>(tsm + s);   // a gather address is this + the record's byte offset
#pragma unroll U
  for (int q = 0; q < nsteps; ++q) {
    const double2 ca = TD2(lt, (c0 + 1 + q) * 2), cb = TD2(lt, (c0 + 1 + q) * 2 + 1);
    const int4 of = TI4(lt, o0 + q);
    a0 = fma(ca.x, *reinterpret_cast<const double *>(cell8 + of.x), a0);
    a1 = fma(ca.y, *reinterpret_cast<const double *>(cell8 + of.y), a1);
    a2 = fma(cb.x, *reinterpret_cast<const double *>(cell8 + of.z), a2);
    a3 = fma(cb.y, *reinterpret_cast<const double *>(cell8 + of.w), a3);
  }
}

// ln a_i = ln m_i + ln gamma_i, then sec_molal_k = exp(lnQK_k - ln gamma_k)   (RTotal, reaction.F90:4090-4122)
template <int N, int CPB, int G>
LANE_DEV void lane_speciate(const LaneTab &lt, Lane<N, G> &c) {
  const int n = lt.n, s = c.s;
#pragma unroll 2
  for (int i = c.l; i < n; i += G)
    tsm[c.vlna + i * CPB] = log(tsm[c.vm + i * CPB]) + tsm[c.vlng + TI(lt, lt.i_pcls + i) * CPB];
  grp_sync<G>(c.gm);
#pragma unroll 1
  for (int g = c.l; g < lt.spec.ng; g += G) {
    const int4 hd = TI4(lt, (lt.spec.g0 >> 2) + 2 * g), h2 = TI4(lt, (lt.spec.g0 >> 2) + 2 * g + 1);
    const int c0 = hd.x, o0 = hd.y, nsteps = hd.z, cb = h2.x >> 2;
    const int4 m0 = TI4(lt, cb), m1 = TI4(lt, cb + 1), m2 = TI4(lt, cb + 2), m3 = TI4(lt, cb + 3);
    double a0, a1, a2, a3;
    if (lt.percell_logK) {
      a0 = tsm[c.vlk + (m0.z < 0 ? 0 : m0.z) * CPB]; a1 = tsm[c.vlk + (m1.z < 0 ? 0 : m1.z) * CPB];
      a2 = tsm[c.vlk + (m2.z < 0 ? 0 : m2.z) * CPB]; a3 = tsm[c.vlk + (m3.z < 0 ? 0 : m3.z) * CPB];
    } else {
      const double2 ia = TD2(lt, c0 * 2), ib = TD2(lt, c0 * 2 + 1);
      a0 = ia.x; a1 = ia.y; a2 = ib.x; a3 = ib.y;
    }
    lane_run_group(lt, s, c0, o0, nsteps, a0, a1, a2, a3);
    if (lt.gamma_state) {                                      // sec_molal = exp(lnQK)/gamma_k, gamma_k as stored in the state (:4112)
      const double e0 = exp(a0) / tsm[m0.y + s], e1 = exp(a1) / tsm[m1.y + s], e2 = exp(a2) / tsm[m2.y + s], e3 = exp(a3) / tsm[m3.y + s];
      tsm[m0.x + s] = e0; tsm[m1.x + s] = e1; tsm[m2.x + s] = e2; tsm[m3.x + s] = e3;
    } else {                                                   // REASSOC: exp(lnQK)/gamma_k -> exp(lnQK - ln gamma_k)
      const double e0 = exp(a0 - tsm[m0.y + s]), e1 = exp(a1 - tsm[m1.y + s]), e2 = exp(a2 - tsm[m2.y + s]),
                   e3 = exp(a3 - tsm[m3.y + s]);
      tsm[m0.x + s] = e0; tsm[m1.x + s] = e1; tsm[m2.x + s] = e2; tsm[m3.x + s] = e3;
    }
  }
  grp_sync<G>(c.gm);
}

// plan A: tot_i <- sum_k nu_ik sm_k ; plan B: Jln_ij = Jln_ji <- (sum_k nu_ik nu_jk sm_k) * scale
template <int N, int CPB, int G, bool JAC>
LANE_DEV void lane_plan(const LaneTab &lt, Lane<N, G> &c, double scale) {
  const LaneStream S = JAC ? lt.planB : lt.planA;
  const int s = c.s, s2 = 2 * c.s;
#pragma unroll 1
  for (int g = c.l; g < S.ng; g += G) {
    const int4 hd = TI4(lt, (S.g0 >> 2) + 2 * g), h2 = TI4(lt, (S.g0 >> 2) + 2 * g + 1);
    const int c0 = hd.x, o0 = hd.y, nsteps = hd.z, mode = hd.w, cb = h2.x >> 2;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    lane_run_group(lt, s, c0, o0, nsteps, a0, a1, a2, a3);
    if (mode == LANE_WIDE) {
      const double a = (a0 + a1) + (a2 + a3);
      const int4 d = TI4(lt, cb);
      if (!JAC) tsm[d.x + s] = a;
      else { const double v = a * scale; tsm[d.x + s2] = v; tsm[d.y + s2] = v; }
    } else if (!JAC) {
      const int4 d = TI4(lt, cb);
      tsm[d.x + s] = a0; tsm[d.y + s] = a1; tsm[d.z + s] = a2; tsm[d.w + s] = a3;
    } else {
      const int4 d0 = TI4(lt, cb), d1 = TI4(lt, cb + 1);
      const double v0 = a0 * scale, v1 = a1 * scale, v2 = a2 * scale, v3 = a3 * scale;
      tsm[d0.x + s2] = v0; tsm[d0.y + s2] = v0; tsm[d0.z + s2] = v1; tsm[d0.w + s2] = v1;
      tsm[d1.x + s2] = v2; tsm[d1.y + s2] = v2; tsm[d1.z + s2] = v3; tsm[d1.w + s2] = v3;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RTotalSorbEqSurfCplx1 — reaction_surf_complex.F90:658-934, one surface complexation reaction.
// Complexes are dealt to the lanes; totals / Jacobian rows go to the lane that owns the row.
//   target_i (tsm[tb + i*ts]: column N of J for equilibrium reactions, the S_eq vector for multirate ones)
//            += total sorbed of primary i
//   addJ: Jln(i, j) += fac * d(total_sorb_i)/d ln m_j   (REASSOC: term by term, no dense temporary)
template <int N, int CPB, int G>
LANE_DEV void lane_srf_rxn(const LaneTab &lt, Lane<N, G> &c, const DevState &S, int irxn, double fac, bool addJ, bool store_conc, int tb,
                           int ts) {
  constexpr int LDJ2 = (N + 2) / 2;
  const int n = lt.n;
  const double tol = 1.0e-12;
  const int c0 = TI(lt, lt.i_rxn_cptr + irxn), c1 = TI(lt, lt.i_rxn_cptr + irxn + 1);
  const int nlk0 = lt.d_nlk + lt.ncplx + lt.nkin;
  double free_site_conc = tsm[c.vfree + irxn * CPB];
  double site_density;
  const int surf_type = TI(lt, lt.i_rxn_surf_type + irxn);
  const double dens = TD(lt, lt.d_rxn_density + irxn);
  if (surf_type == RXN_MINERAL_SURFACE) site_density = dens * tsm[c.vmnrl + (TI(lt, lt.i_rxn_to_surf + irxn) - 1) * CPB];
  else if (surf_type == RXN_ROCK_SURFACE) site_density = dens * c.soil_density * (1.0 - c.porosity);
  else site_density = dens;
  if (site_density < 1.0e-40) return;                         // :749 (uniform over the group)
  const int stoich_flag = TI(lt, lt.i_rxn_flag + irxn);
  bool one_more = false;
  int num_iterations = 0;
  double damping_factor = 1.0;
  grp_sync<G>(c.gm);                                           // every lane has read the warm-start value
#pragma unroll 1
  for (;;) {                                                  // :760-829
    num_iterations = num_iterations + 1;
    const double ln_free_site = c_log(free_site_conc);
    double part = 0.0, part2 = 0.0;
#pragma unroll 1
    for (int j = c0 + c.l; j < c1; j += G) {
      const int icplx = TI(lt, lt.i_rxn_cid + j);
      double lnQK = lt.percell_logK ? tsm[c.vlk + (lt.ncplx + lt.nkin + icplx) * CPB] : TD(lt, nlk0 + icplx);
      const double sh2o = TD(lt, lt.d_sh2o + icplx), site_st = TD(lt, lt.d_site_st + icplx);
      if (sh2o != 0.0) lnQK = lnQK + sh2o * c.ln_act_h2o;
      lnQK = lnQK + site_st * ln_free_site;
      const int p1 = TI(lt, lt.i_sptr + icplx + 1);
#pragma unroll 1
      for (int p = TI(lt, lt.i_sptr + icplx); p < p1; ++p) lnQK = lnQK + TD(lt, lt.d_sst + p) * tsm[c.vlna + TI(lt, lt.i_sid + p) * CPB];
      const double sc = c_exp(lnQK);
      tsm[c.vsc + (j - c0) * CPB] = sc;
      part += site_st * sc;
      part2 += site_st * sc / free_site_conc;
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
      if (num_iterations > 100000) { c.flags |= RXN_FLAG_CAPPED; one_more = true; }   // reference would spin
    } else {
      total = total / free_site_conc;
      free_site_conc = site_density / total;
      one_more = true;
    }
  }
  grp_sync<G>(c.gm);
  if (c.l == 0) tsm[c.vfree + irxn * CPB] = free_site_conc;
  if (store_conc) {
#pragma unroll 1
    for (int j = c0 + c.l; j < c1; j += G) GSL(S, RXN_F_EQSRFCPLX_CONC, TI(lt, lt.i_rxn_cid + j), c.cell) += tsm[c.vsc + (j - c0) * CPB];
  }
  const double inv_free = 1.0 / free_site_conc;
  // The loops below walk the (complex, species) list once; the lane that owns row sid[p] takes the entry
  // (per row the entries arrive in the reference's order: complexes ascending, species ascending).
  if (addJ) {                                                  // :838-866 (tempreal redundantly per lane: few complexes)
#pragma unroll 1
    for (int row = c.l; row < n; row += G) tsm[c.vscr + row * CPB] = 0.0;
    double tempreal = 0.0;
#pragma unroll 1
    for (int j = c0; j < c1; ++j) {
      const int icplx = TI(lt, lt.i_rxn_cid + j);
      const double sc = tsm[c.vsc + (j - c0) * CPB], site_st = TD(lt, lt.d_site_st + icplx);
      const int p1 = TI(lt, lt.i_sptr + icplx + 1);
#pragma unroll 1
      for (int p = TI(lt, lt.i_sptr + icplx); p < p1; ++p) {
        const int row = TI(lt, lt.i_sid + p);
        if (G > 1 && (row & (G - 1)) != c.l) continue;
        const int o = c.vscr + row * CPB;
        tsm[o] = tsm[o] + TD(lt, lt.d_sst + p) * site_st * sc;
      }
      tempreal = tempreal + site_st * site_st * sc;
    }
    tempreal = tempreal / free_site_conc;
    tempreal = tempreal + 1.0;
#pragma unroll 1
    for (int row = c.l; row < n; row += G) tsm[c.vscr + row * CPB] = -tsm[c.vscr + row * CPB] / tempreal;   // dSx/d ln m_row
    grp_sync<G>(c.gm);
  }
#pragma unroll 1
  for (int j = c0; j < c1; ++j) {                              // :872-931
    const int icplx = TI(lt, lt.i_rxn_cid + j);
    const double sc = tsm[c.vsc + (j - c0) * CPB];
    const int p0 = TI(lt, lt.i_sptr + icplx), p1 = TI(lt, lt.i_sptr + icplx + 1);
    const double nui_Si_over_Sx = TD(lt, lt.d_site_st + icplx) * sc * inv_free;
#pragma unroll 1
    for (int p = p0; p < p1; ++p) {
      const int row = TI(lt, lt.i_sid + p);
      if (G > 1 && (row & (G - 1)) != c.l) continue;
      const double stp = TD(lt, lt.d_sst + p);
      const int o = tb + row * ts;
      tsm[o] = tsm[o] + stp * sc;
      if (addJ) {
#pragma unroll 1
        for (int q = p0; q < p1; ++q) {
          const int jc = TI(lt, lt.i_sid + q);
          const double tr = TD(lt, lt.d_sst + q) * sc + nui_Si_over_Sx * tsm[c.vscr + jc * CPB];
          JE(c, row, jc) = JE(c, row, jc) + (stp * tr) * fac;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RKineticMineral — reaction_mineral.F90:564-1000 (tables with prefactors use the cooperative kernel).
// Every lane evaluates the (few) rate laws redundantly; the owner lane of a primary adds its
// residual entry and its Jacobian row.
template <int N, int CPB, int G>
LANE_DEV void lane_kinetic_mineral(const LaneTab &lt, Lane<N, G> &c) {
  constexpr int LDJ2 = (N + 2) / 2;
#pragma unroll 1
  for (int imnrl = 0; imnrl < lt.nkin; ++imnrl) {
    double rate_out = 0.0;
    do {
      double lnQK = lt.percell_logK ? tsm[c.vlk + (lt.ncplx + imnrl) * CPB] : TD(lt, lt.d_nlk + lt.ncplx + imnrl);
      const double h2ost = TD(lt, lt.d_kh2o + imnrl);
      if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
      const int p0 = TI(lt, lt.i_kptr + imnrl), p1 = TI(lt, lt.i_kptr + imnrl + 1);
#pragma unroll 1
      for (int p = p0; p < p1; ++p) lnQK = lnQK + TD(lt, lt.d_kst + p) * tsm[c.vlna + TI(lt, lt.i_kid + p) * CPB];
      double QK;
      if (lnQK <= 6.90776) QK = c_exp(lnQK); else QK = 1.0e3;
      const double k_scale = lt.has_scale ? TD(lt, lt.d_k_scale + imnrl) : 1.0;
      const double k_Temkin = lt.has_Temkin ? TD(lt, lt.d_k_Temkin + imnrl) : 1.0;
      const double k_power = lt.has_power ? TD(lt, lt.d_k_power + imnrl) : 1.0;
      const double k_lim = TD(lt, lt.d_k_lim + imnrl);
      const double k_aff = TD(lt, lt.d_k_aff + imnrl);
      double affinity_factor;
      if (lt.has_Temkin) {
        if (lt.has_scale) affinity_factor = 1.0 - c_pow(QK, 1.0 / (k_scale * k_Temkin));
        else affinity_factor = 1.0 - c_pow(QK, 1.0 / k_Temkin);
      } else if (lt.has_scale) {
        affinity_factor = 1.0 - c_pow(QK, 1.0 / k_scale);
      } else {
        affinity_factor = 1.0 - QK;
      }
      const double sign_ = copysign(1.0, affinity_factor);
      const double volfrac = tsm[c.vmnrl + imnrl * CPB];
      if (!(volfrac > 0 || sign_ < 0.0)) break;                 // :723
      if (k_aff > 0.0) {
        if (sign_ < 0.0 && QK < k_aff) break;
      }
      if (k_lim > 0.0) affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / k_lim);
      double arrhenius_factor = 1.0;
      const double Ea = TD(lt, lt.d_k_Ea + imnrl);
      if (Ea > 0.0) arrhenius_factor = c_exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
      const double sum_prefactor_rate = TD(lt, lt.d_k_rate + imnrl) * arrhenius_factor;
      double Im_const = -tsm[c.vmnrl + (lt.nkin + imnrl) * CPB], Im;
      if (lt.has_scale) Im_const = Im_const / k_scale;
      if (lt.has_power) Im = Im_const * sign_ * c_pow(fabs(affinity_factor), k_power) * sum_prefactor_rate;
      else Im = Im_const * sign_ * fabs(affinity_factor) * sum_prefactor_rate;
      rate_out = Im;

      Im_const = Im_const * c.volume;
      Im = Im * c.volume;
      double dIm_dQK;
      if (lt.has_power) dIm_dQK = -Im * k_power / fabs(affinity_factor);
      else dIm_dQK = -Im_const * sum_prefactor_rate;
      if (lt.has_Temkin) {
        if (lt.has_scale) dIm_dQK = dIm_dQK * (1.0 / (k_scale * k_Temkin)) / QK * (1.0 - affinity_factor);
        else dIm_dQK = dIm_dQK * (1.0 / k_Temkin) / QK * (1.0 - affinity_factor);
      } else if (lt.has_scale) {
        dIm_dQK = dIm_dQK * (1.0 / k_scale) / QK * (1.0 - affinity_factor);
      }
      const double den = (k_lim <= 0.0) ? 1.0 : 1.0 + (1.0 - affinity_factor) / k_lim;
#pragma unroll 1
      for (int p = p0; p < p1; ++p) {
        const int ip = TI(lt, lt.i_kid + p);
        if (G > 1 && (ip & (G - 1)) != c.l) continue;            // owner lane of primary ip
        const double stp = TD(lt, lt.d_kst + p);
        JE(c, ip, N) = JE(c, ip, N) + stp * Im;
#pragma unroll 1
        for (int q = p0; q < p1; ++q) {
          const int jcomp = TI(lt, lt.i_kid + q);
          const double dQK_dCj = TD(lt, lt.d_kst + q) * QK;     // d/d ln m_j: the reference's exp(-ln m_j) factor is not applied
          const double dQK_dmj = dQK_dCj * c.den_kg * 1.0e-3;
          if (k_lim <= 0.0) JE(c, ip, jcomp) = JE(c, ip, jcomp) + stp * dIm_dQK * dQK_dmj;
          else JE(c, ip, jcomp) = JE(c, ip, jcomp) + stp * dIm_dQK * (1.0 + QK / k_lim / den) * dQK_dmj / den;
        }
      }
    } while (false);
    if (c.l == 0) tsm[c.vmnrl + (2 * lt.nkin + imnrl) * CPB] = rate_out;   // :575 (zeroed) / :816
  }
}

// ---------------------------------------------------------------------------------------------
// RSolve (reaction.F90:4835-4880) + ludcmp/lubksb (utility.F90:393-523) on [Jln | b] in shared memory.
// Row i belongs to lane i mod G.  Right-looking elimination with the pivot row in registers; per element
// the same a(i,j) -= a(i,k)*a(k,j), k ascending, as Crout; pivot = last maximum of vv(i)*|a(i,k)|, i >= k.
// b (column N) goes through the elimination (= forward substitution of lubksb); row-oriented back
// substitution in the reference's order.  The k loop is a run-time loop; the column range of step k is
// selected once per step (dispatch on the first pair) so the row loop has no per-row control.  The solution
// replaces b.  Returns 1 if a row is all zero (reference: MPI_Abort).
template <int N, int CPB, int G, int P0>
LANE_DEV void lane_lu_elim(const Lane<N, G> &c, int i0, int ek, int kodd, double dum, const double2 *pr, double &best, int &imax) {
  constexpr int LDJ2 = (N + 2) / 2;
#ifndef LANE_ELIM_UNROLL
#define LANE_ELIM_UNROLL 1
#endif
  constexpr int UE = LANE_ELIM_UNROLL;
#pragma unroll UE
  for (int i = i0; i < N; i += G) {
    const double lik = tsm[ek + i * (2 * LDJ2 * CPB)] * dum;
    const double vvi = tsm[c.vscr + i * CPB];
    const int ri = c.jb + i * (LDJ2 * CPB);
    double nxt = 0.0;
#pragma unroll
    for (int p = P0; p < LDJ2; ++p) {                          // whole pairs: a stale column <= k may be rewritten, it is dead
      double2 a = TSM2[ri + p * CPB];
      a.x = a.x - lik * pr[p].x;
      a.y = a.y - lik * pr[p].y;
      TSM2[ri + p * CPB] = a;
      if (p == P0) nxt = kodd ? a.x : a.y;                     // column k+1 of this row: even k+1 -> .x of pair (k+1)/2
    }
    // pivot search of the next step (ludcmp :440-449) on the fly: rows > k are exactly the candidates of step k+1
    const double cand = vvi * fabs(nxt);
    if (cand >= best) { best = cand; imax = i; }
  }
}
template <int N, int CPB, int G, int P>
LANE_DEV void lane_lu_elim_from(const Lane<N, G> &c, int pe, int i0, int rk, int ek, int kodd, double dum, double &best, int &imax) {
  constexpr int LDJ2 = (N + 2) / 2;
  if constexpr (P < LDJ2) {
    if (pe == P) {
      double2 pr[LDJ2];
#pragma unroll
      for (int p = P; p < LDJ2; ++p) pr[p] = TSM2[rk + p * CPB];
      lane_lu_elim<N, CPB, G, P>(c, i0, ek, kodd, dum, pr, best, imax);
    } else {
      lane_lu_elim_from<N, CPB, G, P + 1>(c, pe, i0, rk, ek, kodd, dum, best, imax);
    }
  }
}

template <int N, int CPB, int G>
LANE_DEV int lane_rsolve(const LaneTab &lt, Lane<N, G> &c) {
  constexpr int LDJ2 = (N + 2) / 2;
  const double tiny = 1.0e-20;
  const bool use_log = lt.use_log != 0;
  bool zero = false;
  double best = -1.0;                                          // running pivot search: value / row of the next step
  int imax = -1;
  // rows scaled by 1/max(1, max_j |J_ij|), J_ij = Jln_ij/m_j (:4851-4858), log form: times m_j (:4866-4870);
  // implicit-scaling factors vv(i) = 1/max_j |a(i,j)| (ludcmp :413-425) -> scratch.  1/m_j: each lane computes a
  // share, all lanes keep the N values in registers.
  {
    double invm[N];
    if (G == 1) {
#pragma unroll
      for (int j = 0; j < N; ++j) invm[j] = 1.0 / tsm[c.vm + j * CPB];
    } else {
#pragma unroll 1
      for (int j = c.l; j < N; j += G) tsm[c.vscr + j * CPB] = 1.0 / tsm[c.vm + j * CPB];
      grp_sync<G>(c.gm);
#pragma unroll
      for (int j = 0; j < N; ++j) invm[j] = tsm[c.vscr + j * CPB];
      grp_sync<G>(c.gm);
    }
#pragma unroll 1
    for (int i = c.l; i < N; i += G) {
      double2 r[LDJ2];
#pragma unroll
      for (int p = 0; p < LDJ2; ++p) r[p] = JP(c, i, p);
      double mx = 0.0, mraw = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double av = fabs((j & 1) ? r[j >> 1].y : r[j >> 1].x), v = av * invm[j];
        if (v > mx) mx = v;
        if (av > mraw) mraw = av;
      }
      const double norm = 1.0 / ((mx > 1.0) ? mx : 1.0);
#pragma unroll
      for (int j = 0; j <= N; ++j) {
        double v = (j & 1) ? r[j >> 1].y : r[j >> 1].x;
        if (j < N && !use_log) v = v * invm[j];
        v = v * norm;
        if (j & 1) r[j >> 1].y = v; else r[j >> 1].x = v;
      }
#pragma unroll
      for (int p = 0; p < LDJ2; ++p) JP(c, i, p) = r[p];
      // max_j |a(i,j)|: rounding is monotone, so in the log form it is |.|max of the unscaled row times norm
      const double aamax = use_log ? mraw * norm : mx * norm;
      if (aamax <= 0.0) zero = true;
      const double vvi = 1.0 / aamax;
      tsm[c.vscr + i * CPB] = vvi;
      const double cand = vvi * fabs(r[0].x);                   // pivot search of step 0 (ludcmp :440-449)
      if (cand >= best) { best = cand; imax = i; }
    }
  }
  if (grp_any<G>(zero, c.gm)) return 1;
  grp_sync<G>(c.gm);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const int ek = 2 * (c.jb + (k >> 1) * CPB) + (k & 1);       // element (0, k); row stride 2*LDJ2*CPB
    grp_argmax_last<G>(best, imax, c.gm);
    if (imax < 0) imax = k;
    const int rk = c.jb + k * (LDJ2 * CPB), ri = c.jb + imax * (LDJ2 * CPB);
    if (imax != k) {
      // swap rows k and imax from the pair holding column k on (columns left of it are never read again)
#pragma unroll 1
      for (int p = (k >> 1) + c.l; p < LDJ2; p += G) {
        const double2 a = TSM2[ri + p * CPB], b = TSM2[rk + p * CPB];
        TSM2[ri + p * CPB] = b;
        TSM2[rk + p * CPB] = a;
      }
      if (c.l == 0) tsm[c.vscr + imax * CPB] = tsm[c.vscr + k * CPB];
    }
    grp_sync<G>(c.gm);
    double piv = tsm[ek + k * (2 * LDJ2 * CPB)];
    if (piv == 0.0) piv = tiny;
    const double dum = 1.0 / piv;
    if (c.l == 0) tsm[c.vscr + k * CPB] = dum;                 // vv(k) is dead: keep 1/a(k,k) for the back substitution
    int i1 = c.l;                                               // first owned row > k
    if (i1 <= k) i1 += ((k - i1) / G + 1) * G;
    best = -1.0; imax = -1;
    lane_lu_elim_from<N, CPB, G, 0>(c, (k + 1) >> 1, i1, rk, ek, k & 1, dum, best, imax);
    grp_sync<G>(c.gm);
  }
  if (c.l == 0) {
    double x[N];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {                         // lubksb :511-520 (REASSOC: times 1/a(i,i))
      double2 r[LDJ2];
#pragma unroll
      for (int p = ((i + 1) >> 1); p < LDJ2; ++p) r[p] = JP(c, i, p);
      double sum = (N & 1) ? r[N >> 1].y : r[N >> 1].x;
#pragma unroll
      for (int j = i + 1; j < N; ++j) sum = sum - ((j & 1) ? r[j >> 1].y : r[j >> 1].x) * x[j];
      x[i] = sum * tsm[c.vscr + i * CPB];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) JE(c, i, N) = x[i];
  }
  grp_sync<G>(c.gm);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// lane life cycle: load a cell -> trips (one Newton iteration each) -> finish (closing RTAuxVarCompute + write back)

// x / d with r = RN(1/d): RN(q + fma(-d, q, x) r), q = RN(x r), is the correctly rounded quotient (Markstein)
LANE_DEV double hpt_div(double x, double d, double r) {
#ifndef RXN_LANE_HOST
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
#else
  (void)r;
  return x / d;
#endif
}
// RUpdateTempDependentCoefs reaction.F90:5433-5524 -> -logK*LOG_TO_LN of this cell's T (and P)
template <int CPB, int G>
LANE_COLD void lane_percell_logK(int l, int ncoef, int logK_mode, int vlk, double temp, double pres, const double *blob_d, DSpec s0,
                                 DSpec s1, DSpec s2) {
  const double tk = temp + 273.15;
  // the cell's T, P terms of the hpt fit, once per cell (the same values the reference forms inside every evaluation)
  const double tr = tk / 273.15, pr = pres / 1.0e7;
  double logtr = 0.0, sqtr = 0.0, itr = 0.0, ipr = 0.0;
  if (logK_mode == RXN_LOGK_HPT) { logtr = log(tr) / log(10.0); sqtr = sqrt(tr); itr = 1.0 / tr; ipr = 1.0 / pr; }
  int o0 = 0;
#pragma unroll 1
  for (int q = 0; q < 3; ++q) {
    const DSpec sp = q == 0 ? s0 : q == 1 ? s1 : s2;
#pragma unroll 1
    for (int r = l; r < sp.n; r += G) {
      double lk;
      const bool fixed = sp.o_coef < 0 || (q == 2 && logK_mode == RXN_LOGK_HPT);   // :5517-5521: hpt not applied to srfcplx
      if (fixed) lk = blob_d[sp.o_logK + r];
      else {
        const double *cf = blob_d + sp.o_coef + r * ncoef;
        if (logK_mode == RXN_LOGK_HPT) {                      // reaction_aux.F90:1529-1571
          // the reference's expression term by term; divisions by tr / pr as Markstein-corrected products with the reciprocal
          // (hpt_div: the correctly rounded quotient, i.e. the division's own bits - the fit's terms cancel, so none may change)
          lk = cf[0] + cf[1] * tr + hpt_div(cf[2], tr, itr) + cf[3] * logtr + cf[4] * tr * tr + hpt_div(hpt_div(cf[5], tr, itr), tr, itr) +
               cf[6] * sqtr + cf[7] * pr + cf[8] * pr * tr + hpt_div(cf[9] * pr, tr, itr) + cf[10] * pr * logtr + hpt_div(cf[11], pr, ipr) +
               hpt_div(cf[12], pr, ipr) * tr + hpt_div(hpt_div(cf[13], pr, ipr), tr, itr) + cf[14] * pr * pr + cf[15] * pr * pr * tr +
               hpt_div(cf[16] * pr * pr, tr, itr);
        } else {                                              // reaction_aux.F90:1461-1488
          lk = cf[0] * log(tk) + cf[1] + cf[2] * tk + cf[3] / tk + cf[4] / (tk * tk);
        }
      }
      tsm[vlk + (o0 + r) * CPB] = -lk * RXN_LOG_TO_LN;
    }
    o0 += sp.n;
  }
}

template <int N, int CPB, int G>
LANE_DEV void lane_load(const LaneTab &lt, Lane<N, G> &c, const DevState &S, const double *blob_d, const int *blob_i, const DevTab &h,
                        long long item, long long cell, const double *tran_xx, double tran_dt) {
  constexpr int R = Lane<N, G>::R;
  const int n = lt.n;
  c.item = item; c.cell = cell;
  c.flags = 0; c.iter = 0;
  c.ln_act_h2o = GSL(S, RXN_F_LN_ACT_H2O, 0, cell);
  c.den_kg = GSL(S, RXN_F_DEN_KG, 0, cell);
  c.temp = GSL(S, RXN_F_TEMP, 0, cell);
  c.volume = GSL(S, RXN_F_VOLUME, 0, cell);
  c.porosity = GSL(S, RXN_F_POROSITY, 0, cell);
  c.soil_density = GSL(S, RXN_F_SOIL_PARTICLE_DENSITY, 0, cell);
  const double sat = GSL(S, RXN_F_SAT, 0, cell);
  c.psv = c.porosity * sat * 1000.0 * c.volume;
  c.psvd = c.porosity * sat * 1000.0 * c.volume / tran_dt;                     // :5189
  c.v_t = c.volume / tran_dt;                                                  // :4590
  c.den_kg_per_L = c.den_kg * 1.0 * 1.0e-3;
  {
    // all row loads first (clamped row index: no control flow between them), then the stores
    double pm[R], xx[R], ts[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = c.l + r * G, ic = i < n ? i : n - 1;
      pm[r] = GSL(S, RXN_F_PRI_MOLAL, ic, cell);
      xx[r] = tran_xx[item * n + ic];
      ts[r] = lt.neqsorb > 0 ? GSL(S, RXN_F_TOTAL_SORB_EQ, ic, cell) : 0.0;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = c.l + r * G;
      if (i < n) {
        tsm[c.vm + i * CPB] = pm[r];
        double fx = c.psv * xx[r];                               // :3370, RTAccumulation :5072-5148
        if (lt.neqsorb > 0) fx = fx + ts[r] * c.volume;          // RAccumulationSorb :4539-4568
        c.fix[r] = fx;
      } else {
        // padding row of the shape: m = 1, no complexes -> total = den, residual = psv*den - fix = 0 exactly,
        // Jln_ii = den*psvd: the row stays decoupled and its Newton update is 0
        if (i < N) tsm[c.vm + i * CPB] = 1.0;
        c.fix[r] = c.psv * ((1.0 + 0.0) * c.den_kg_per_L);
      }
    }
  }
  if (c.l == 0) {
    tsm[c.vlna + n * CPB] = 0.0;
    tsm[c.vlna + (n + 1) * CPB] = c.ln_act_h2o;
    tsm[c.vsm + lt.ncplx * CPB] = 0.0;
  }
  if (lt.act_off) {                                            // ln gamma from the state, one class per species
#pragma unroll 1
    for (int i = c.l; i < n; i += G) tsm[c.vlng + i * CPB] = c_log(GSL(S, RXN_F_PRI_ACT_COEF, i, cell));
#pragma unroll 1
    for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vlng + (n + k) * CPB] = c_log(GSL(S, RXN_F_SEC_ACT_COEF, k, cell));
  }
  if (!lt.coop_io && !lt.act_off) {
#pragma unroll 4
    for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vsm + k * CPB] = GSL(S, RXN_F_SEC_MOLAL, k, cell);   // lagged, for I
  }
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) tsm[c.vfree + q * CPB] = GSL(S, RXN_F_FREE_SITE_CONC, q, cell);
#pragma unroll 1
  for (int q = c.l; q < lt.nkin; q += G) {                     // read-only inside RReact
    tsm[c.vmnrl + q * CPB] = GSL(S, RXN_F_MNRL_VOLFRAC, q, cell);
    tsm[c.vmnrl + (lt.nkin + q) * CPB] = GSL(S, RXN_F_MNRL_AREA, q, cell);
  }
  if (lt.percell_logK)
    lane_percell_logK<CPB, G>(c.l, lt.ncoef, lt.logK_mode, c.vlk, c.temp, GSL(S, RXN_F_PRES, 0, cell), blob_d, h.cplx, h.kin, h.srf);
  grp_sync<G>(c.gm);
}

// Warp-cooperative part of taking / finishing a cell: the long per-complex arrays move with all W lanes of the warp
// (lane w takes elements w, w+W, ...), whichever group the cell belongs to.  On the host W = 1.
//   in : lagged sec_molal (for the ionic strength, reaction.F90:3994-4010) and, for multirate sorption,
//        R0_i = sum_r k_r/(1+k_r dt) S_r,i (multirate_prepare, rxn_device.cuh; REASSOC: even and odd rates summed
//        separately, then added)
//   out: sec_molal and sec_act_coef
// lagged sec_molal of up to 3 cells at once (their loads overlap)
template <int N, int CPB>
LANE_DEV void lane_coop_in_sm(const LaneTab &lt, const DevState &S, const int (&slot)[3], const long long (&cell)[3], int cnt, int w, int W) {
  if (lt.act_off) return;
#pragma unroll 1
  for (int k0 = w; k0 < lt.ncplx; k0 += 4 * W) {
    double b[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * W;
        b[g][u] = (g < cnt && k < lt.ncplx) ? GSL(S, RXN_F_SEC_MOLAL, k, cell[g]) : 0.0;
      }
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * W;
        if (g < cnt && k < lt.ncplx) tsm[lt.o_vec + (lt.s_sm + k) * CPB + slot[g]] = b[g][u];
      }
  }
}
template <int N, int CPB>
LANE_DEV void lane_coop_in_mr(const LaneTab &lt, const DevState &S, const DevTab &h, const double *blob_d, const int *blob_i, int slot,
                              long long cell, double tran_dt, int w, int W) {
  const int vr0 = lt.o_vec + lt.s_r0 * CPB + slot;
  const int n = lt.n;
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr) {
    const int nrate = blob_i[h.o_mr_nrate + ikr];
    const long long row0 = ((long long)ikr * (h.mr_ld + 1) + 1) * n;
    // lane (i, half): half = 0 even rates, 1 odd rates
    const int H = (W >= 2 * n) ? 2 : 1, per = W / H;
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += per) {
      const int i = i0 + (w % per), half = w / per;
      double acc0 = 0.0, acc1 = 0.0;
      if (i < n && half < H) {
        if (H == 2) {
#pragma unroll 4
          for (int irate = half; irate < nrate; irate += 2) {
            const double rate = blob_d[h.o_mr_rate + ikr * h.mr_ld + irate];
            const double kk = rate / (1.0 + rate * tran_dt);
            acc0 = acc0 + kk * GSL(S, RXN_F_KINMR_TOTAL_SORB, row0 + (long long)irate * n + i, cell);
          }
        } else {
#pragma unroll 1
          for (int irate = 0; irate < nrate; irate += 2) {
            const double rate = blob_d[h.o_mr_rate + ikr * h.mr_ld + irate];
            acc0 = acc0 + rate / (1.0 + rate * tran_dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, row0 + (long long)irate * n + i, cell);
            if (irate + 1 < nrate) {
              const double rate1 = blob_d[h.o_mr_rate + ikr * h.mr_ld + irate + 1];
              acc1 = acc1 + rate1 / (1.0 + rate1 * tran_dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, row0 + (long long)(irate + 1) * n + i, cell);
            }
          }
        }
      }
#ifndef RXN_LANE_HOST
      if (H == 2) acc1 = __shfl_xor_sync(0xffffffffu, acc0, per);   // the odd-rate sum of lane (i, 1)
#endif
      if (i < n && half == 0) tsm[vr0 + (ikr * N + i) * CPB] = acc0 + acc1;
    }
  }
}

template <int N, int CPB>
LANE_DEV void lane_coop_out(const LaneTab &lt, const DevState &S, int slot, long long cell, int w, int W) {
  const int vsm = lt.o_vec + lt.s_sm * CPB + slot, vlng = lt.o_vec + lt.s_lng * CPB + slot;
#pragma unroll 4
  for (int k = w; k < lt.ncplx; k += W) {
    GSL(S, RXN_F_SEC_MOLAL, k, cell) = tsm[vsm + k * CPB];
    if (!lt.act_off) GSL(S, RXN_F_SEC_ACT_COEF, k, cell) = tsm[vlng + TI(lt, lt.i_ccls + k) * CPB];
  }
}

// x / d with r = 1/d precomputed (one Newton correction: the quotient the division unit returns, bar double rounding)
LANE_DEV double lane_div(double x, double d, double r) {
#ifndef RXN_LANE_HOST
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
#else
  (void)r;
  return x / d;
#endif
}

// One trip of a lane group through the Newton loop of RReact (reaction.F90:3411-3500): one iteration, or -
// when `closing` - the shortened last pass that redoes RTotal for the closing RTAuxVarCompute (:3507) after an
// abnormal exit changed pri_molal.  Returns 0 to continue, -1 after a closing pass, else the exit reason /
// flag; `recompute` is set when the closing RTAuxVarCompute needs such a pass.  All return values are
// uniform over the group.
template <int N, int CPB, int G>
LANE_DEV int lane_trip(const LaneTab &lt, Lane<N, G> &c, const DevState &S, double tran_dt, double inv_dt, int dt_mode, bool closing,
                       bool &recompute) {
  constexpr int LDJ2 = (N + 2) / 2;
  constexpr int R = Lane<N, G>::R;
  const int n = lt.n;
  const int bcol = 2 * (c.jb + (N >> 1) * CPB) + (N & 1), brow = 2 * LDJ2 * CPB;   // b_i = tsm[bcol + i*brow]
  const int bown = bcol + c.l * brow;                                              // own rows: + r*G*brow
  recompute = false;
  if (!closing) {
    c.iter = c.iter + 1;
    // :3407-3409 (once, before the loop) and :3413-3418 (every iteration): the call before the loop and
    // the call of iteration 1 see identical inputs, so one evaluation serves both
    if (!lt.act_off && (c.iter == 1 || lt.act_newton_iter)) lane_act_coefs<N, CPB, G>(lt, c);
  }
  // RTAuxVarCompute :3419 -> RTotal + RTotalSorb
  lane_speciate<N, CPB, G>(lt, c);
  lane_plan<N, CPB, G, false>(lt, c, 0.0);
  if (closing) {
    grp_sync<G>(c.gm);
#pragma unroll 1
    for (int i = c.l; i < n; i += G) tsm[c.vtot + i * CPB] = (tsm[c.vm + i * CPB] + tsm[c.vtot + i * CPB]) * c.den_kg_per_L;
    grp_sync<G>(c.gm);
    return -1;
  }
  {
    double2 z; z.x = 0.0; z.y = 0.0;
#pragma unroll 8
    for (int e = c.l; e < N * LDJ2; e += G) TSM2[c.jb + e * CPB] = z;
  }
  grp_sync<G>(c.gm);
  const double dp = c.den_kg_per_L * c.psvd;                   // dtotal * psvd_t  (:3429-3437; RTAccumulationDerivative :5189-5204)
  lane_plan<N, CPB, G, true>(lt, c, dp);
  grp_sync<G>(c.gm);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = c.l + r * G;
    if (i < N) JE(c, i, i) = fma(tsm[c.vm + i * CPB], dp, JE(c, i, i));   // REASSOC: (1 + D_ii/m_i) m_i
  }
  // sorption: equilibrium reactions (RTotalSorb :4182-4216, sorbed totals -> b) and the equilibrium part of the
  // multirate reactions (RMultiRateSorption reaction_surf_complex.F90:566-654, S_eq -> its own vector), one call site.
  // REASSOC: the multirate derivative block enters J before the mineral block.
#pragma unroll 1
  for (int task = 0; task < lt.neq + lt.nmr; ++task) {
    const bool eq = task < lt.neq;
    const int ikr = task - lt.neq;
    int tb = bcol, ts = brow;
    double fac = c.v_t;
    if (!eq) {
      tb = c.vseq + ikr * N * CPB; ts = CPB;
      fac = c.volume * lt.mrK1[ikr];
#pragma unroll 1
      for (int i = c.l; i < n; i += G) tsm[tb + i * ts] = 0.0;
    }
    lane_srf_rxn<N, CPB, G>(lt, c, S, TI(lt, (eq ? lt.i_eq_rxn : lt.i_mr_rxn - lt.neq) + task), fac, true, false, tb, ts);
  }
  const bool consistent = dt_mode == RXN_DT_CONSISTENT;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = c.l + r * G;
    if (i < N) {
      const double tot = (tsm[c.vm + i * CPB] + tsm[c.vtot + i * CPB]) * c.den_kg_per_L;   // :4095,4124,4148
      tsm[c.vtot + i * CPB] = tot;
      double res = c.psv * tot;
      res = res - c.fix[r];                                    // :3424-3426
      if (lt.neqsorb > 0) res = res + tsm[bown + r * G * brow] * c.volume;
      if (consistent) res = lane_div(res, tran_dt, inv_dt);
      tsm[bown + r * G * brow] = res;
    }
  }
  // RReaction :3440 (minerals, then multirate)
  if (lt.nkin > 0) lane_kinetic_mineral<N, CPB, G>(lt, c);
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr) {
#pragma unroll 1
    for (int i = c.l; i < n; i += G)
      tsm[bcol + i * brow] += c.volume * (lt.mrK1[ikr] * tsm[c.vseq + (ikr * N + i) * CPB] - tsm[c.vr0 + (ikr * N + i) * CPB]);
  }
  double mx = 0.0;
  bool bad = false;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (c.l + r * G < N) {
      const double v = tsm[bown + r * G * brow];
      mx = fmax(mx, fabs(v));
      if (!isfinite(v)) bad = true;
    }
  }
  if (grp_any<G>(bad, c.gm)) { recompute = true; return RXN_FLAG_NONFINITE; }
  mx = grp_max<G>(mx, c.gm);
  if (mx < lt.res_tol) return RXN_EXIT_RESIDUAL;               // :3443
  grp_sync<G>(c.gm);
  if (lane_rsolve<N, CPB, G>(lt, c)) { recompute = true; return RXN_FLAG_LU_ZERO_ROW; }
  double maxrel = 0.0, min_ratio = 1.0e20;
  if (!lt.use_log) {                                           // :3459-3471
#pragma unroll 1
    for (int i = c.l; i < n; i += G) {
      const double prev = tsm[c.vm + i * CPB], u = tsm[bcol + i * brow];
      if (prev <= u) {
        const double ratio = fabs(prev / u);
        if (ratio < min_ratio) min_ratio = ratio;
      }
    }
    min_ratio = grp_min<G>(min_ratio, c.gm);
  }
  // the new solution is staged in b: it is discarded when the relative change has converged (:3476)
#pragma unroll 2
  for (int i = c.l; i < n; i += G) {
    double u = tsm[bcol + i * brow];
    const double prev = tsm[c.vm + i * CPB];
    double nw;
    if (lt.use_log) {                                          // :3454-3458
      u = copysign(1.0, u) * fmin(fabs(u), lt.max_dlnC);
      nw = prev * exp(-u);
    } else {
      if (min_ratio < 1.0) u = u * min_ratio * 0.99;
      nw = prev - u;
    }
    const double rc = fabs((nw - prev) / prev);
    if (!isfinite(rc)) bad = true;
    maxrel = fmax(maxrel, rc);
    if (c.iter > 50) nw = 0.1 * (nw - prev) + prev;            // :3478-3496
    tsm[bcol + i * brow] = nw;
  }
  if (grp_any<G>(bad, c.gm)) { recompute = true; return RXN_FLAG_NONFINITE; }
  maxrel = grp_max<G>(maxrel, c.gm);
  if (maxrel < lt.rel_tol) return RXN_EXIT_REL_CHANGE;         // :3476 (update discarded)
#pragma unroll 4
  for (int i = c.l; i < n; i += G) tsm[c.vm + i * CPB] = tsm[bcol + i * brow];   // :3498
  grp_sync<G>(c.gm);
  if (c.iter >= lt.maxit) { recompute = true; return RXN_FLAG_CAPPED; }   // GPU-only guard (reference spins)
  return 0;
}

// closing RTAuxVarCompute (:3507) + write back (store_cell of the thread-per-cell path + reactive_transport.F90:1711).
// After a normal exit pri_molal and the activity coefficients are those of the last RTotal, so sec_molal and
// total are already final; only RTotalSorb sees a different input (the warm-start free-site concentration).
template <int N, int CPB, int G>
LANE_DEV void lane_finish(const LaneTab &lt, Lane<N, G> &c, const DevState &S, const DevTab &h, double *tran_xx, int32_t *iters,
                          int32_t *flags, int status) {
  constexpr int LDJ2 = (N + 2) / 2;
  const int n = lt.n;
  const long long cell = c.cell;
  const int bcol = 2 * (c.jb + (N >> 1) * CPB) + (N & 1), brow = 2 * LDJ2 * CPB;
  grp_sync<G>(c.gm);
  if (lt.neqsorb > 0) {
#pragma unroll 1
    for (int i = c.l; i < n; i += G) tsm[bcol + i * brow] = 0.0;
    if (lt.neq > 0) {                                          // RZeroSorb :4162-4178
#pragma unroll 1
      for (int k = c.l; k < lt.nsrf; k += G) GSL(S, RXN_F_EQSRFCPLX_CONC, k, cell) = 0.0;
    }
#pragma unroll 1
    for (int ieq = 0; ieq < lt.neq; ++ieq)
      lane_srf_rxn<N, CPB, G>(lt, c, S, TI(lt, lt.i_eq_rxn + ieq), 0.0, false, true, bcol, brow);
  }
#pragma unroll 2
  for (int i = c.l; i < n; i += G) {
    const double mm = tsm[c.vm + i * CPB];
    tran_xx[c.item * n + i] = mm;
    GSL(S, RXN_F_PRI_MOLAL, i, cell) = mm;
    GSL(S, RXN_F_TOTAL, i, cell) = tsm[c.vtot + i * CPB];
    if (lt.neqsorb > 0) GSL(S, RXN_F_TOTAL_SORB_EQ, i, cell) = tsm[bcol + i * brow];
  }
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr)
#pragma unroll 1
    for (int i = c.l; i < n; i += G)
      GSL(S, RXN_F_KINMR_TOTAL_SORB, (long long)ikr * (h.mr_ld + 1) * n + i, cell) = tsm[c.vseq + (ikr * N + i) * CPB];
  if (!lt.act_off) {
#pragma unroll 1
    for (int q = c.l; q < lt.ncls; q += G) tsm[c.vlng + q * CPB] = c_exp(tsm[c.vlng + q * CPB]);   // gamma per class
    grp_sync<G>(c.gm);
#pragma unroll 1
    for (int i = c.l; i < n; i += G) GSL(S, RXN_F_PRI_ACT_COEF, i, cell) = tsm[c.vlng + TI(lt, lt.i_pcls + i) * CPB];
  }
  if (!lt.coop_io) {
#pragma unroll 2
    for (int k = c.l; k < lt.ncplx; k += G) {
      GSL(S, RXN_F_SEC_MOLAL, k, cell) = tsm[c.vsm + k * CPB];
      if (!lt.act_off) GSL(S, RXN_F_SEC_ACT_COEF, k, cell) = tsm[c.vlng + TI(lt, lt.i_ccls + k) * CPB];
    }
  }
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) GSL(S, RXN_F_FREE_SITE_CONC, q, cell) = tsm[c.vfree + q * CPB];
#pragma unroll 1
  for (int q = c.l; q < lt.nkin; q += G) GSL(S, RXN_F_MNRL_RATE, q, cell) = tsm[c.vmnrl + (2 * lt.nkin + q) * CPB];
  if (c.l == 0) {
    GSL(S, RXN_F_LN_ACT_H2O, 0, cell) = c.ln_act_h2o;
    if (iters) iters[c.item] = c.iter;
    if (flags) flags[c.item] = status | c.flags;
  }
  grp_sync<G>(c.gm);
}

// ---------------------------------------------------------------------------------------------
// Global-implicit block of one cell: accumulation + reaction parts of RTResidualNonFlux
// (reactive_transport.F90:2545-2586, 2735-2758) and RTJacobianNonFlux (:3342-3389, 3445-3465), as
// cell_residual_jacobian of the thread-per-cell path: total / dtotal re-evaluated from pri_molal and the state's
// activity coefficients, res = (RTAccumulation + RAccumulationSorb)/dt + RReaction, jac = accumulation derivative block
// + reaction derivative block (column-major, with respect to m_j: the ln-m block divided by m_j on the way out).
template <int N, int CPB, int G>
LANE_DEV void lane_gi_cell(const LaneTab &lt, Lane<N, G> &c, const DevState &S, const double *blob_d, const int *blob_i, const DevTab &h,
                           long long item, long long cell, double dt, double *res_out, double *jac_out) {
  constexpr int LDJ2 = (N + 2) / 2;
  const int n = lt.n;
  const int bcol = 2 * (c.jb + (N >> 1) * CPB) + (N & 1), brow = 2 * LDJ2 * CPB;
  c.item = item; c.cell = cell; c.flags = 0; c.iter = 0;
  c.ln_act_h2o = GSL(S, RXN_F_LN_ACT_H2O, 0, cell);
  c.den_kg = GSL(S, RXN_F_DEN_KG, 0, cell);
  c.temp = GSL(S, RXN_F_TEMP, 0, cell);
  c.volume = GSL(S, RXN_F_VOLUME, 0, cell);
  c.porosity = GSL(S, RXN_F_POROSITY, 0, cell);
  c.soil_density = GSL(S, RXN_F_SOIL_PARTICLE_DENSITY, 0, cell);
  const double sat = GSL(S, RXN_F_SAT, 0, cell);
  c.psv = c.porosity * sat * 1000.0 * c.volume;
  c.psvd = c.porosity * sat * 1000.0 * c.volume / dt;
  c.v_t = c.volume / dt;
  c.den_kg_per_L = c.den_kg * 1.0 * 1.0e-3;
  grp_sync<G>(c.gm);
#pragma unroll 2
  for (int i = c.l; i < N; i += G) {
    tsm[c.vm + i * CPB] = i < n ? GSL(S, RXN_F_PRI_MOLAL, i, cell) : 1.0;
    if (i < n) tsm[c.vlng + i * CPB] = log(GSL(S, RXN_F_PRI_ACT_COEF, i, cell));      // ln_act = ln_conc + log(pri_act_coef) :4090
  }
#pragma unroll 8
  for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vlng + (n + k) * CPB] = GSL(S, RXN_F_SEC_ACT_COEF, k, cell);
  if (c.l == 0) {
    tsm[c.vlna + n * CPB] = 0.0;
    tsm[c.vlna + (n + 1) * CPB] = c.ln_act_h2o;
    tsm[c.vsm + lt.ncplx * CPB] = 0.0;
  }
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) tsm[c.vfree + q * CPB] = GSL(S, RXN_F_FREE_SITE_CONC, q, cell);
#pragma unroll 1
  for (int q = c.l; q < lt.nkin; q += G) {
    tsm[c.vmnrl + q * CPB] = GSL(S, RXN_F_MNRL_VOLFRAC, q, cell);
    tsm[c.vmnrl + (lt.nkin + q) * CPB] = GSL(S, RXN_F_MNRL_AREA, q, cell);
  }
  if (lt.percell_logK)
    lane_percell_logK<CPB, G>(c.l, lt.ncoef, lt.logK_mode, c.vlk, c.temp, GSL(S, RXN_F_PRES, 0, cell), blob_d, h.cplx, h.kin, h.srf);
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr) {                     // multirate_prepare (rxn_device.cuh)
    const int nrate = blob_i[h.o_mr_nrate + ikr];
#pragma unroll 1
    for (int i = c.l; i < n; i += G) {
      double acc0 = 0.0, acc1 = 0.0;                            // even / odd rates, as lane_coop_in_mr
#pragma unroll 4
      for (int irate = 0; irate < nrate; irate += 2) {
        const double rate = blob_d[h.o_mr_rate + ikr * h.mr_ld + irate];
        acc0 = acc0 + rate / (1.0 + rate * dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, ((long long)ikr * (h.mr_ld + 1) + irate + 1) * n + i, cell);
        if (irate + 1 < nrate) {
          const double rate1 = blob_d[h.o_mr_rate + ikr * h.mr_ld + irate + 1];
          acc1 = acc1 + rate1 / (1.0 + rate1 * dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, ((long long)ikr * (h.mr_ld + 1) + irate + 2) * n + i, cell);
        }
      }
      tsm[c.vr0 + (ikr * N + i) * CPB] = acc0 + acc1;
    }
  }
  grp_sync<G>(c.gm);
  // RTAuxVarCompute: RTotal + RTotalSorb
  lane_speciate<N, CPB, G>(lt, c);
  lane_plan<N, CPB, G, false>(lt, c, 0.0);
  {
    double2 z; z.x = 0.0; z.y = 0.0;
#pragma unroll 8
    for (int e = c.l; e < N * LDJ2; e += G) TSM2[c.jb + e * CPB] = z;
  }
  grp_sync<G>(c.gm);
  const double dp = c.den_kg_per_L * c.psvd;
  lane_plan<N, CPB, G, true>(lt, c, dp);
  grp_sync<G>(c.gm);
#pragma unroll 1
  for (int i = c.l; i < N; i += G) JE(c, i, i) = fma(tsm[c.vm + i * CPB], dp, JE(c, i, i));
  if (lt.neq > 0) {                                            // RZeroSorb :4162-4178
#pragma unroll 1
    for (int k = c.l; k < lt.nsrf; k += G) GSL(S, RXN_F_EQSRFCPLX_CONC, k, cell) = 0.0;
  }
#pragma unroll 1
  for (int task = 0; task < lt.neq + lt.nmr; ++task) {
    const bool eq = task < lt.neq;
    const int ikr = task - lt.neq;
    int tb = bcol, ts = brow;
    double fac = c.v_t;
    if (!eq) {
      tb = c.vseq + ikr * N * CPB; ts = CPB;
      fac = c.volume * lt.mrK1[ikr];
#pragma unroll 1
      for (int i = c.l; i < n; i += G) tsm[tb + i * ts] = 0.0;
    }
    lane_srf_rxn<N, CPB, G>(lt, c, S, TI(lt, (eq ? lt.i_eq_rxn : lt.i_mr_rxn - lt.neq) + task), fac, true, eq, tb, ts);
  }
#pragma unroll 1
  for (int i = c.l; i < n; i += G) {
    const double tot = (tsm[c.vm + i * CPB] + tsm[c.vtot + i * CPB]) * c.den_kg_per_L;
    GSL(S, RXN_F_TOTAL, i, cell) = tot;
    double res = c.psv * tot;                                  // RTAccumulation :5072-5148
    if (lt.neqsorb > 0) {
      const double tsorb = tsm[bcol + i * brow];
      GSL(S, RXN_F_TOTAL_SORB_EQ, i, cell) = tsorb;
      res = res + tsorb * c.volume;                            // RAccumulationSorb :4539-4568
    }
    tsm[bcol + i * brow] = res / dt;
  }
  if (lt.nkin > 0) lane_kinetic_mineral<N, CPB, G>(lt, c);
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr) {
#pragma unroll 1
    for (int i = c.l; i < n; i += G) {
      tsm[bcol + i * brow] += c.volume * (lt.mrK1[ikr] * tsm[c.vseq + (ikr * N + i) * CPB] - tsm[c.vr0 + (ikr * N + i) * CPB]);
      GSL(S, RXN_F_KINMR_TOTAL_SORB, (long long)ikr * (h.mr_ld + 1) * n + i, cell) = tsm[c.vseq + (ikr * N + i) * CPB];
    }
  }
  grp_sync<G>(c.gm);
  // outputs
  {
    bool bad = false;
#pragma unroll 1
    for (int i = c.l; i < n; i += G) if (!isfinite(tsm[bcol + i * brow])) bad = true;
    int fl = c.flags | (bad ? RXN_FLAG_NONFINITE : 0);
    if (fl != 0 && S.fail) {
#ifndef RXN_LANE_HOST
      atomicOr(S.fail, (unsigned int)fl);
#else
      __atomic_fetch_or(S.fail, (unsigned int)fl, __ATOMIC_RELAXED);
#endif
    }
  }
  if (res_out) {
#pragma unroll 1
    for (int i = c.l; i < n; i += G) res_out[item * n + i] = tsm[bcol + i * brow];
  }
  if (jac_out) {
#pragma unroll 1
    for (int j = c.l; j < n; j += G) tsm[c.vscr + j * CPB] = 1.0 / tsm[c.vm + j * CPB];
    grp_sync<G>(c.gm);
    double *jo = jac_out + item * (long long)(n * n);
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      const double invm = tsm[c.vscr + j * CPB];
#pragma unroll 2
      for (int i = c.l; i < n; i += G) jo[i + j * n] = JE(c, i, j) * invm;
    }
  }
#pragma unroll 4
  for (int k = c.l; k < lt.ncplx; k += G) GSL(S, RXN_F_SEC_MOLAL, k, cell) = tsm[c.vsm + k * CPB];
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) GSL(S, RXN_F_FREE_SITE_CONC, q, cell) = tsm[c.vfree + q * CPB];
#pragma unroll 1
  for (int q = c.l; q < lt.nkin; q += G) GSL(S, RXN_F_MNRL_RATE, q, cell) = tsm[c.vmnrl + (2 * lt.nkin + q) * CPB];
  grp_sync<G>(c.gm);
}

}  // namespace lane
}  // namespace rxn
