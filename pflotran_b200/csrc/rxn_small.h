// rxn_small.h — "register" RReact kernel for SMALL chemistries (naq <= 4, <= 8 aqueous complexes, <= 2 kinetic minerals, no
// sorption): plan structure and host-side builder (pure C++: the CPU-only test harness compiles it as well).
//
// Why.  For the calcite-class chemistries of BASELINE configs 2 and 4 (4 primaries / 5 complexes / 1 mineral) the resident-lane
// kernel spends 3 750 thread instructions per cell and Newton iteration on machinery built for 15 x 15 systems (term-stream
// records, shared-memory vectors, generic LU with run-time pivot addressing, persistent-lane bookkeeping), about 5x what the
// arithmetic needs (profiles/r02_o_*: issue slots 49 % busy, FP64 pipe 22 %).  With N <= 4 EVERYTHING of a cell fits in
// registers when every loop is unrolled at compile time: the tables become DENSE (stoichiometry nu[k][j] with zeros, so no
// species index ever addresses a register array), they travel as a __grid_constant__ kernel parameter (constant bank: table
// operands are immediates of the FMAs, loop control runs on the uniform datapath), the Newton matrix is 4 x 5 registers and
// ludcmp's pivoting is a chain of selects.  One thread = one cell, as the north star words it.
//
// Arithmetic: as the resident-lane kernel (ln-m Jacobian, REASSOC notes of rxn_lane_dev.cuh), sums over species in ascending
// species order (a dense row adds exact zeros for the species a complex does not contain).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "rxn_tab.h"

namespace rxn {

enum { SMALL_N = 4, SMALL_MAXC = 8, SMALL_MAXK = 2, SMALL_NCOEF = 17 };

// POD handed to the kernel by value (__grid_constant__): ~3.5 KB
struct SmallTab {
  int n, ncplx, nkin;
  int act_off, act_newton_iter, use_act_h2o, h2o_aq_id, use_log, logK_mode, ncoef, maxit;
  int has_Temkin, has_scale, has_power;
  int cplx_fit[SMALL_MAXC], kin_fit[SMALL_MAXK];       // 1: logK from the per-cell T(,P) fit, 0: fixed
  double debyeA, debyeB, debyeBdot, max_dlnC, rel_tol, res_tol;
  double pz2[SMALL_N], pa0[SMALL_N];                    // primaries: Z^2, a0 (LAG threshold |Z| > 1e-10 applied at build: z2 = 0 -> gamma = 1)
  int pcharged[SMALL_N], ccharged[SMALL_MAXC];
  double cz2[SMALL_MAXC], ca0[SMALL_MAXC];
  double nu[SMALL_MAXC][SMALL_N];                       // dense stoichiometry of the complexes
  double ch2o[SMALL_MAXC], cnlk[SMALL_MAXC];            // H2O stoichiometry, -logK*LOG_TO_LN (fixed logK)
  double ccoef[SMALL_MAXC][SMALL_NCOEF];
  double nuk[SMALL_MAXK][SMALL_N];                      // minerals
  double kh2o[SMALL_MAXK], knlk[SMALL_MAXK];
  double kcoef[SMALL_MAXK][SMALL_NCOEF];
  double k_rate[SMALL_MAXK], k_Ea[SMALL_MAXK], k_aff[SMALL_MAXK], k_lim[SMALL_MAXK], k_Temkin[SMALL_MAXK], k_scale[SMALL_MAXK],
      k_power[SMALL_MAXK];
};

struct SmallPlan {
  bool usable = false;
  std::string err = "not built";
  SmallTab st;
};

// bd / bi: the packed main tables (rxn_pack.h)
inline int small_plan_build(const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi, SmallPlan *p) {
  p->usable = false;
  SmallTab &t = p->st;
  memset(&t, 0, sizeof t);
  auto no = [&](const char *why) { p->err = why; return RXN_OK; };
  if (h.naq > SMALL_N) return no("more than 4 primary species");
  if (h.ncplx > SMALL_MAXC) return no("more than 8 aqueous complexes");
  if (h.nkin > SMALL_MAXK) return no("more than 2 kinetic minerals");
  if (h.nrxn > 0 || h.nionx > 0 || h.nkd > 0 || h.neqsorb > 0) return no("sorption runs on the other kernels");
  if (h.maxpref > 0) return no("mineral prefactors run on the thread-per-cell kernel");
  if (h.ngen > 0 || h.ndecay > 0 || h.nkinrxn > 0) return no("general / decay / kinetic surface complexation reactions run on the thread-per-cell kernel");
  if (h.act_alg == RXN_ACT_COEF_ALGORITHM_NEWTON && h.act_freq != RXN_ACT_COEF_FREQUENCY_OFF)
    return no("NEWTON activity-coefficient algorithm runs on the thread-per-cell kernel");
  if (h.logK_mode != RXN_LOGK_FIXED && h.ncoef > SMALL_NCOEF) return no("logK fit with more than 17 coefficients");
  t.n = h.naq; t.ncplx = h.ncplx; t.nkin = h.nkin;
  t.act_off = h.act_freq == RXN_ACT_COEF_FREQUENCY_OFF;
  t.act_newton_iter = h.act_freq == RXN_ACT_COEF_FREQUENCY_NEWTON_ITER;
  t.use_act_h2o = h.use_act_h2o; t.h2o_aq_id = h.h2o_aq_id; t.use_log = h.use_log;
  t.logK_mode = h.logK_mode; t.ncoef = h.ncoef; t.maxit = h.maxit;
  t.has_Temkin = h.has_Temkin; t.has_scale = h.has_scale; t.has_power = h.has_power;
  t.debyeA = h.debyeA; t.debyeB = h.debyeB; t.debyeBdot = h.debyeBdot;
  t.max_dlnC = h.max_dlnC; t.rel_tol = h.rel_tol; t.res_tol = h.res_tol;
  for (int i = 0; i < h.naq; ++i) {
    const double Z = bd[h.o_Z + i];
    t.pz2[i] = Z * Z; t.pa0[i] = bd[h.o_a0 + i]; t.pcharged[i] = std::fabs(Z) > 1.0e-10;
  }
  auto fill = [&](const DSpec &sp, int r, double *nu_row, double *h2o, double *nlk, double *coef, int *fit, bool is_srf) {
    for (int a = bi[sp.o_ptr + r]; a < bi[sp.o_ptr + r + 1]; ++a) nu_row[bi[sp.o_id + a]] += bd[sp.o_st + a];   // a species listed twice adds up
    *h2o = bd[sp.o_h2ost + r];
    *nlk = -bd[sp.o_logK + r] * 2.30258509299;
    *fit = h.logK_mode != RXN_LOGK_FIXED && sp.o_coef >= 0 && !(is_srf && h.logK_mode == RXN_LOGK_HPT);
    if (*fit) for (int q = 0; q < h.ncoef; ++q) coef[q] = bd[sp.o_coef + r * h.ncoef + q];
  };
  for (int k = 0; k < h.ncplx; ++k) {
    // a species listed twice in one complex is summed term by term by the reference: keep that out of the dense form
    for (int a = bi[h.cplx.o_ptr + k]; a < bi[h.cplx.o_ptr + k + 1]; ++a)
      for (int b = a + 1; b < bi[h.cplx.o_ptr + k + 1]; ++b)
        if (bi[h.cplx.o_id + a] == bi[h.cplx.o_id + b]) return no("a complex lists a species twice");
    fill(h.cplx, k, t.nu[k], &t.ch2o[k], &t.cnlk[k], t.ccoef[k], &t.cplx_fit[k], false);
    const double Z = bd[h.o_cplxZ + k];
    t.cz2[k] = Z * Z; t.ca0[k] = bd[h.o_cplxa0 + k]; t.ccharged[k] = std::fabs(Z) > 1.0e-10;
  }
  for (int q = 0; q < h.nkin; ++q) {
    for (int a = bi[h.kin.o_ptr + q]; a < bi[h.kin.o_ptr + q + 1]; ++a)
      for (int b = a + 1; b < bi[h.kin.o_ptr + q + 1]; ++b)
        if (bi[h.kin.o_id + a] == bi[h.kin.o_id + b]) return no("a mineral lists a species twice");
    fill(h.kin, q, t.nuk[q], &t.kh2o[q], &t.knlk[q], t.kcoef[q], &t.kin_fit[q], false);
    t.k_rate[q] = bd[h.o_k_rate + q]; t.k_Ea[q] = bd[h.o_k_Ea + q]; t.k_aff[q] = bd[h.o_k_aff + q]; t.k_lim[q] = bd[h.o_k_lim + q];
    t.k_Temkin[q] = h.has_Temkin ? bd[h.o_k_Temkin + q] : 1.0;
    t.k_scale[q] = h.has_scale ? bd[h.o_k_scale + q] : 1.0;
    t.k_power[q] = h.has_power ? bd[h.o_k_power + q] : 1.0;
  }
  p->usable = true;
  p->err.clear();
  return RXN_OK;
}

}  // namespace rxn
