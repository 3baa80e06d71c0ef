// rxn_lane.h — "resident lane" RReact kernel: plan structures and the host-side plan builder
// (pure C++, no CUDA: the CPU-only test harness compiles it as well).
//
// Design.  One THREAD solves one cell, as the north star asks, but nothing of a
// cell lives in local memory: the Newton system [J | b] (naq x (naq+1) doubles), sec_molal,
// ln activities and the few per-cell vectors sit in shared memory in a lane-fastest layout
// (element e of lane t at  e*CPB + t : every warp access is conflict free and needs no address
// arithmetic beyond an immediate), while the fixed accumulation and the per-cell scalars stay in
// registers.  Because all 32 lanes of a warp walk the chemistry tables in lock step, a table read is
// one shared-memory broadcast and the loop control is shared by 32 cells (round 1's cooperative kernel, 8 lanes per
// cell, paid it once per 4 cells).  Lanes are PERSISTENT: a lane whose cell has converged
// writes it back and takes the next cell from a global counter, so a warp never waits for its
// slowest cell (trip efficiency ~100 % instead of max-of-32 iterations).
//
// The sparse sums of RTotal are host-compiled into TERM STREAMS: groups of 4 independent
// accumulators that advance together (4 DFMA chains in flight per thread), each step one
// {coef[4], offset[4]} record.  A group either feeds 4 short sums (QUAD) or one long sum split 4
// ways (WIDE).  Three streams: speciation (lnQK_k), totals (plan A) and the symmetric
// d total/d ln m block (plan B).
//
// The Jacobian is assembled with respect to ln m_j (column j of the reference's matrix times m_j):
// every term of dtotal, of the sorption derivatives and of the mineral derivatives carries the
// factor 1/m_j, and the log formulation multiplies column j by m_j again (RSolve,
// reaction.F90:4866-4870), so neither the 1/m_j nor the m_j multiplication is executed.  The row
// norms of RSolve, which the reference takes on the 1/m_j-scaled matrix, are formed on the fly.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "rxn_tab.h"

namespace rxn {

enum { LANE_QUAD = 0, LANE_WIDE = 1 };

struct LaneStream {     // a range of term-stream groups
  int g0, ng;           // first group header (int index, 8 ints per group), number of groups
};

// POD handed to the kernel by value (__grid_constant__).  Offsets: *_d double index, *_i int index
// into the staged plan blob [doubles][ints]; s_* shared-memory vector slots (element e of slot s of
// lane t = tsm[o_vec + (s + e)*CPB + t]).
struct LaneTab {
  int N, CPB, LDJ2, cells;
  int n, ncplx, nkin, nsrf, nrxn, neq, nmr, neqsorb;
  int ncls, act_off, act_newton_iter, use_act_h2o, h2o_aq_id, use_log, percell_logK, maxit;
  int has_Temkin, has_scale, has_power, maxsrf;
  int logK_mode, ncoef;
  int coop_io;          // long per-complex arrays move warp-cooperatively (many complexes / multirate), else by the cell's own lanes
  int gamma_state;      // global-implicit plan: activity coefficients are the state's (one class per species; complexes divide by gamma)
  double debyeA, debyeB, debyeBdot, max_dlnC, rel_tol, res_tol;
  int blob_dbl, blob_int;
  int o_J2, o_vec, smem_dbl;
  int s_m, s_lna, s_lng, s_sm, s_tot, s_scr, s_sc, s_free, s_mnrl, s_r0, s_seq, s_lk;
  int jsink;            // J-region sink element (double index relative to 2*t) for padded plan-B closers
  int tm, tmG;          // tensor-memory kernel (rxn_tm_dev.cuh): J lives in TMEM (row i, column j at 32-bit column 2*(16 i + j)),
                        // plan-B closers are {column of (i,j), column of (j,i), offset of m_i for diagonal entries or -1, 0}
  int s_res, s_x;       // TMEM kernel: residual / right-hand side / solution vector (N), exchange slots of the G member warps (4 G)
  // term streams
  LaneStream spec, planA, planB;
  int d_coef, i_off;    // coef blocks (4 doubles per step, init block first) / offset blocks (4 ints per step, BYTE offsets
                        // into shared memory so that a gather address is one add: offset + 8*cell)
  // activity classes
  int d_cls_z2, d_cls_a0, d_pz2, d_cz2, i_pcls, i_ccls;
  int d_nlk;            // -logK*LOG_TO_LN for [complexes | minerals | surface complexes] (fixed logK)
  // minerals (CSR) and their rate parameters
  int i_kptr, i_kid, d_kst, d_kh2o, d_k_rate, d_k_Ea, d_k_aff, d_k_lim, d_k_Temkin, d_k_scale, d_k_power;
  // surface complexation
  int i_sptr, i_sid, d_sst, d_sh2o, d_site_st;
  int i_rxn_cptr, i_rxn_cid, i_rxn_surf_type, i_rxn_to_surf, i_rxn_flag, d_rxn_density, i_eq_rxn, i_mr_rxn;
  double mrK1[2];       // sum_r k_r/(1+k_r dt) f_r of the multirate reactions: set per launch (depends on dt)
};

// arguments of the global-implicit pass on the tensor-memory layout (tm_gi_cell, rxn_tm_dev.cuh)
enum { GI_AUX = 1, GI_RJ = 2 };
struct GiArgs {
  int mode, update_act;
  const double *xx;        // free-ion iterate, AoS [row][n] (row = cell for the auxvar update, item for the accumulation), or NULL
  int xx_by_item;
  double *accum_out;       // GI_AUX: [item][n] fixed accumulation, or NULL
  double *res_out, *jac_out;   // GI_RJ
  double dt;
};

struct LanePlan {
  bool usable = false;
  std::string err = "not built";
  LaneTab lt;
  std::vector<unsigned char> blob;   // [doubles][ints]
  size_t smem_bytes = 0;
  int terms_spec = 0, terms_A = 0, terms_B = 0, steps_spec = 0, steps_A = 0, steps_B = 0;
};

// supported (N, CPB) shapes, largest CPB first per N (instantiated in rxn_lane_variant.cu)
struct LaneShape { int N, CPB, G; };

inline int lane_N_for(int naq) {
  const int cand[] = {4, 8, 12, 15, 16, 24};
  if (naq == 15) return 15;
  for (int c : cand) if (c != 15 && naq <= c) return c;
  return 0;
}

// Build the plan for shape (N, CPB).  bd/bi: packed main tables (rxn_pack.h).  smem_max: bytes of
// shared memory one CTA may use.  Returns RXN_OK and sets p->usable (false + p->err if this chemistry
// or shape cannot use the kernel).
inline int lane_plan_build(const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi, int N, int CPB,
                           size_t smem_max, LanePlan *p, bool gamma_state = false, int tmG = 0) {
  p->usable = false;
  LaneTab &lt = p->lt;
  memset(&lt, 0, sizeof lt);
  auto unusable = [&](const char *why) { p->err = why; return RXN_OK; };
  const int n = h.naq;
  if (!gamma_state && h.act_alg == RXN_ACT_COEF_ALGORITHM_NEWTON && h.act_freq != RXN_ACT_COEF_FREQUENCY_OFF)
    return unusable("NEWTON activity-coefficient algorithm runs on the thread-per-cell kernel");
  if (h.nionx > 0 || h.nkd > 0) return unusable("ion exchange / KD isotherms run on the thread-per-cell kernel");
  if (h.maxpref > 0) return unusable("mineral prefactors run on the thread-per-cell kernel");
  if (h.ngen > 0 || h.ndecay > 0 || h.nkinrxn > 0 || h.nmic > 0 || h.nim > 0 || h.nimdecay > 0)
    return unusable("general / radioactive decay / kinetic surface complexation / microbial reactions and immobile species run on the thread-per-cell kernel");
  if (n > N) return unusable("naq exceeds the shape");
  if (tmG > 0 && N > 15) return unusable("tensor-memory kernel: a row [J_i | b_i] must fit 16 doubles (N <= 15)");
  lt.N = N; lt.CPB = CPB; lt.LDJ2 = (N + 2) / 2;
  lt.n = n; lt.ncplx = h.ncplx; lt.nkin = h.nkin; lt.nsrf = h.nsrf; lt.nrxn = h.nrxn; lt.neq = h.neq; lt.nmr = h.nmr;
  lt.neqsorb = h.neqsorb;
  lt.act_off = h.act_freq == RXN_ACT_COEF_FREQUENCY_OFF;
  lt.act_newton_iter = h.act_freq == RXN_ACT_COEF_FREQUENCY_NEWTON_ITER;
  lt.use_act_h2o = h.use_act_h2o; lt.h2o_aq_id = h.h2o_aq_id; lt.use_log = h.use_log;
  lt.percell_logK = h.logK_mode != RXN_LOGK_FIXED; lt.maxit = h.maxit;
  lt.has_Temkin = h.has_Temkin; lt.has_scale = h.has_scale; lt.has_power = h.has_power;
  lt.logK_mode = h.logK_mode; lt.ncoef = h.ncoef;
  lt.coop_io = h.ncplx >= 32 || h.nmr > 0;
  lt.gamma_state = gamma_state;
  lt.tm = tmG > 0; lt.tmG = tmG;
  lt.debyeA = h.debyeA; lt.debyeB = h.debyeB; lt.debyeBdot = h.debyeBdot;
  lt.max_dlnC = h.max_dlnC; lt.rel_tol = h.rel_tol; lt.res_tol = h.res_tol;

  std::vector<double> pd;
  std::vector<int32_t> pi;
  auto D = [&](const std::vector<double> &v) { int o = (int)pd.size(); pd.insert(pd.end(), v.begin(), v.end()); return o; };
  auto I = [&](const std::vector<int32_t> &v) { int o = (int)pi.size(); pi.insert(pi.end(), v.begin(), v.end()); return o; };
  auto Dsub = [&](int o, int cnt) { return D(std::vector<double>(bd.begin() + o, bd.begin() + o + std::max(cnt, 0))); };
  auto Isub = [&](int o, int cnt) { return I(std::vector<int32_t>(bi.begin() + o, bi.begin() + o + std::max(cnt, 0))); };

  // ---- activity classes: one Debye-Hueckel exponent per distinct (Z^2, a0); class 0 = neutral
  // (LAG threshold |Z| > 1e-10, reaction.F90:4013,4029).  With activity coefficients OFF every species
  // is its own class and ln gamma is taken from the state at load time.
  std::vector<double> z2(1, 0.0), a0(1, 0.0);
  std::vector<int32_t> pcls(n), ccls(std::max(h.ncplx, 1), 0);
  if (lt.act_off || gamma_state) {
    z2.assign(n + h.ncplx, 0.0); a0.assign(n + h.ncplx, 0.0);
    for (int i = 0; i < n; ++i) pcls[i] = i;
    for (int k = 0; k < h.ncplx; ++k) ccls[k] = n + k;
  } else {
    std::map<std::pair<double, double>, int> cls;
    auto class_of = [&](double Z, double a) {
      if (!(std::fabs(Z) > 1.0e-10)) return 0;
      auto key = std::make_pair(Z * Z, a);
      auto it = cls.find(key);
      if (it != cls.end()) return it->second;
      const int id = (int)z2.size();
      z2.push_back(Z * Z); a0.push_back(a);
      cls[key] = id;
      return id;
    };
    for (int i = 0; i < n; ++i) pcls[i] = class_of(bd[h.o_Z + i], bd[h.o_a0 + i]);
    for (int k = 0; k < h.ncplx; ++k) ccls[k] = class_of(bd[h.o_cplxZ + k], bd[h.o_cplxa0 + k]);
  }
  lt.ncls = (int)z2.size();
  int maxsrf = 1;
  for (int r = 0; r < h.nrxn; ++r) maxsrf = std::max(maxsrf, bi[h.o_rxn_cptr + r + 1] - bi[h.o_rxn_cptr + r]);
  lt.maxsrf = maxsrf;

  // ---- shared-memory slots (lane-fastest)
  int s = 0;
  auto slot = [&](int len) { const int at = s; s += len; return at; };
  lt.s_m = slot(N);
  lt.s_lna = slot(n + 2);                 // [n] = 0 (padding terms), [n+1] = ln a_H2O
  lt.s_lng = slot(lt.ncls);
  lt.s_sm = slot(h.ncplx + 2);            // [ncplx] = 0 (padding terms), [ncplx+1] = sink of padded closers
  lt.s_tot = slot(N);
  lt.s_scr = slot(N);                     // sorption dSx/dln m, then the LU row scales vv
  lt.s_sc = slot(h.nrxn > 0 ? maxsrf : 0);
  lt.s_free = slot(h.nrxn);
  lt.s_mnrl = slot(3 * h.nkin);           // volfrac | area | rate
  lt.s_r0 = slot(h.nmr * N); lt.s_seq = slot(h.nmr * N);
  lt.s_lk = slot(lt.percell_logK ? (h.ncplx + h.nkin + h.nsrf) : 0);
  lt.s_res = slot(tmG > 0 ? N : 0);
  lt.s_x = slot(tmG > 0 ? 4 * tmG : 0);
  const int nslots = s;
  const int jpairs = tmG > 0 ? 0 : N * lt.LDJ2 + 1;     // + one sink pair; the TMEM kernel keeps no J in shared memory

  // element offsets (double index; add the lane id t, or 2*t inside the J region)
  // J region starts at double2 index o_J2; vector region at double index o_vec (set below, after the blob size is known):
  // offsets are stored relative to those bases and rebased at the end.
  auto voff = [&](int sl, int e) { return (sl + e) * CPB; };                       // + o_vec
  auto joff = [&](int i, int j) {
    if (tmG > 0) return 2 * (16 * i + j);                                           // TMEM column (32-bit units), not rebased
    return 2 * ((i * lt.LDJ2 + (j >> 1)) * CPB) + (j & 1);                          // + 2*o_J2
  };
  const int jsink_rel = tmG > 0 ? 2 * (16 * 15) : 2 * (N * lt.LDJ2 * CPB);          // TMEM: row 15 is never a matrix row

  // ---- term streams
  struct Entry { int dest0, dest1; std::vector<std::pair<int, double>> terms; double init; int kk; int moff = -1; };
  struct Built { std::vector<int32_t> hdr; int ng = 0; int nterms = 0, nsteps = 0; };
  std::vector<double> coef;      // blocks: [init[4]] [steps][4]
  std::vector<int32_t> toff;     // blocks: [steps][4]
  std::vector<int32_t> ghdr;     // 8 ints per group: step0(coef block, in units of 4 doubles) , off0 (units of 4 ints), nsteps, mode, c0..c3 (closer index base / aux)
  std::vector<int32_t> closers;  // ints
  enum { REL_VEC = 0, REL_J = 1 };
  // zero-term padding: coef 0, offset -> sm[ncplx] (= 0.0), vector region
  const int zero_off = voff(lt.s_sm, h.ncplx);
  auto emit_group = [&](Built &B, int mode, const std::vector<const Entry *> &mem, int ndest) {
    // mem: QUAD up to 4 entries; WIDE exactly 1
    int nsteps = 0;
    if (mode == LANE_QUAD) for (auto *e : mem) nsteps = std::max(nsteps, (int)e->terms.size());
    else nsteps = ((int)mem[0]->terms.size() + 3) / 4;
    const int c0 = (int)coef.size() / 4, o0 = (int)toff.size() / 4;
    for (int a = 0; a < 4; ++a) {
      double init = 0.0;
      if (mode == LANE_QUAD) { if (a < (int)mem.size()) init = mem[a]->init; }
      else if (a == 0) init = mem[0]->init;
      coef.push_back(init);
    }
    for (int t = 0; t < nsteps; ++t)
      for (int a = 0; a < 4; ++a) {
        const Entry *e = nullptr; int q = -1;
        if (mode == LANE_QUAD) { if (a < (int)mem.size()) { e = mem[a]; q = t; } }
        else { e = mem[0]; q = t * 4 + a; }
        if (e && q < (int)e->terms.size()) { coef.push_back(e->terms[q].second); toff.push_back(e->terms[q].first); ++B.nterms; }
        else { coef.push_back(0.0); toff.push_back(zero_off); }
      }
    const int cbase = (int)closers.size();
    const int nmem = mode == LANE_QUAD ? 4 : 1;
    for (int a = 0; a < nmem; ++a) {
      const Entry *e = a < (int)mem.size() ? mem[a] : nullptr;
      if (ndest == 3) {                    // speciation: {sm offset, ln gamma offset, k}
        closers.push_back(e ? e->dest0 : voff(lt.s_sm, h.ncplx + 1));
        closers.push_back(e ? e->dest1 : voff(lt.s_lng, 0));
        closers.push_back(e ? e->kk : -1);
        closers.push_back(0);
      } else if (ndest == 1) {
        closers.push_back(e ? e->dest0 : voff(lt.s_sm, h.ncplx + 1));
      } else if (ndest == 4) {             // TMEM plan B: {column (i,j), column (j,i), m_i offset (diagonal) or -1, 0}
        closers.push_back(e ? e->dest0 : jsink_rel);
        closers.push_back(e ? e->dest1 : jsink_rel);
        closers.push_back(e ? e->moff : -1);
        closers.push_back(0);
      } else {
        closers.push_back(e ? e->dest0 : jsink_rel);
        closers.push_back(e ? e->dest1 : jsink_rel);
      }
    }
    while (closers.size() & 3) closers.push_back(0);
    ghdr.insert(ghdr.end(), {c0, o0, nsteps, mode, cbase, (int)mem.size(), 0, 0});
    ++B.ng; B.nsteps += nsteps;
  };
  auto build_stream = [&](std::vector<Entry> &E, int ndest, bool allow_wide, Built &B, LaneStream &S) {
    S.g0 = (int)ghdr.size();
    std::vector<const Entry *> wide, quad;
    for (auto &e : E) (allow_wide && e.terms.size() >= 8 ? wide : quad).push_back(&e);
    std::stable_sort(quad.begin(), quad.end(), [](const Entry *a, const Entry *b) { return a->terms.size() > b->terms.size(); });
    // groups in descending step count: the lanes of a cell take consecutive groups, so their loop trip counts stay close
    struct Grp { int mode, nsteps; std::vector<const Entry *> mem; };
    std::vector<Grp> groups;
    for (auto *e : wide) groups.push_back(Grp{LANE_WIDE, ((int)e->terms.size() + 3) / 4, {e}});
    for (size_t g = 0; g < quad.size(); g += 4) {
      Grp q{LANE_QUAD, 0, std::vector<const Entry *>(quad.begin() + g, quad.begin() + std::min(g + 4, quad.size()))};
      for (auto *e : q.mem) q.nsteps = std::max(q.nsteps, (int)e->terms.size());
      groups.push_back(q);
    }
    std::stable_sort(groups.begin(), groups.end(), [](const Grp &a, const Grp &b) { return a.nsteps > b.nsteps; });
    for (auto &g : groups) emit_group(B, g.mode, g.mem, ndest);
    S.ng = B.ng;
  };

  const int *ptr = bi.data() + h.cplx.o_ptr, *id = bi.data() + h.cplx.o_id;
  const double *st = bd.data() + h.cplx.o_st, *h2ost = bd.data() + h.cplx.o_h2ost;
  // speciation: lnQK_k = -logK_k*LOG_TO_LN [+ nu_w ln a_w] + sum nu ln a  (reaction.F90:4104-4118)
  std::vector<Entry> ES(h.ncplx), EA, EB;
  for (int k = 0; k < h.ncplx; ++k) {
    Entry &e = ES[k];
    e.dest0 = voff(lt.s_sm, k); e.dest1 = voff(lt.s_lng, ccls[k]); e.kk = k;
    e.init = lt.percell_logK ? 0.0 : -bd[h.cplx.o_logK + k] * 2.30258509299;
    if (h2ost[k] != 0.0) e.terms.push_back({voff(lt.s_lna, n + 1), h2ost[k]});
    for (int a = ptr[k]; a < ptr[k + 1]; ++a) e.terms.push_back({voff(lt.s_lna, id[a]), st[a]});
  }
  // plan A: total_i - m_i = sum_k nu_ik sm_k ; plan B: D_ij = sum_k nu_ik nu_jk sm_k, i <= j  (reaction.F90:4124-4146)
  {
    std::vector<Entry> rowsA(N);                              // padding rows of the shape: tot_i = 0 every trip
    std::map<int, int> bmap;
    for (int i = 0; i < N; ++i) { rowsA[i].dest0 = voff(lt.s_tot, i); rowsA[i].init = 0.0; rowsA[i].kk = 0; rowsA[i].dest1 = 0; }
    for (int k = 0; k < h.ncplx; ++k)
      for (int a = ptr[k]; a < ptr[k + 1]; ++a) {
        rowsA[id[a]].terms.push_back({voff(lt.s_sm, k), st[a]});
        for (int b = ptr[k]; b < ptr[k + 1]; ++b) {
          const int i = id[a], j = id[b];
          if (i > j) continue;
          auto it = bmap.find((i << 8) | j);
          if (it == bmap.end()) {
            it = bmap.insert({(i << 8) | j, (int)EB.size()}).first;
            Entry e; e.dest0 = joff(i, j); e.dest1 = joff(j, i); e.init = 0.0; e.kk = 0;
            EB.push_back(e);
          }
          // a species listed twice in one complex contributes twice, as in the reference loops
          EB[it->second].terms.push_back({voff(lt.s_sm, k), st[a] * st[b]});
        }
      }
    for (int i = 0; i < N; ++i) EA.push_back(rowsA[i]);       // rows without terms still write tot_i = 0
    if (tmG > 0) {
      // every diagonal entry exists (also for species in no complex and for the padding rows of the shape) and carries the
      // offset of m_i: its closer forms Jln_ii = fma(m_i, dp, D_ii dp) in one store, no read-modify-write on TMEM
      for (int i = 0; i < N; ++i) {
        auto it = bmap.find((i << 8) | i);
        if (it == bmap.end()) {
          it = bmap.insert({(i << 8) | i, (int)EB.size()}).first;
          Entry e; e.dest0 = joff(i, i); e.dest1 = joff(i, i); e.init = 0.0; e.kk = 0;
          EB.push_back(e);
        }
        EB[it->second].moff = voff(lt.s_m, i);
      }
    }
  }
  Built BS, BA, BB;
  build_stream(ES, 3, false, BS, lt.spec);
  build_stream(EA, 1, true, BA, lt.planA);
  build_stream(EB, tmG > 0 ? 4 : 2, true, BB, lt.planB);
  p->terms_spec = BS.nterms; p->terms_A = BA.nterms; p->terms_B = BB.nterms;
  p->steps_spec = BS.nsteps; p->steps_A = BA.nsteps; p->steps_B = BB.nsteps;

  // ---- assemble the blob.  Doubles: coef first (16-byte aligned records).  Ints: toff, closers, ghdr (multiples of 4).
  lt.d_coef = D(coef);
  lt.d_cls_z2 = D(z2); lt.d_cls_a0 = D(a0);
  {
    std::vector<double> pz2(n), cz2(std::max(h.ncplx, 1), 0.0), nlk;
    for (int i = 0; i < n; ++i) pz2[i] = bd[h.o_Z + i] * bd[h.o_Z + i];
    for (int k = 0; k < h.ncplx; ++k) cz2[k] = bd[h.o_cplxZ + k] * bd[h.o_cplxZ + k];
    for (int k = 0; k < h.ncplx; ++k) nlk.push_back(-bd[h.cplx.o_logK + k] * 2.30258509299);
    for (int k = 0; k < h.nkin; ++k) nlk.push_back(-bd[h.kin.o_logK + k] * 2.30258509299);
    for (int k = 0; k < h.nsrf; ++k) nlk.push_back(-bd[h.srf.o_logK + k] * 2.30258509299);
    if (nlk.empty()) nlk.push_back(0.0);
    lt.d_pz2 = D(pz2); lt.d_cz2 = D(cz2); lt.d_nlk = D(nlk);
  }
  const int nk = h.nkin;
  const int knnz = nk > 0 ? bi[h.kin.o_ptr + nk] : 0;
  lt.d_kst = Dsub(h.kin.o_st, knnz); lt.d_kh2o = Dsub(h.kin.o_h2ost, nk);
  lt.d_k_rate = Dsub(h.o_k_rate, nk); lt.d_k_Ea = Dsub(h.o_k_Ea, nk); lt.d_k_aff = Dsub(h.o_k_aff, nk);
  lt.d_k_lim = Dsub(h.o_k_lim, nk); lt.d_k_Temkin = Dsub(h.o_k_Temkin, nk); lt.d_k_scale = Dsub(h.o_k_scale, nk);
  lt.d_k_power = Dsub(h.o_k_power, nk);
  const int snnz = h.nsrf > 0 ? bi[h.srf.o_ptr + h.nsrf] : 0;
  lt.d_sst = Dsub(h.srf.o_st, snnz); lt.d_sh2o = Dsub(h.srf.o_h2ost, h.nsrf); lt.d_site_st = Dsub(h.o_srf_site_st, h.nsrf);
  lt.d_rxn_density = Dsub(h.o_rxn_density, h.nrxn);
  if (pd.size() & 1) pd.push_back(0.0);

  lt.i_off = I(toff);
  const int i_closers = I(closers);
  while (pi.size() & 3) pi.push_back(0);
  const int i_ghdr = I(ghdr);
  lt.i_pcls = I(pcls); lt.i_ccls = I(ccls);
  lt.i_kptr = Isub(h.kin.o_ptr, nk + 1); lt.i_kid = Isub(h.kin.o_id, knnz);
  lt.i_sptr = Isub(h.srf.o_ptr, h.nsrf + 1); lt.i_sid = Isub(h.srf.o_id, snnz);
  lt.i_rxn_cptr = Isub(h.o_rxn_cptr, h.nrxn + 1);
  lt.i_rxn_cid = Isub(h.o_rxn_cid, h.nrxn > 0 ? bi[h.o_rxn_cptr + h.nrxn] : 0);
  lt.i_rxn_surf_type = Isub(h.o_rxn_surf_type, h.nrxn); lt.i_rxn_to_surf = Isub(h.o_rxn_to_surf, h.nrxn);
  lt.i_rxn_flag = Isub(h.o_rxn_flag, h.nrxn);
  lt.i_eq_rxn = Isub(h.o_eq_rxn, h.neq); lt.i_mr_rxn = Isub(h.o_mr_rxn, h.nmr);
  while (pi.size() & 3) pi.push_back(0);
  lt.blob_dbl = (int)pd.size(); lt.blob_int = (int)pi.size();
  // group headers: absolute int offsets of their blocks
  lt.spec.g0 += i_ghdr; lt.planA.g0 += i_ghdr; lt.planB.g0 += i_ghdr;

  // ---- shared-memory layout: [blob][J region (double2)][vector region]
  const int blob_words = lt.blob_dbl + lt.blob_int / 2;            // doubles
  lt.o_J2 = (blob_words + 1) / 2;                                  // double2 index
  lt.o_vec = 2 * (lt.o_J2 + jpairs * CPB);
  const size_t fixed_bytes = (size_t)lt.o_J2 * 16;
  const size_t per_cell = (size_t)jpairs * 16 + (size_t)nslots * 8;
  // the regions are strided by CPB lanes whether or not every lane holds a cell
  const size_t need = fixed_bytes + per_cell * CPB;
  lt.cells = CPB;
  if (need > smem_max) {
    // fewer resident cells than lanes: idle lanes still own (unused) columns, so the stride CPB must fit
    return unusable("per-cell state of this shape does not fit in shared memory");
  }
  lt.smem_dbl = (int)(need / 8);
  p->smem_bytes = need;
  lt.jsink = tmG > 0 ? jsink_rel : 2 * lt.o_J2 + jsink_rel;
  // rebase offsets: vector-region offsets += o_vec ; J-region offsets += 2*o_J2
  {
    int32_t *ti = pi.data();
    const int32_t *gh = ti + i_ghdr;
    const int ngroups = (int)ghdr.size() / 8;
    for (int g = 0; g < ngroups; ++g) {
      const int o0 = gh[g * 8 + 1], nsteps = gh[g * 8 + 2];
      for (int q = 0; q < nsteps * 4; ++q) ti[lt.i_off + o0 * 4 + q] = (ti[lt.i_off + o0 * 4 + q] + lt.o_vec) * 8;   // byte offsets
    }
    auto fix_closers = [&](const LaneStream &S, int ndest) {
      for (int g = 0; g < S.ng; ++g) {
        int32_t *hd = ti + S.g0 + g * 8;
        const int nmem = hd[3] == LANE_QUAD ? 4 : 1;
        int32_t *c = ti + i_closers + hd[4];
        for (int a = 0; a < nmem; ++a) {
          if (ndest == 3) { c[a * 4] += lt.o_vec; c[a * 4 + 1] += lt.o_vec; }
          else if (ndest == 4) { if (c[a * 4 + 2] >= 0) c[a * 4 + 2] += lt.o_vec; }
          else if (ndest == 1) c[a] += lt.o_vec;
          else { c[a * 2] += 2 * lt.o_J2; c[a * 2 + 1] += 2 * lt.o_J2; }
        }
        hd[0] = lt.d_coef / 4 + hd[0];       // coef block index in units of 4 doubles (d_coef is 0: first array)
        hd[1] = lt.i_off / 4 + hd[1];
        hd[4] = i_closers + hd[4];
      }
    };
    fix_closers(lt.spec, 3); fix_closers(lt.planA, 1); fix_closers(lt.planB, tmG > 0 ? 4 : 2);
  }
  if (lt.d_coef != 0 || (lt.i_off & 3) || (i_closers & 3) || (i_ghdr & 3)) return unusable("internal: plan blob misaligned");
  p->blob.resize((size_t)lt.blob_dbl * 8 + (size_t)lt.blob_int * 4);
  memcpy(p->blob.data(), pd.data(), (size_t)lt.blob_dbl * 8);
  memcpy(p->blob.data() + (size_t)lt.blob_dbl * 8, pi.data(), (size_t)lt.blob_int * 4);
  p->usable = true;
  p->err.clear();
  return RXN_OK;
}

}  // namespace rxn
