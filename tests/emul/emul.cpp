// emul.cpp — HOST COMPILATION OF THE DEVICE ROUTINES, for tests only.
//
// TEST INFRASTRUCTURE.  This file compiles pflotran_b200/csrc/rxn_device.cuh (the per-cell
// device code of the CUDA kernels) with g++ by defining the CUDA qualifiers away, and runs the
// per-cell bodies in a plain loop over a host SoA image.  It exists so that the `-m "not gpu"`
// suite can check the device code's logic and the table packer (rxn_pack.h) against the oracle
// in the build container, which has no GPU.  It is NOT part of the product, is never loaded by
// pflotran_b200, and is not a fallback: the library (librxn_b200.so) fails with
// RXN_ERR_NO_DEVICE when no CUDA device is present.
#define __device__
#define __host__
#define __forceinline__ inline
#include <cmath>
#include <thread>
using std::isfinite;
#include "../../pflotran_b200/csrc/rxn_pack.h"
#include "../../pflotran_b200/csrc/rxn_device.cuh"
#undef RXN_LOG_TO_LN
#undef RXN_IDEAL_GAS_CONSTANT
#define RXN_LANE_HOST 1
#include "../../pflotran_b200/csrc/rxn_lane_dev.cuh"
#define RXN_TM_HOST 1
#include "../../pflotran_b200/csrc/rxn_tm_dev.cuh"
#include "../../pflotran_b200/csrc/rxn_flux.h"

using namespace rxn;

struct HostView { int64_t ncells, ld; double *f[RXN_F_COUNT]; };

struct Emu {
  PackResult R;
  std::vector<unsigned char> blob;
  Tab T;
  int nv;
};

static unsigned int g_cell_flags = 0;   // DevState::fail of the emulated launches (OR of the cell flags)
static DevState mk_state(const HostView *v, const uint8_t *active) {
  DevState S;
  for (int f = 0; f < RXN_F_COUNT; ++f) S.f[f] = v->f[f];
  S.ld = v->ld; S.ncells = v->ncells; S.active = active;
  S.fail = &g_cell_flags;
  return S;
}

#define EMU_DISPATCH(nv, CALL)          \
  switch (nv) {                         \
    case 4: { constexpr int N = 4; CALL; } break;   \
    case 8: { constexpr int N = 8; CALL; } break;   \
    case 16: { constexpr int N = 16; CALL; } break; \
    default: { constexpr int N = 24; CALL; } break; \
  }

// resident-lane RReact kernel (rxn_lane_dev.cuh): plan built for one resident cell (CPB = 1); the G lanes of the
// cell's group are G host threads that run the per-lane routines exactly as the persistent CUDA lanes do
// (load -> trips -> closing pass -> finish), meeting at a barrier where the device has __syncwarp / shuffles.
struct LaneJob {
  const LanePlan *P; const Emu *e; DevState *S; double *tran_xx; const int32_t *l2g; int64_t nlocal; double dt; int dt_mode;
  int32_t *iters, *flags;
};
template <int N, int G>
static void lane_thread(const LaneJob *J, LaneTab lt, int l) {
  using namespace rxn::lane;
  g_hl = l;
  const DevTab &h = J->e->R.h;
  const DevState &S = *J->S;
  const double inv_dt = 1.0 / J->dt;
  for (long long i = 0; i < J->nlocal; ++i) {
    const long long cell = J->l2g ? J->l2g[i] : i;
    if (S.active && !S.active[cell]) {
      if (l == 0) {
        if (J->iters) J->iters[i] = 0;
        if (J->flags) J->flags[i] = RXN_FLAG_INACTIVE;
      }
      continue;
    }
    Lane<N, G> c;
    lane_bind<N, 1, G>(lt, c, 0, l, 0u);
    lane_load<N, 1, G>(lt, c, S, J->e->T.d, J->e->T.i, h, i, cell, J->tran_xx, J->dt);
    if (l == 0 && lt.coop_io) {                                  // the warp-cooperative part, one "lane"
      const int slot3[3] = {0, 0, 0};
      const long long cell3[3] = {cell, cell, cell};
      lane_coop_in_sm<N, 1>(lt, S, slot3, cell3, 1, 0, 1);
      if (lt.nmr > 0) lane_coop_in_mr<N, 1>(lt, S, h, J->e->T.d, J->e->T.i, 0, cell, J->dt, 0, 1);
    }
    grp_sync<G>(0u);
    int pending = 0;
    for (;;) {
      bool recompute;
      const int st = lane_trip<N, 1, G>(lt, c, S, J->dt, inv_dt, J->dt_mode, pending != 0, recompute);
      bool fin = false;
      if (pending != 0) { lane_finish<N, 1, G>(lt, c, S, h, J->tran_xx, J->iters, J->flags, pending); fin = true; }
      else if (st != 0) {
        if (recompute) pending = st;
        else { lane_finish<N, 1, G>(lt, c, S, h, J->tran_xx, J->iters, J->flags, st); fin = true; }
      }
      if (fin) {
        if (l == 0 && lt.coop_io) lane_coop_out<N, 1>(lt, S, 0, cell, 0, 1);
        grp_sync<G>(0u);
        break;
      }
    }
  }
}
template <int N, int G>
static void lane_cells(const LaneJob &J) {
  using namespace rxn::lane;
  LaneTab lt = J.P->lt;
  const DevTab &h = J.e->R.h;
  for (int ikr = 0; ikr < lt.nmr && ikr < 2; ++ikr) {
    double K1 = 0.0;
    for (int irate = 0; irate < J.e->T.i[h.o_mr_nrate + ikr]; ++irate) {
      const double rate = J.e->T.d[h.o_mr_rate + ikr * h.mr_ld + irate], frac = J.e->T.d[h.o_mr_frac + ikr * h.mr_ld + irate];
      const double kdt = rate * J.dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
    }
    lt.mrK1[ikr] = K1;
  }
  std::vector<double> sm((size_t)lt.smem_dbl + 16, 0.0);
  memcpy(sm.data(), J.P->blob.data(), J.P->blob.size());
  tsm = sm.data();
  HostGroup hg;
  pthread_barrier_init(&hg.bar, nullptr, G);
  g_hg = &hg;
  std::vector<std::thread> th;
  for (int l = 1; l < G; ++l) th.emplace_back(lane_thread<N, G>, &J, lt, l);
  lane_thread<N, G>(&J, lt, 0);
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&hg.bar);
  g_hg = nullptr;
  tsm = nullptr;
}
struct LaneGiJob {
  const LanePlan *P; const Emu *e; DevState *S; const int32_t *l2g; int64_t nlocal; double dt; double *res, *jac;
};
template <int N, int G>
static void lane_gi_thread(const LaneGiJob *J, LaneTab lt, int l) {
  using namespace rxn::lane;
  g_hl = l;
  const DevTab &h = J->e->R.h;
  const DevState &S = *J->S;
  for (long long i = 0; i < J->nlocal; ++i) {
    const long long cell = J->l2g ? J->l2g[i] : i;
    if (S.active && !S.active[cell]) continue;
    Lane<N, G> c;
    lane_bind<N, 1, G>(lt, c, 0, l, 0u);
    lane_gi_cell<N, 1, G>(lt, c, S, J->e->T.d, J->e->T.i, h, i, cell, J->dt, J->res, J->jac);
  }
}
template <int N, int G>
static void lane_gi_cells(const LaneGiJob &J) {
  using namespace rxn::lane;
  LaneTab lt = J.P->lt;
  const DevTab &h = J.e->R.h;
  for (int ikr = 0; ikr < lt.nmr && ikr < 2; ++ikr) {
    double K1 = 0.0;
    for (int irate = 0; irate < J.e->T.i[h.o_mr_nrate + ikr]; ++irate) {
      const double rate = J.e->T.d[h.o_mr_rate + ikr * h.mr_ld + irate], frac = J.e->T.d[h.o_mr_frac + ikr * h.mr_ld + irate];
      const double kdt = rate * J.dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
    }
    lt.mrK1[ikr] = K1;
  }
  std::vector<double> sm((size_t)lt.smem_dbl + 16, 0.0);
  memcpy(sm.data(), J.P->blob.data(), J.P->blob.size());
  tsm = sm.data();
  HostGroup hg;
  pthread_barrier_init(&hg.bar, nullptr, G);
  g_hg = &hg;
  std::vector<std::thread> th;
  for (int l = 1; l < G; ++l) th.emplace_back(lane_gi_thread<N, G>, &J, lt, l);
  lane_gi_thread<N, G>(&J, lt, 0);
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&hg.bar);
  g_hg = nullptr;
  tsm = nullptr;
}
template <int N>
static void lane_gi_cells_g(const LaneGiJob &J, int G) {
  if (G == 1) lane_gi_cells<N, 1>(J);
  else if (G == 2) lane_gi_cells<N, 2>(J);
  else if (G == 8) { if constexpr (N == 24) lane_gi_cells<N, 8>(J); }      // the library's 24_20_8 / 24_16_8 shapes
  else lane_gi_cells<N, 4>(J);
}

template <int N>
static void lane_cells_g(const LaneJob &J, int G) {
  if (G == 1) lane_cells<N, 1>(J);
  else if (G == 2) lane_cells<N, 2>(J);
  else if (G == 8) { if constexpr (N == 24) lane_cells<N, 8>(J); }
  else lane_cells<N, 4>(J);
}

static const double *g_mr_kk = nullptr;     // the CTA's table of k_r/(1 + k_r dt) (tm_mr_kk_fill)
template <class JOB> static double tm_job_dt(const JOB &J) { return J.dt; }
#define TM_JOB_DT(J) tm_job_dt(J)
// tensor-memory RReact kernel (rxn_tm_dev.cuh): plan built for one resident cell (CPB = 1, J in the emulated TMEM lane); the G
// member warps of the cell are G host threads (one lane each) that run the rounds of k_react_tm: load -> trips (run, then
// closing after an abnormal exit) -> finish, meeting at a barrier where the device has the quad's named barrier.
template <int N, int G>
static void tm_thread(const LaneJob *J, LaneTab lt, int l) {
  using namespace rxn::tmk;
  const DevTab &h = J->e->R.h;
  const DevState &S = *J->S;
  const double inv_dt = 1.0 / J->dt;
  Ctx<N, G> c;
  tm_bind<N, 1, G>(lt, c, 0, l, 0, 0u);
  tm_init_column<N, 1, G>(lt, c);
  for (long long i = 0; i < J->nlocal; ++i) {
    const long long cell = J->l2g ? J->l2g[i] : i;
    if (S.active && !S.active[cell]) {
      if (l == 0) {
        if (J->iters) J->iters[i] = 0;
        if (J->flags) J->flags[i] = RXN_FLAG_INACTIVE;
      }
      continue;
    }
    tm_load<N, 1, G>(lt, c, S, J->e->T.d, J->e->T.i, h, i, cell, J->tran_xx, J->dt);
    if (lt.nmr > 0) tm_coop_in_mr<N, 1, G>(lt, S, h, J->e->T.d, J->e->T.i, l, 0, cell, J->dt, 0, 1, (cell & 1) ? g_mr_kk : nullptr);   // both forms of k_r/(1 + k_r dt)
    bool closing = false;
    int pending = 0;
    for (;;) {
      grp_sync<G>(c);
      int st;
      bool recompute;
      tm_trip<N, 1, G>(lt, c, S, J->dt, inv_dt, J->dt_mode, !closing, closing, st, recompute);
      bool fin = false;
      int status = 0;
      if (closing) { fin = true; status = pending; }
      else if (st != 0) {
        if (recompute) { pending = st; closing = true; }
        else { fin = true; status = st; }
      }
      if (fin) {
        tm_finish<N, 1, G>(lt, c, S, h, J->tran_xx, J->iters, J->flags, true, status);
        break;
      }
    }
  }
}
template <int N, int G>
static void tm_cells(const LaneJob &J) {
  using namespace rxn::tmk;
  LaneTab lt = J.P->lt;
  const DevTab &h = J.e->R.h;
  for (int ikr = 0; ikr < lt.nmr && ikr < 2; ++ikr) {
    double K1 = 0.0;
    for (int irate = 0; irate < J.e->T.i[h.o_mr_nrate + ikr]; ++irate) {
      const double rate = J.e->T.d[h.o_mr_rate + ikr * h.mr_ld + irate], frac = J.e->T.d[h.o_mr_frac + ikr * h.mr_ld + irate];
      const double kdt = rate * J.dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
    }
    lt.mrK1[ikr] = K1;
  }
  std::vector<double> sm((size_t)lt.smem_dbl + 16, 0.0), tmem(256, 0.0);
  memcpy(sm.data(), J.P->blob.data(), J.P->blob.size());
  static double kk_tab[TM_MR_KK];
  rxn::tmk::tm_mr_kk_fill(kk_tab, 0, 1, J.e->T.d, h, TM_JOB_DT(J));
  g_mr_kk = h.nmr * h.mr_ld <= TM_MR_KK ? kk_tab : nullptr;
  rxn::tmk::tsm = sm.data();
  rxn::tmk::tmh = tmem.data();
  rxn::tmk::HostGroup hg;
  pthread_barrier_init(&hg.bar, nullptr, G);
  rxn::tmk::g_hg = &hg;
  std::vector<std::thread> th;
  for (int l = 1; l < G; ++l) th.emplace_back(tm_thread<N, G>, &J, lt, l);
  tm_thread<N, G>(&J, lt, 0);
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&hg.bar);
  rxn::tmk::g_hg = nullptr;
  rxn::tmk::tsm = nullptr;
  rxn::tmk::tmh = nullptr;
}
template <int N>
static void tm_cells_g(const LaneJob &J, int G) {
  if (G == 1) tm_cells<N, 1>(J);
  else if (G == 2) tm_cells<N, 2>(J);
  else if (G == 3) tm_cells<N, 3>(J);
  else tm_cells<N, 4>(J);
}

// global-implicit pass on the tensor-memory layout (tm_gi_cell): G host threads per cell as above
struct TmGiJob {
  const LanePlan *P; Emu *e; DevState *S; const int32_t *l2g; int64_t nlocal; rxn::GiArgs a;
};
static double tm_job_dt(const TmGiJob &J) { return J.a.dt; }
template <int N, int G>
static void tm_gi_thread(const TmGiJob *J, LaneTab lt, int l) {
  using namespace rxn::tmk;
  const DevTab &h = J->e->R.h;
  const DevState &S = *J->S;
  Ctx<N, G> c;
  tm_bind<N, 1, G>(lt, c, 0, l, 0, 0u);
  tm_init_column<N, 1, G>(lt, c);
  for (long long i = 0; i < J->nlocal; ++i) {
    const long long cell = J->l2g ? J->l2g[i] : i;
    const bool on = !(S.active && !S.active[cell]);            // an inactive cell is walked with its stores off, as on the device
    tm_gi_cell<N, 1, G>(lt, c, S, h, J->e->T.d, J->e->T.i, J->a, i, cell, on, (cell & 1) ? g_mr_kk : nullptr);
  }
}
template <int N, int G>
static void tm_gi_cells(const TmGiJob &J) {
  LaneTab lt = J.P->lt;
  const DevTab &h = J.e->R.h;
  for (int ikr = 0; ikr < lt.nmr && ikr < 2; ++ikr) {
    double K1 = 0.0;
    for (int irate = 0; irate < J.e->T.i[h.o_mr_nrate + ikr]; ++irate) {
      const double rate = J.e->T.d[h.o_mr_rate + ikr * h.mr_ld + irate], frac = J.e->T.d[h.o_mr_frac + ikr * h.mr_ld + irate];
      const double kdt = rate * J.a.dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
    }
    lt.mrK1[ikr] = K1;
  }
  std::vector<double> sm((size_t)lt.smem_dbl + 16, 0.0), tmem(256, 0.0);
  memcpy(sm.data(), J.P->blob.data(), J.P->blob.size());
  static double kk_tab[TM_MR_KK];
  rxn::tmk::tm_mr_kk_fill(kk_tab, 0, 1, J.e->T.d, h, TM_JOB_DT(J));
  g_mr_kk = h.nmr * h.mr_ld <= TM_MR_KK ? kk_tab : nullptr;
  rxn::tmk::tsm = sm.data();
  rxn::tmk::tmh = tmem.data();
  rxn::tmk::HostGroup hg;
  pthread_barrier_init(&hg.bar, nullptr, G);
  rxn::tmk::g_hg = &hg;
  std::vector<std::thread> th;
  for (int l = 1; l < G; ++l) th.emplace_back(tm_gi_thread<N, G>, &J, lt, l);
  tm_gi_thread<N, G>(&J, lt, 0);
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&hg.bar);
  rxn::tmk::g_hg = nullptr;
  rxn::tmk::tsm = nullptr;
  rxn::tmk::tmh = nullptr;
}
template <int N>
static void tm_gi_cells_g(const TmGiJob &J, int G) {
  if (G == 1) tm_gi_cells<N, 1>(J);
  else if (G == 2) tm_gi_cells<N, 2>(J);
  else if (G == 3) tm_gi_cells<N, 3>(J);
  else tm_gi_cells<N, 4>(J);
}

extern "C" {

// mode: 1 = GI_AUX (auxvar update / fixed accumulation), 2 = GI_RJ (residual + Jacobian blocks); see rxn_lane.h GiArgs
int emu_gi_tm(void *hh, const HostView *v, const uint8_t *active, const int32_t *l2g, int64_t nlocal, int mode, int update_act,
              const double *xx, int xx_by_item, double *accum_out, double *res_out, double *jac_out, double dt, int G, char *err,
              int errlen) {
  Emu *e = (Emu *)hh;
  DevState S = mk_state(v, active);
  LanePlan P;
  const int N = e->R.h.naq <= 12 ? 12 : 15;
  if (G < 1 || G > 4) { if (err) snprintf(err, errlen, "G must be 1, 2, 3 or 4"); return RXN_ERR_INVALID; }
  int rc = lane_plan_build(e->R.h, e->R.P.d, e->R.P.i, N, 1, (size_t)1 << 30, &P, false, G);
  if (rc != RXN_OK || !P.usable) { if (err) snprintf(err, errlen, "%s", P.err.c_str()); return RXN_ERR_UNSUPPORTED; }
  if (P.lt.act_off && update_act) { if (err) snprintf(err, errlen, "activity update with activity coefficients off"); return RXN_ERR_UNSUPPORTED; }
  TmGiJob J{&P, e, &S, l2g, nlocal, rxn::GiArgs{mode, update_act, xx, xx_by_item, accum_out, res_out, jac_out, dt}};
  switch (N) {
    case 12: tm_gi_cells_g<12>(J, G); break;
    default: tm_gi_cells_g<15>(J, G); break;
  }
  return 0;
}

int emu_react_tm(void *hh, const HostView *v, double *tran_xx, const uint8_t *active, const int32_t *l2g, int64_t nlocal, double dt,
                 int dt_mode, int32_t *iters, int32_t *flags, int G, int forceN, char *err, int errlen) {
  Emu *e = (Emu *)hh;
  DevState S = mk_state(v, active);
  LanePlan P;
  const int N = forceN >= e->R.h.naq ? forceN : (e->R.h.naq <= 12 ? 12 : 15);
  if (G != 1 && G != 2 && G != 3 && G != 4) { if (err) snprintf(err, errlen, "G must be 1, 2, 3 or 4"); return RXN_ERR_INVALID; }
  int rc = lane_plan_build(e->R.h, e->R.P.d, e->R.P.i, N, 1, (size_t)1 << 30, &P, false, G);
  if (rc != RXN_OK || !P.usable) { if (err) snprintf(err, errlen, "%s", P.err.c_str()); return RXN_ERR_UNSUPPORTED; }
  LaneJob J{&P, e, &S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags};
  switch (N) {
    case 12: tm_cells_g<12>(J, G); break;
    case 15: tm_cells_g<15>(J, G); break;
    default: if (err) snprintf(err, errlen, "no tensor-memory shape N=%d in the harness", N); return RXN_ERR_UNSUPPORTED;
  }
  return 0;
}

void *emu_create(const RxnTablesDesc *d, char *err, int errlen) {
  Emu *e = new Emu();
  int rc = pack_tables(d, e->R);
  if (rc != RXN_OK) {
    if (err && errlen > 0) { strncpy(err, e->R.err.c_str(), errlen - 1); err[errlen - 1] = 0; }
    delete e;
    return nullptr;
  }
  e->blob = blob_bytes(e->R);
  e->T.d = reinterpret_cast<const double *>(e->blob.data());
  e->T.i = reinterpret_cast<const int *>(e->blob.data() + (size_t)e->R.h.ndbl * 8);
  e->T.h = &e->R.h;
  e->nv = variant_for(e->R.h.ncomp);
  return e;
}
void emu_destroy(void *h) { delete (Emu *)h; }
int emu_pack_status(const RxnTablesDesc *d, char *err, int errlen) {
  PackResult R;
  int rc = pack_tables(d, R);
  if (rc != RXN_OK && err && errlen > 0) { strncpy(err, R.err.c_str(), errlen - 1); err[errlen - 1] = 0; }
  return rc;
}
unsigned int emu_cell_flags(int reset) { const unsigned int v = g_cell_flags; if (reset) g_cell_flags = 0; return v; }
int emu_field_rows(void *h, int f) { return ((Emu *)h)->R.rows[f]; }
void emu_set_maxit(void *h, int maxit) { ((Emu *)h)->R.h.maxit = maxit; }

int emu_react_batch(void *h, const HostView *v, double *tran_xx, const uint8_t *active, const int32_t *l2g, int64_t nlocal,
                    double dt, int dt_mode, int32_t *iters, int32_t *flags) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long i = 0; i < nlocal; ++i) EMU_DISPATCH(e->nv, cell_react<N>(e->T, S, i, tran_xx, l2g, dt, dt_mode, iters, flags));
  return 0;
}
int emu_react_lane(void *hh, const HostView *v, double *tran_xx, const uint8_t *active, const int32_t *l2g, int64_t nlocal, double dt,
                   int dt_mode, int32_t *iters, int32_t *flags, int G, int forceN, char *err, int errlen, int32_t *stats) {
  Emu *e = (Emu *)hh;
  DevState S = mk_state(v, active);
  LanePlan P;
  const int N = forceN >= e->R.h.naq ? forceN : lane_N_for(e->R.h.naq);
  if (N == 0) { if (err) snprintf(err, errlen, "naq exceeds the compiled shapes"); return RXN_ERR_UNSUPPORTED; }
  if (G != 1 && G != 2 && G != 4 && !(G == 8 && N == 24)) { if (err) snprintf(err, errlen, "G must be 1, 2 or 4 (8 with N = 24)"); return RXN_ERR_INVALID; }
  int rc = lane_plan_build(e->R.h, e->R.P.d, e->R.P.i, N, 1, (size_t)1 << 30, &P);
  if (rc != RXN_OK || !P.usable) { if (err) snprintf(err, errlen, "%s", P.err.c_str()); return RXN_ERR_UNSUPPORTED; }
  if (stats) {
    stats[0] = P.lt.N; stats[1] = (int)P.blob.size(); stats[2] = (int)(P.smem_bytes - (size_t)P.lt.o_J2 * 16);
    stats[3] = P.terms_spec; stats[4] = P.steps_spec; stats[5] = P.terms_A; stats[6] = P.steps_A; stats[7] = P.terms_B; stats[8] = P.steps_B;
    stats[9] = P.lt.ncls;
  }
  LaneJob J{&P, e, &S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags};
  switch (N) {
    case 4: lane_cells_g<4>(J, G); break;
    case 8: lane_cells_g<8>(J, G); break;
    case 12: lane_cells_g<12>(J, G); break;
    case 15: lane_cells_g<15>(J, G); break;
    case 16: lane_cells_g<16>(J, G); break;
    default: lane_cells_g<24>(J, G); break;
  }
  return 0;
}

int emu_gi_lane(void *hh, const HostView *v, const uint8_t *active, const int32_t *l2g, int64_t nlocal, double dt, double *res_out,
                double *jac_out, int G, char *err, int errlen) {
  Emu *e = (Emu *)hh;
  DevState S = mk_state(v, active);
  LanePlan P;
  const int N = lane_N_for(e->R.h.naq);
  if (N == 0) { if (err) snprintf(err, errlen, "naq exceeds the compiled shapes"); return RXN_ERR_UNSUPPORTED; }
  if (G != 1 && G != 2 && G != 4 && !(G == 8 && N == 24)) { if (err) snprintf(err, errlen, "G must be 1, 2 or 4 (8 with N = 24)"); return RXN_ERR_INVALID; }
  int rc = lane_plan_build(e->R.h, e->R.P.d, e->R.P.i, N, 1, (size_t)1 << 30, &P, true);
  if (rc != RXN_OK || !P.usable) { if (err) snprintf(err, errlen, "%s", P.err.c_str()); return RXN_ERR_UNSUPPORTED; }
  LaneGiJob J{&P, e, &S, l2g, nlocal, dt, res_out, jac_out};
  switch (N) {
    case 4: lane_gi_cells_g<4>(J, G); break;
    case 8: lane_gi_cells_g<8>(J, G); break;
    case 12: lane_gi_cells_g<12>(J, G); break;
    case 15: lane_gi_cells_g<15>(J, G); break;
    case 16: lane_gi_cells_g<16>(J, G); break;
    default: lane_gi_cells_g<24>(J, G); break;
  }
  return 0;
}

int emu_update_auxvars_batch(void *h, const HostView *v, const double *xx_loc, const uint8_t *active, int update_act_coefs) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long c = 0; c < v->ncells; ++c) EMU_DISPATCH(e->nv, cell_update_auxvars<N>(e->T, S, c, xx_loc, update_act_coefs));
  return 0;
}
int emu_fixed_accum_batch(void *h, const HostView *v, const double *xx, const uint8_t *active, const int32_t *l2g, int64_t nlocal,
                          double *accum_out) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long i = 0; i < nlocal; ++i) EMU_DISPATCH(e->nv, cell_fixed_accum<N>(e->T, S, i, xx, l2g, accum_out));
  return 0;
}
int emu_residual_jacobian_batch(void *h, const HostView *v, const uint8_t *active, const int32_t *l2g, int64_t nlocal, double dt,
                                double *res_out, double *jac_out) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long i = 0; i < nlocal; ++i) EMU_DISPATCH(e->nv, cell_residual_jacobian<N>(e->T, S, i, l2g, dt, res_out, jac_out));
  return 0;
}
int emu_equilibrate_batch(void *h, const HostView *v, const uint8_t *active, const int32_t *ctype, const double *conc, int64_t conc_stride,
                          const int32_t *cid, const double *guess, int use_prev, int init_molal, int64_t nlocal, double *basis,
                          int32_t *iters, int32_t *status) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  const int n = e->R.h.naq;
  for (long long i = 0; i < nlocal; ++i) {
    if (active && !active[i]) { iters[i] = 0; status[i] = 0; continue; }
    int nit = 0, rc = 0;
    EMU_DISPATCH(e->nv, rc = cell_equilibrate<N>(e->T, S, i, ctype, conc + i * conc_stride, cid, guess, use_prev, init_molal,
                                                 basis ? basis + i * n : nullptr, &nit));
    iters[i] = nit; status[i] = rc;
  }
  return 0;
}
int emu_update_kinetic_state_batch(void *h, const HostView *v, const uint8_t *active, double dt) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long c = 0; c < v->ncells; ++c) EMU_DISPATCH(e->nv, cell_update_kinetic_state<N>(e->T, S, c, dt));
  return 0;
}

// Flux side (rxn_flux.h): the structure builder the library runs and the per-row arithmetic the kernels of rxn_flux.cuh
// implement (k_flux_residual's remainder path calls it; the unrolled paths restate it), run in plain loops with the kernels' index maps (coefficients SoA [component][connection], block CSR output).
// Returns the number of Jacobian blocks, < 0 on a structure error; row_ptr (nlocal+1) is always written, the other
// outputs only when non-NULL.
int64_t emu_flux(const HostView *v, const uint8_t *active, int n, int64_t nconn, const int32_t *id_up, const int32_t *id_dn,
                 const int32_t *g2l, int64_t nlocal, const double *area, const double *velocity, const double *disp,
                 const double *fraction_upwind, int use_upwinding, int32_t *row_ptr, int32_t *col, double *res, double *val) {
  FluxRows R;
  if (!flux_rows_build(v->ncells, nlocal, nconn, id_up, id_dn, g2l, active, &R)) return -1;
  memcpy(row_ptr, R.row_ptr.data(), R.row_ptr.size() * 4);
  if (col) memcpy(col, R.col.data(), R.col.size() * 4);
  std::vector<double> Tu((size_t)n * nconn), Td((size_t)n * nconn);
  for (int64_t c = 0; c < nconn; ++c)
    for (int i = 0; i < n; ++i)
      flux_coef(velocity[c], disp[c * n + i], area[c], use_upwinding ? 0.0 : fraction_upwind[c], use_upwinding, &Tu[(size_t)i * nconn + c],
                &Td[(size_t)i * nconn + c]);
  const double *tot = v->f[RXN_F_TOTAL], *D = v->f[RXN_F_DTOTAL];
  for (int64_t r = 0; r < nlocal; ++r) {
    const int s0 = R.row_ptr[r], s1 = R.row_ptr[r + 1];
    const int32_t own = R.l2g[r];
    if (res)
      for (int i = 0; i < n; ++i)
        res[r * n + i] = flux_row_residual(R.ent.data(), R.col.data(), s0, s1, own, tot + (int64_t)i * v->ld, &Tu[(size_t)i * nconn], &Td[(size_t)i * nconn]);
    if (val)
      for (int k = 0; k < s1 - s0; ++k)
        for (int e = 0; e < n * n; ++e) {
          const int i = e % n;
          double *dst = val + (int64_t)(s0 + k) * n * n;
          dst[e] = k == 0 ? flux_row_jac_diag(R.ent.data(), s0, s1, D[(int64_t)e * v->ld + own], &Tu[(size_t)i * nconn], &Td[(size_t)i * nconn])
                          : flux_row_jac_off(R.ent[s0 + k], D[(int64_t)e * v->ld + R.col[s0 + k]], &Tu[(size_t)i * nconn], &Td[(size_t)i * nconn]);
        }
  }
  return R.nnzb;
}

// The flux Jacobian by block COLUMNS (FluxCols: the traversal of k_flux_jacobian_cols) - same products and sums as the row
// walk, every block written once.  val [nnzb][n*n] must be sized by a previous emu_flux call.  Returns nnzb, < 0 on error.
int64_t emu_flux_cols(const HostView *v, const uint8_t *active, int n, int64_t nconn, const int32_t *id_up, const int32_t *id_dn,
                      const int32_t *g2l, int64_t nlocal, const double *area, const double *velocity, const double *disp,
                      const double *fraction_upwind, int use_upwinding, double *val) {
  FluxRows R;
  if (!flux_rows_build(v->ncells, nlocal, nconn, id_up, id_dn, g2l, active, &R)) return -1;
  FluxCols C;
  flux_cols_build(R, &C);
  std::vector<double> Tu((size_t)n * nconn), Td((size_t)n * nconn);
  for (int64_t c = 0; c < nconn; ++c)
    for (int i = 0; i < n; ++i)
      flux_coef(velocity[c], disp[c * n + i], area[c], use_upwinding ? 0.0 : fraction_upwind[c], use_upwinding, &Tu[(size_t)i * nconn + c],
                &Td[(size_t)i * nconn + c]);
  const double *D = v->f[RXN_F_DTOTAL];
  std::vector<char> written(R.nnzb, 0);
  for (int64_t c = 0; c < C.nghosted; ++c)
    for (int t = C.col_ptr[c]; t < C.col_ptr[c + 1]; ++t) {
      const int32_t en = C.tgt_ent[t], slot = C.tgt_slot[t];
      if (written[slot]++) return -2;
      double *dst = val + (int64_t)slot * n * n;
      for (int e = 0; e < n * n; ++e) {
        const int i = e % n;
        const double d = D[(int64_t)e * v->ld + c];
        if (en >= 0) dst[e] = flux_row_jac_off(en, d, &Tu[(size_t)i * nconn], &Td[(size_t)i * nconn]);
        else {
          const int row = C.col_row[c];
          dst[e] = flux_row_jac_diag(R.ent.data(), R.row_ptr[row], R.row_ptr[row + 1], d, &Tu[(size_t)i * nconn], &Td[(size_t)i * nconn]);
        }
      }
    }
  for (int64_t s = 0; s < R.nnzb; ++s) if (!written[s]) return -3;
  return R.nnzb;
}

// Coupler connections (boundary / source-sink) through the row view and per-row arithmetic of rxn_flux.h, i.e. what
// k_coupler_residual / k_coupler_jacobian implement.  c_ext / c_cell: nconn x n coefficient arrays for kind 0 (from
// flux_coef with fraction_upwind = 0.5, computed here from area / velocity / disp) or qsrc / type for kind 1.
// res [nlocal][n] and diag [nlocal][n*n] are updated in place; flux_out [nconn][n] optional.  Returns 0, < 0 on a structure error.
int emu_coupler(const HostView *v, const uint8_t *active, int kind, int n, int64_t nconn, const int32_t *id_dn, const int32_t *g2l,
                int64_t nlocal, const double *area, const double *velocity, const double *disp, int use_upwinding, const double *qsrc,
                const int32_t *ss_type, const double *ext_total, double *res, double *flux_out, double *diag) {
  CouplerRows R;
  if (!coupler_rows_build(v->ncells, nlocal, nconn, id_dn, g2l, active, &R)) return -1;
  std::vector<double> cx((size_t)n * nconn), cc((size_t)n * nconn), ext((size_t)n * nconn);
  for (int64_t c = 0; c < nconn; ++c)
    for (int i = 0; i < n; ++i) {
      if (kind == COUPLER_BOUNDARY) flux_coef(velocity[c], disp[c * n + i], area[c], 0.5, use_upwinding, &cx[(size_t)i * nconn + c], &cc[(size_t)i * nconn + c]);
      else ss_coef(qsrc[c], ss_type[c], &cc[(size_t)i * nconn + c], &cx[(size_t)i * nconn + c]);
      ext[(size_t)i * nconn + c] = ext_total[c * n + i];
    }
  const double sgn = kind == COUPLER_BOUNDARY ? -1.0 : 1.0;
  const double *tot = v->f[RXN_F_TOTAL], *D = v->f[RXN_F_DTOTAL];
  for (int64_t q = 0; q < R.nrows; ++q) {
    const int s0 = R.row_ptr[q], s1 = R.row_ptr[q + 1];
    const int64_t row = R.row[q], own = R.own[q];
    for (int i = 0; i < n; ++i) {
      const double t_own = tot[(int64_t)i * v->ld + own];
      if (res) res[row * n + i] = coupler_row_residual(res[row * n + i], R.conn.data(), s0, s1, sgn, &ext[(size_t)i * nconn], &cx[(size_t)i * nconn], &cc[(size_t)i * nconn], t_own);
      if (flux_out)
        for (int s = s0; s < s1; ++s) {
          const int32_t c = R.conn[s];
          flux_out[(int64_t)c * n + i] = fl_mul(sgn, coupler_res(cx[(size_t)i * nconn + c], ext[(size_t)i * nconn + c], cc[(size_t)i * nconn + c], t_own));
        }
    }
    if (diag)
      for (int e = 0; e < n * n; ++e)
        diag[row * n * n + e] = coupler_row_jac(diag[row * n * n + e], R.conn.data(), s0, s1, sgn, &cc[(size_t)(e % n) * nconn], D[(int64_t)e * v->ld + own]);
  }
  return 0;
}

}  // extern "C"

