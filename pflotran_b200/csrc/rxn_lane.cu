// rxn_lane.cu — host side of the resident-lane RReact kernel: shape selection, plan upload, launch dispatch.
#include <cstdio>
#include <cstdlib>

#include "rxn_lane.cuh"

namespace rxn {

#define RXN_LANE_DECL(n, cpb, g)                                                                                                 \
  template <> int lane_launch_variant<n, cpb, g>(const LaneTab &, size_t, int, const DevTab &, const double *, const double *,    \
                                                 const DevState &, double *, const int32_t *, long long, double, int, int32_t *, \
                                                 int32_t *, unsigned long long *, long long, cudaStream_t);
RXN_LANE_SHAPES(RXN_LANE_DECL)
#undef RXN_LANE_DECL
#define RXN_LANE_DECL(n, cpb, g)                                                                                                    \
  template <> int lane_launch_gi_variant<n, cpb, g>(const LaneTab &, size_t, int, const DevTab &, const double *, const double *,   \
                                                    const DevState &, const int32_t *, long long, double, double *, double *, cudaStream_t);
RXN_LANE_SHAPES(RXN_LANE_DECL)
#undef RXN_LANE_DECL

#define RXN_TM_DECL(n, q, g)                                                                                                   \
  template <> int tm_launch_variant<n, q, g>(const LaneTab &, size_t, int, const DevTab &, const double *, const double *,     \
                                             const DevState &, double *, const int32_t *, long long, double, int, int32_t *,   \
                                             int32_t *, unsigned long long *, long long, cudaStream_t);
RXN_TM_SHAPES(RXN_TM_DECL)
#undef RXN_TM_DECL

#define RXN_TM_DECL(n, q, g)                                                                                                   \
  template <> int tm_launch_gi_variant<n, q, g>(const LaneTab &, size_t, int, const DevTab &, const double *, const double *,  \
                                                const DevState &, const int32_t *, long long, const GiArgs &, cudaStream_t);
RXN_TM_SHAPES(RXN_TM_DECL)
#undef RXN_TM_DECL

// tensor-memory kernel: same plan with J in TMEM; the first compiled shape (N, QUADS, G) whose vectors fit
static int tm_kernel_build(const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi, const cudaDeviceProp &prop,
                           int N, LaneKernel *k) {
  LanePlan &p = k->plan_tm;
  p.usable = false;
  p.err = "tensor-memory kernel disabled (RXN_TM=0)";
  if (const char *e = getenv("RXN_TM")) { if (atoi(e) == 0) return RXN_OK; }
  if (prop.major != 10) { p.err = "tensor memory needs sm_100"; return RXN_OK; }
  int force_g = 3, force_q = 0;                                  // measured on B200, 300A chemistry: G = 3 > 4 > 2 > 1
  if (const char *e = getenv("RXN_TM_G")) force_g = atoi(e);
  if (const char *e = getenv("RXN_TM_QUADS")) force_q = atoi(e);
  struct Shape { int N, Q, G; };
  static const Shape shapes[] = {
#define RXN_TM_ROW(n, q, g) {n, q, g},
      RXN_TM_SHAPES(RXN_TM_ROW)
#undef RXN_TM_ROW
  };
  p.err = "no tensor-memory shape for this matrix dimension";
  for (const Shape &s : shapes) {
    if (s.N != N || s.G != force_g) continue;
    if (force_q && s.Q != force_q) continue;
    int rc = lane_plan_build(h, bd, bi, s.N, 32 * s.Q, prop.sharedMemPerBlockOptin - 1024, &p, false, s.G);
    if (rc != RXN_OK) return rc;
    k->G_tm = s.G; k->quads_tm = s.Q;
    if (p.usable) break;
    if (p.err.find("does not fit") == std::string::npos) break;
  }
  if (!p.usable) return RXN_OK;
  if (cudaMalloc(&k->d_blob_tm, p.blob.size()) != cudaSuccess ||
      cudaMemcpy(k->d_blob_tm, p.blob.data(), p.blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    p.usable = false;
    p.err = std::string("plan upload failed: ") + cudaGetErrorString(cudaGetLastError());
    return RXN_ERR_CUDA;
  }
  return RXN_OK;
}

int lane_kernel_build(const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi, int device, LaneKernel *k) {
  LanePlan &p = k->plan;
  p.usable = false;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { p.err = "cudaGetDeviceProperties failed"; return RXN_OK; }
  k->sm_count = prop.multiProcessorCount;
  int N = lane_N_for(h.naq);
  if (const char *e = getenv("RXN_LANE_N")) { if (atoi(e) >= h.naq) N = atoi(e); }   // tests: a padded shape on purpose
  if (N == 0) { p.err = "naq exceeds the compiled shapes"; return RXN_OK; }
  int force_cpb = 0, force_g = 0;
  if (const char *e = getenv("RXN_LANE_CPB")) force_cpb = atoi(e);
  if (const char *e = getenv("RXN_LANE_G")) force_g = atoi(e);
  static const LaneShape shapes[] = {
#define RXN_LANE_ROW(n, cpb, g) {n, cpb, g},
      RXN_LANE_SHAPES(RXN_LANE_ROW)
#undef RXN_LANE_ROW
  };
  auto pick = [&](LanePlan &pl, int &Gout, bool gamma_state) {
    for (const LaneShape &s : shapes) {
      if (s.N != N) continue;
      if (force_cpb && s.CPB != force_cpb) continue;
      if (force_g && s.G != force_g) continue;
      int rc = lane_plan_build(h, bd, bi, s.N, s.CPB, prop.sharedMemPerBlockOptin, &pl, gamma_state);
      if (rc != RXN_OK) return rc;
      Gout = s.G;
      if (pl.usable) break;
      if (pl.err.find("does not fit") == std::string::npos) break;     // chemistry, not shape, is the obstacle
    }
    return (int)RXN_OK;
  };
  int rc0 = pick(p, k->G, false);
  if (rc0 != RXN_OK) return rc0;
  rc0 = tm_kernel_build(h, bd, bi, prop, N, k);
  if (rc0 != RXN_OK) return rc0;
  rc0 = pick(k->plan_gi, k->G_gi, true);
  if (rc0 != RXN_OK) return rc0;
  if (h.nmr > 0 && k->plan_gi.usable) {
    // measured (profiles/r01_r9_bench_gi.txt): the 750 multirate sorbed totals per cell are read faster by the
    // thread-per-cell kernel (one thread per cell at full occupancy) than by 2 lanes x 48 resident cells
    k->plan_gi.usable = false;
    k->plan_gi.err = "multirate sorption: residual/Jacobian blocks stay on the thread-per-cell kernel";
  }
  if (k->plan_gi.usable) {
    if (cudaMalloc(&k->d_blob_gi, k->plan_gi.blob.size()) != cudaSuccess ||
        cudaMemcpy(k->d_blob_gi, k->plan_gi.blob.data(), k->plan_gi.blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      k->plan_gi.usable = false;
      k->plan_gi.err = std::string("plan upload failed: ") + cudaGetErrorString(cudaGetLastError());
      return RXN_ERR_CUDA;
    }
  }
  k->mr_ld = h.mr_ld;
  k->mr_nrate.clear();
  for (int i = 0; i < h.nmr; ++i) k->mr_nrate.push_back(bi[h.o_mr_nrate + i]);
  k->mr_rate.assign(bd.begin() + h.o_mr_rate, bd.begin() + h.o_mr_rate + (size_t)h.nmr * h.mr_ld);
  k->mr_frac.assign(bd.begin() + h.o_mr_frac, bd.begin() + h.o_mr_frac + (size_t)h.nmr * h.mr_ld);
  if (!p.usable) return RXN_OK;
  if (cudaMalloc(&k->d_blob, p.blob.size()) != cudaSuccess ||
      cudaMemcpy(k->d_blob, p.blob.data(), p.blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    p.usable = false;
    p.err = std::string("plan upload failed: ") + cudaGetErrorString(cudaGetLastError());
    return RXN_ERR_CUDA;
  }
  if (getenv("RXN_LANE_VERBOSE"))
    fprintf(stderr, "[rxn lane] N=%d CPB=%d G=%d smem=%zu B blob=%zu B classes=%d spec %d terms/%d steps, planA %d/%d, planB %d/%d\n",
            p.lt.N, p.lt.CPB, k->G, p.smem_bytes, p.blob.size(), p.lt.ncls, p.terms_spec, p.steps_spec, p.terms_A, p.steps_A, p.terms_B,
            p.steps_B);
  return RXN_OK;
}

void lane_kernel_free(LaneKernel *k) {
  if (k->d_blob) cudaFree(k->d_blob);
  if (k->d_blob_gi) cudaFree(k->d_blob_gi);
  if (k->d_blob_tm) cudaFree(k->d_blob_tm);
  k->d_blob = k->d_blob_gi = k->d_blob_tm = nullptr;
  k->plan.usable = k->plan_gi.usable = k->plan_tm.usable = false;
}

static void lane_set_mrK1(const LaneKernel &k, LaneTab &lt, double dt) {
  // K1 = sum_r k_r/(1 + k_r dt) f_r (multirate_prepare, rxn_device.cuh): the same for every cell
  for (int ikr = 0; ikr < lt.nmr && ikr < 2; ++ikr) {
    double K1 = 0.0;
    for (int irate = 0; irate < k.mr_nrate[ikr]; ++irate) {
      const double rate = k.mr_rate[(size_t)ikr * k.mr_ld + irate], frac = k.mr_frac[(size_t)ikr * k.mr_ld + irate];
      const double kdt = rate * dt;
      const double one_plus_kdt = 1.0 + kdt;
      const double kk = rate / one_plus_kdt;
      K1 = K1 + kk * frac;
    }
    lt.mrK1[ikr] = K1;
  }
}

bool tm_gi_usable(const LaneKernel &k, int update_act) {
  if (!k.plan_tm.usable) return false;
  if (const char *e = getenv("RXN_GI_TM")) { if (atoi(e) == 0) return false; }
  return !(k.plan_tm.lt.act_off && update_act);
}

int tm_launch_gi(LaneKernel &k, const DevTab &h, const double *blob, const DevState &S, const int32_t *l2g, long long nlocal,
                 const GiArgs &a, cudaStream_t stream) {
  LaneTab lt = k.plan_tm.lt;
  lane_set_mrK1(k, lt, a.dt);
#define RXN_TM_CASE(n, q, g)                                                                                                   \
  if (lt.N == n && k.quads_tm == q && k.G_tm == g)                                                                              \
    return tm_launch_gi_variant<n, q, g>(lt, k.plan_tm.smem_bytes, k.sm_count, h, k.d_blob_tm, blob, S, l2g, nlocal, a, stream);
  RXN_TM_SHAPES(RXN_TM_CASE)
#undef RXN_TM_CASE
  return RXN_ERR_UNSUPPORTED;
}

int lane_launch_gi(LaneKernel &k, const DevTab &h, const double *blob, const DevState &S, const int32_t *l2g, long long nlocal, double dt,
                   double *res_out, double *jac_out, cudaStream_t stream) {
  LaneTab lt = k.plan_gi.lt;
  lane_set_mrK1(k, lt, dt);
#define RXN_LANE_CASE(n, cpb, g)                                                                                                   \
  if (lt.N == n && lt.CPB == cpb && k.G_gi == g)                                                                                   \
    return lane_launch_gi_variant<n, cpb, g>(lt, k.plan_gi.smem_bytes, k.sm_count, h, k.d_blob_gi, blob, S, l2g, nlocal, dt, res_out, \
                                             jac_out, stream);
  RXN_LANE_SHAPES(RXN_LANE_CASE)
#undef RXN_LANE_CASE
  return RXN_ERR_UNSUPPORTED;
}

int lane_launch_react(LaneKernel &k, const DevTab &h, const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                      long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter,
                      cudaStream_t stream, long long cell0) {
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXN_ERR_CUDA;
  if (k.plan_tm.usable) {
    LaneTab lt = k.plan_tm.lt;
    lane_set_mrK1(k, lt, dt);
#define RXN_TM_CASE(n, q, g)                                                                                                     \
    if (lt.N == n && k.quads_tm == q && k.G_tm == g)                                                                              \
      return tm_launch_variant<n, q, g>(lt, k.plan_tm.smem_bytes, k.sm_count, h, k.d_blob_tm, blob, S, tran_xx, l2g, nlocal, dt,  \
                                        dt_mode, iters, flags, counter, cell0, stream);
    RXN_TM_SHAPES(RXN_TM_CASE)
#undef RXN_TM_CASE
    return RXN_ERR_UNSUPPORTED;
  }
  LaneTab lt = k.plan.lt;
  lane_set_mrK1(k, lt, dt);
#define RXN_LANE_CASE(n, cpb, g)                                                                                                 \
  if (lt.N == n && lt.CPB == cpb && k.G == g)                                                                                    \
    return lane_launch_variant<n, cpb, g>(lt, k.plan.smem_bytes, k.sm_count, h, k.d_blob, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, \
                                       iters, flags, counter, cell0, stream);
  RXN_LANE_SHAPES(RXN_LANE_CASE)
#undef RXN_LANE_CASE
  return RXN_ERR_UNSUPPORTED;
}

}  // namespace rxn
