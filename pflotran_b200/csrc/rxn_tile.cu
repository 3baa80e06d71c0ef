// rxn_tile.cu — host side of the cooperative RReact kernel: accumulation-plan builder, shared-memory
// layout, launch dispatch.  Device code: rxn_tile_dev.cuh (instantiated per shape in rxn_tile_variant.cu).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

#include "rxn_device.cuh"
#include "rxn_tile.cuh"

namespace rxn {

constexpr int TILE_MAX_THREADS = 768;
constexpr int CODE_LAST = 1 << 16;

template <int G, int R>
void tile_launch_variant(const TilePlan &p, const DevTab &h, const double *blob, const DevState &S, double *tran_xx,
                         const int32_t *l2g, long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags,
                         cudaStream_t stream);
#define RXN_TILE_SHAPES(X) X(4, 1) X(4, 2) X(4, 4) X(8, 1) X(8, 2) X(8, 3) X(16, 1) X(16, 2) X(32, 1)
#define RXN_TILE_DECL(g, r)                                                                                              \
  extern template void tile_launch_variant<g, r>(const TilePlan &, const DevTab &, const double *, const DevState &, double *, \
                                                 const int32_t *, long long, double, int, int32_t *, int32_t *, cudaStream_t);
RXN_TILE_SHAPES(RXN_TILE_DECL)
#undef RXN_TILE_DECL

// =============================================================================================
// plan builder
int tile_plan_build(const RxnTablesDesc *d, const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi,
                    size_t main_blob_bytes, int device, TilePlan *p) {
  p->usable = false;
  TileTab &tt = p->tt;
  memset(&tt, 0, sizeof tt);
  const int n = h.naq;
  auto unusable = [&](const char *why) { p->err = why; return RXN_OK; };
  if (h.act_alg == RXN_ACT_COEF_ALGORITHM_NEWTON && h.act_freq != RXN_ACT_COEF_FREQUENCY_OFF)
    return unusable("NEWTON activity-coefficient algorithm runs on the thread-per-cell kernel");
  if (h.nionx > 0 || h.nkd > 0) return unusable("ion exchange / KD isotherms run on the thread-per-cell kernel");
  if (n > 64) return unusable("naq > 64");
  if (h.ncplx >= 0xffff) return unusable("too many complexes");
  // lane-group shape: G lanes x R rows per lane
  int G = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 8 : 16, R = (n + G - 1) / G;
  if (const char *e = getenv("RXN_TILE_G")) {
    const int g = atoi(e);
    if (g == 4 || g == 8 || g == 16 || g == 32) { G = g; R = (n + G - 1) / G; }
  }
  if (R > 4) return unusable("more than 4 rows per lane");
  tt.G = G; tt.R = R; tt.NP = G * R; tt.LDJ = tt.NP + 1;
  if ((tt.LDJ & 1) == 0) tt.LDJ += 1;

  std::vector<double> pd;
  std::vector<int32_t> pi;
  auto D = [&](const std::vector<double> &v) { int o = (int)pd.size(); pd.insert(pd.end(), v.begin(), v.end()); return o; };
  auto I = [&](const std::vector<int32_t> &v) { int o = (int)pi.size(); pi.insert(pi.end(), v.begin(), v.end()); return o; };

  // activity classes (Z^2, a0); class 0 = neutral (LAG threshold |Z| > 1e-10, reaction.F90:4013,4029)
  std::vector<double> z2(1, 0.0), a0(1, 0.0);
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
  std::vector<int32_t> pcls(n), ccls(h.ncplx);
  for (int i = 0; i < n; ++i) pcls[i] = class_of(bd[h.o_Z + i], bd[h.o_a0 + i]);
  for (int k = 0; k < h.ncplx; ++k) ccls[k] = class_of(bd[h.o_cplxZ + k], bd[h.o_cplxa0 + k]);
  tt.ncls = (int)z2.size();
  tt.o_cls_z2 = D(z2); tt.o_cls_a0 = D(a0);
  tt.o_pri_cls = I(pcls); tt.o_cplx_cls = I(ccls);

  // -logK * LOG_TO_LN (fixed-temperature tables)
  std::vector<double> nlk;
  for (int k = 0; k < h.ncplx; ++k) nlk.push_back(-bd[h.cplx.o_logK + k] * RXN_LOG_TO_LN);
  for (int k = 0; k < h.nkin; ++k) nlk.push_back(-bd[h.kin.o_logK + k] * RXN_LOG_TO_LN);
  for (int k = 0; k < h.nsrf; ++k) nlk.push_back(-bd[h.srf.o_logK + k] * RXN_LOG_TO_LN);
  if (nlk.empty()) nlk.push_back(0.0);
  tt.o_nlk = D(nlk);
  tt.percell_logK = h.logK_mode != RXN_LOGK_FIXED;

  // plans
  struct Entry { int key; std::vector<std::pair<int, double>> terms; };
  const int *ptr = bi.data() + h.cplx.o_ptr, *id = bi.data() + h.cplx.o_id;
  const double *st = bd.data() + h.cplx.o_st;
  std::vector<Entry> EA(n), EB;
  std::map<int, int> bmap;
  for (int i = 0; i < n; ++i) EA[i].key = i;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) { bmap[(i << 8) | j] = (int)EB.size(); EB.push_back(Entry{(i << 8) | j, {}}); }
  for (int k = 0; k < h.ncplx; ++k)
    for (int a = ptr[k]; a < ptr[k + 1]; ++a) {
      EA[id[a]].terms.push_back({k, st[a]});
      for (int b = ptr[k]; b < ptr[k + 1]; ++b) {
        const int i = id[a], j = id[b];
        if (i > j) continue;
        auto &terms = EB[bmap[(i << 8) | j]].terms;
        // a species listed twice in one complex contributes twice, as in the reference loops
        terms.push_back({k, st[a] * st[b]});
      }
    }
  auto build = [&](std::vector<Entry> &E, int &T, int &o_coef, int &o_code, int &o_ent, int &nterms) {
    for (auto &e : E) if (e.terms.empty()) e.terms.push_back({h.ncplx, 0.0});
    std::vector<int> order(E.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return E[a].terms.size() > E[b].terms.size(); });
    std::vector<std::vector<int>> lane(G);
    std::vector<int> load(G, 0);
    for (int e : order) {
      int best = 0;
      for (int l = 1; l < G; ++l) if (load[l] < load[best]) best = l;
      lane[best].push_back(e);
      load[best] += (int)E[e].terms.size() + 1;               // +1: cost of closing an entry
    }
    T = 0; nterms = 0;
    size_t maxent = 0;
    for (int l = 0; l < G; ++l) {
      int t = 0;
      for (int e : lane[l]) t += (int)E[e].terms.size();
      T = std::max(T, t);
      nterms += t;
      maxent = std::max(maxent, lane[l].size());
    }
    std::vector<double> coef((size_t)T * G, 0.0);
    std::vector<int32_t> code((size_t)T * G, h.ncplx), ent(std::max<size_t>(maxent, 1) * G, 0);
    for (int l = 0; l < G; ++l) {
      int t = 0, ne = 0;
      for (int e : lane[l]) {
        for (size_t q = 0; q < E[e].terms.size(); ++q, ++t) {
          coef[(size_t)t * G + l] = E[e].terms[q].second;
          code[(size_t)t * G + l] = E[e].terms[q].first | (q + 1 == E[e].terms.size() ? CODE_LAST : 0);
        }
        ent[(size_t)ne * G + l] = E[e].key;
        ++ne;
      }
    }
    o_coef = D(coef); o_code = I(code); o_ent = I(ent);
  };
  build(EA, tt.TA, tt.o_A_coef, tt.o_A_code, tt.o_A_ent, p->termsA);
  build(EB, tt.TB, tt.o_B_coef, tt.o_B_code, tt.o_B_ent, p->termsB);
  if (pi.size() & 1) pi.push_back(0);
  tt.ndbl = (int)pd.size(); tt.nint = (int)pi.size();

  // per-cell shared-memory layout (doubles)
  int maxsrf = 1;
  for (int r = 0; r < h.nrxn; ++r) maxsrf = std::max(maxsrf, bi[h.o_rxn_cptr + r + 1] - bi[h.o_rxn_cptr + r]);
  tt.maxsrf = maxsrf;
  tt.need_gam = h.maxpref > 0;
  int o = tt.NP * tt.LDJ;
  tt.c_m = o; o += tt.NP;
  tt.c_invm = o; o += tt.NP;
  tt.c_lna = o; o += tt.NP;
  tt.c_tot = o; o += tt.NP;
  tt.c_gam = o; o += tt.need_gam ? tt.NP : 0;
  tt.c_sm = o; o += h.ncplx + 1;
  tt.c_lng = o; o += tt.ncls;
  tt.c_sc = o; o += h.nrxn > 0 ? maxsrf : 0;
  tt.c_dsx = o; o += h.nrxn > 0 ? tt.NP : 0;
  tt.c_free = o; o += h.nrxn;
  tt.c_lk = o; o += tt.percell_logK ? (h.ncplx + h.nkin + h.nsrf) : 0;
  if ((o & 1) == 0) o += 1;                                   // odd stride: groups of a warp hit different banks
  tt.pc_dbl = o;

  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return unusable("cudaGetDeviceProperties failed");
  const size_t smem_max = prop.sharedMemPerBlockOptin;
  const size_t fixed = main_blob_bytes + (size_t)tt.ndbl * 8 + (size_t)tt.nint * 4;
  int threads = TILE_MAX_THREADS;
  if (const char *e = getenv("RXN_TILE_THREADS")) threads = std::max(32, std::min(TILE_MAX_THREADS, atoi(e) / 32 * 32));
  while (threads > 32 && fixed + (size_t)(threads / G) * tt.pc_dbl * 8 > smem_max) threads -= 32;
  if (fixed + (size_t)(threads / G) * tt.pc_dbl * 8 > smem_max) return unusable("tables do not fit in shared memory");
  tt.threads = threads;
  tt.cpb = threads / G;
  p->smem_bytes = fixed + (size_t)tt.cpb * tt.pc_dbl * 8;
  p->grid = prop.multiProcessorCount;                         // x resident CTAs per SM (occupancy query at launch)

  p->blob_bytes = (size_t)tt.ndbl * 8 + (size_t)tt.nint * 4;
  std::vector<unsigned char> blob(p->blob_bytes);
  memcpy(blob.data(), pd.data(), (size_t)tt.ndbl * 8);
  memcpy(blob.data() + (size_t)tt.ndbl * 8, pi.data(), (size_t)tt.nint * 4);
  if (cudaMalloc(&p->d_blob, p->blob_bytes) != cudaSuccess ||
      cudaMemcpy(p->d_blob, blob.data(), p->blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    p->err = std::string("plan upload failed: ") + cudaGetErrorString(cudaGetLastError());
    return RXN_ERR_CUDA;
  }
  p->usable = true;
  p->err.clear();
  if (getenv("RXN_TILE_VERBOSE"))
    fprintf(stderr, "[rxn tile] G=%d R=%d threads=%d cells/CTA=%d smem=%zu B (per cell %d B) grid=%d ncls=%d planA T=%d (%d terms) planB T=%d (%d terms)\n",
            G, R, threads, tt.cpb, p->smem_bytes, tt.pc_dbl * 8, p->grid, tt.ncls, tt.TA, p->termsA, tt.TB, p->termsB);
  return RXN_OK;
}

void tile_plan_free(TilePlan *p) {
  if (p->d_blob) cudaFree(p->d_blob);
  p->d_blob = nullptr;
  p->usable = false;
}

int tile_launch_react(const TilePlan &p, const DevTab &h, const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                      long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags, cudaStream_t stream) {
#define RXN_TILE_CASE(g, r) \
  if (p.tt.G == g && p.tt.R == r) { tile_launch_variant<g, r>(p, h, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, stream); return RXN_OK; }
  RXN_TILE_SHAPES(RXN_TILE_CASE)
#undef RXN_TILE_CASE
  return RXN_ERR_UNSUPPORTED;
}

}  // namespace rxn
