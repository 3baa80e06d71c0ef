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

constexpr int CODE_LAST = 1 << 16;

template <int G>
void tile_launch_variant(const TilePlan &p, const DevTab &h, const double *blob, const DevState &S, double *tran_xx,
                         const int32_t *l2g, long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags,
                         cudaStream_t stream);
#define RXN_TILE_SHAPES(X) X(1) X(2) X(4) X(8) X(16) X(32)
#define RXN_TILE_DECL(g)                                                                                              \
  extern template void tile_launch_variant<g>(const TilePlan &, const DevTab &, const double *, const DevState &, double *, \
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
    if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) { G = g; R = (n + G - 1) / G; }
  }
  tt.G = G; tt.NP = (n + 1) & ~1; if (tt.NP < G) tt.NP = G; tt.LDJ = tt.NP + 2;             // b in column NP, rows 16-byte aligned

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
  std::vector<double> pz2(n), cz2(std::max(h.ncplx, 1), 0.0);
  for (int i = 0; i < n; ++i) pz2[i] = bd[h.o_Z + i] * bd[h.o_Z + i];
  for (int k = 0; k < h.ncplx; ++k) cz2[k] = bd[h.o_cplxZ + k] * bd[h.o_cplxZ + k];
  std::vector<int32_t> pcls(n), ccls(h.ncplx);
  for (int i = 0; i < n; ++i) pcls[i] = class_of(bd[h.o_Z + i], bd[h.o_a0 + i]);
  for (int k = 0; k < h.ncplx; ++k) ccls[k] = class_of(bd[h.o_cplxZ + k], bd[h.o_cplxa0 + k]);
  tt.ncls = (int)z2.size();

  // -logK * LOG_TO_LN (fixed-temperature tables)
  std::vector<double> nlk;
  for (int k = 0; k < h.ncplx; ++k) nlk.push_back(-bd[h.cplx.o_logK + k] * RXN_LOG_TO_LN);
  for (int k = 0; k < h.nkin; ++k) nlk.push_back(-bd[h.kin.o_logK + k] * RXN_LOG_TO_LN);
  for (int k = 0; k < h.nsrf; ++k) nlk.push_back(-bd[h.srf.o_logK + k] * RXN_LOG_TO_LN);
  if (nlk.empty()) nlk.push_back(0.0);
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
  auto build = [&](std::vector<Entry> &E, int &T, int &o_rec, int &nterms) {
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
      load[best] += (int)E[e].terms.size() + 2;               // +2: cost of closing an entry
    }
    T = 0; nterms = 0;
    for (int l = 0; l < G; ++l) {
      int t = 0;
      for (int e : lane[l]) t += (int)E[e].terms.size();
      T = std::max(T, t);
      nterms += t;
    }
    struct Rec { double coef; int32_t code; int32_t key; };
    static_assert(sizeof(Rec) == 16, "plan record must be 16 bytes");
    std::vector<Rec> rec((size_t)T * G, Rec{0.0, h.ncplx, 0});
    for (int l = 0; l < G; ++l) {
      int t = 0;
      for (int e : lane[l])
        for (size_t q = 0; q < E[e].terms.size(); ++q, ++t) {
          const bool last = q + 1 == E[e].terms.size();
          rec[(size_t)t * G + l] = Rec{E[e].terms[q].second, E[e].terms[q].first | (last ? CODE_LAST : 0), last ? E[e].key : 0};
        }
    }
    std::vector<double> raw(rec.size() * 2);
    if (!rec.empty()) memcpy(raw.data(), rec.data(), rec.size() * 16);
    o_rec = D(raw);
  };
  build(EA, tt.TA, tt.o_A_rec, p->termsA);                   // records first: 16-byte aligned in shared memory
  build(EB, tt.TB, tt.o_B_rec, p->termsB);
  tt.o_cls_z2 = D(z2); tt.o_cls_a0 = D(a0);
  tt.o_pz2 = D(pz2); tt.o_cz2 = D(cz2);
  tt.o_nlk = D(nlk);
  tt.o_pri_cls = I(pcls); tt.o_cplx_cls = I(ccls);
  if (pi.size() & 1) pi.push_back(0);
  tt.ndbl = (int)pd.size(); tt.nint = (int)pi.size();

  // per-cell shared-memory layout (doubles)
  int maxsrf = 1;
  for (int r = 0; r < h.nrxn; ++r) maxsrf = std::max(maxsrf, bi[h.o_rxn_cptr + r + 1] - bi[h.o_rxn_cptr + r]);
  tt.maxsrf = maxsrf;
  tt.need_gam = h.maxpref > 0;
  int o = tt.NP * tt.LDJ;
  auto vec = [&](int len) { const int at = o; o += len; return at; };
  tt.c_m = vec(tt.NP); tt.c_invm = vec(tt.NP); tt.c_lna = vec(tt.NP); tt.c_tot = vec(tt.NP);
  tt.c_lgp = vec(tt.NP); tt.c_fix = vec(tt.NP);
  tt.c_tsorb = vec(h.neqsorb > 0 ? tt.NP : 0);
  tt.c_gam = vec(tt.need_gam ? tt.NP : 0);
  tt.c_sm = vec(h.ncplx + 1);
  tt.c_lng = vec(tt.ncls);
  tt.c_sc = vec(h.nrxn > 0 ? maxsrf : 0);
  tt.c_dsx = vec(tt.NP);                                      // sorption dSx/dm, reused as LU row scales
  tt.c_free = vec(h.nrxn);
  tt.c_mnrl = vec(2 * h.nkin);                                // mnrl_volfrac | mnrl_area
  tt.c_r0 = vec(h.nmr * tt.NP); tt.c_seq = vec(h.nmr * tt.NP);
  tt.c_lk = vec(tt.percell_logK ? (h.ncplx + h.nkin + h.nsrf) : 0);
  while ((o & 15) != 2) ++o;                                  // 16-byte aligned; groups of a warp start in different banks
  tt.pc_dbl = o;

  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return unusable("cudaGetDeviceProperties failed");
  const size_t smem_max = prop.sharedMemPerBlockOptin;
  const size_t fixed = main_blob_bytes + (size_t)tt.ndbl * 8 + (size_t)tt.nint * 4 + 32;
  const size_t avail = smem_max > fixed ? smem_max - fixed : 0;
  int cpb = (int)std::min<size_t>(avail / ((size_t)tt.pc_dbl * 8), (size_t)tile_max_threads(G) / G);
  if (cpb < 1) return unusable("tables do not fit in shared memory");
  const int maxgpw = 32 / G;                                  // lane groups a warp can hold
  int warps = (cpb + maxgpw - 1) / maxgpw;
  if (const char *e = getenv("RXN_TILE_WARPS")) warps = std::max(warps, std::min(tile_max_threads(G) / 32, atoi(e)));
  if (const char *e = getenv("RXN_TILE_CELLS")) cpb = std::max(1, std::min(cpb, atoi(e)));
  tt.gpw = (cpb + warps - 1) / warps;
  tt.cpb = cpb;
  tt.threads = 32 * warps;
  const int threads = tt.threads;
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
    fprintf(stderr, "[rxn tile] G=%d rows/lane=%d threads=%d (cells/warp %d) cells/CTA=%d smem=%zu B (per cell %d B) grid=%d ncls=%d planA T=%d (%d terms) planB T=%d (%d terms)\n",
            G, R, threads, tt.gpw, tt.cpb, p->smem_bytes, tt.pc_dbl * 8, p->grid, tt.ncls, tt.TA, p->termsA, tt.TB, p->termsB);
  return RXN_OK;
}

void tile_plan_free(TilePlan *p) {
  if (p->d_blob) cudaFree(p->d_blob);
  p->d_blob = nullptr;
  p->usable = false;
}

int tile_launch_react(const TilePlan &p, const DevTab &h, const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                      long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags, cudaStream_t stream) {
#define RXN_TILE_CASE(g) \
  if (p.tt.G == g) { tile_launch_variant<g>(p, h, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, stream); return RXN_OK; }
  RXN_TILE_SHAPES(RXN_TILE_CASE)
#undef RXN_TILE_CASE
  return RXN_ERR_UNSUPPORTED;
}

}  // namespace rxn
