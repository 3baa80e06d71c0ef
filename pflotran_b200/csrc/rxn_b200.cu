// rxn_b200.cu — C ABI (include/rxn_b200.h) of the B200-native batched reaction path.
//
// Host side: validates and packs the reference's reaction tables into one device blob, owns
// the SoA FP64 cell state in HBM, and launches the batched kernels (rxn_kernels.cuh).
// There is NO CPU fallback: without a usable CUDA device every entry point fails with
// RXN_ERR_NO_DEVICE / RXN_ERR_CUDA.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>   // work order of the resident-lane kernel (react_order)
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3 (dlopens the tool's injection library; a no-op without a profiler)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rxn_b200.h"
#include "rxn_kernels.cuh"
#include "rxn_pack.h"
#include "rxn_lane.cuh"
#include "rxn_flux.cuh"

using namespace rxn;

namespace {

thread_local std::string g_err;
long long g_launches = 0;

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// NVTX ranges named after the reference's PetscLogEvents (logging.F90:327-381): a timeline of the GPU path reads like
// the reference's -log_view
struct Nvtx {
  explicit Nvtx(const char *name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
};

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RXN_ERR_NO_DEVICE : RXN_ERR_CUDA, \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

}  // namespace

struct RxnTables {
  DevTab h;
  mutable LaneKernel lane; // resident-lane (thread per cell, state in shared memory) kernel, rxn_lane.cuh
  double *d_blob = nullptr;
  size_t blob_bytes = 0;
  int device = 0;
  int rows[RXN_F_COUNT];
  int nvariant = 0;        // template instantiation (max naq)
};

struct RxnState {
  const RxnTables *t = nullptr;
  DevState S;
  uint8_t *d_active = nullptr;
  long long ncells = 0, ld = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  float last_ms = 0.f;
  // grow-only scratch for the host-buffer entry points
  void *scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[4] = {0, 0, 0, 0};
  int react_kernel = 0;    // 0 auto, 1 thread-per-cell, 3 resident lane / tensor memory (2: the cooperative kernel of round 1, removed)
  unsigned long long *d_counter = nullptr;   // work counter of the resident-lane kernel
  // host-buffer RReact: chunks of the batch move over PCIe while the previous / next chunk is being solved
  // (16 chunks; consecutive chunks are solved on two alternating streams with their own work counters, so the persistent CTAs of
  // chunk c+1 move onto the SMs that chunk c's CTAs have drained instead of waiting for its slowest cell)
  enum { NCHUNK = 16 };
  cudaStream_t h2d = nullptr, d2h = nullptr, stream2 = nullptr;
  unsigned long long *d_counters = nullptr;   // one work counter per chunk
  cudaEvent_t ev_in[NCHUNK] = {}, ev_k[NCHUNK] = {};
  int flux_generic = 0;    // RXN_FLUX_GENERIC=1: flux Jacobian through the run-time-n kernel (tests)
  int flux_rows = 0;       // RXN_FLUX_ROWS=1: N = 15 flux Jacobian by block rows (k_flux_jacobian_t) instead of block columns
  int gi_kernel = 0;       // global-implicit loops (RXN_GI_KERNEL): 0 auto (tensor-memory layout, else resident lanes, else thread per cell), 1 thread per cell, 2 resident lanes
  unsigned int *d_fail = nullptr;   // OR of the cell flags of the running global-implicit launch (DevState::fail)
  // work order of the resident-lane RReact kernel for tail-bound chemistries (react_order): items sorted by the Newton iteration
  // count of the previous call, slowest first
  int32_t *d_prev_it = nullptr, *d_keys = nullptr, *d_iota = nullptr, *d_order = nullptr;
  void *d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  long long order_cap = 0, prev_n = 0;     // allocated items; items of the remembered iteration counts (0: none)
};

// Row view of a connection list + the flux coefficients of the current flow field (rxn_flux.h)
struct RxnConnSet {
  RxnState *s = nullptr;
  FluxRows R;
  int n = 0, device = 0;      // device kept here: the set may outlive its state
  int32_t *d_row_ptr = nullptr, *d_col = nullptr, *d_ent = nullptr, *d_l2g = nullptr;
  int32_t *d_col_ptr = nullptr, *d_tgt_slot = nullptr, *d_tgt_ent = nullptr, *d_col_row = nullptr;   // column view (FluxCols)
  double *d_T = nullptr;        // [T_up | T_dn], each SoA [component][connection]
  double *d_Ta = nullptr;       // the same coefficients AoS [connection][up | dn][component] (column walk of the Jacobian)
  bool have_coefs = false;
};

struct RxnCouplerSet {
  RxnState *s = nullptr;
  CouplerRows R;
  int kind = 0, n = 0, device = 0;
  int32_t *d_row = nullptr, *d_own = nullptr, *d_row_ptr = nullptr, *d_conn = nullptr;
  double *d_ext = nullptr, *d_cx = nullptr, *d_cc = nullptr;   // external totals, coef_ext, coef_cell: SoA [component][connection]
  bool have_coefs = false, have_totals = false;
};

namespace {

int ensure_scratch(RxnState *s, int k, size_t bytes, void **out) {
  if (s->scratch_bytes[k] < bytes) {
    if (s->scratch[k]) cudaFree(s->scratch[k]);
    s->scratch[k] = nullptr;
    s->scratch_bytes[k] = 0;
    CU(cudaMalloc(&s->scratch[k], bytes));
    s->scratch_bytes[k] = bytes;
  }
  *out = s->scratch[k];
  return RXN_OK;
}

inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// dispatch on the compile-time bound of naqcomp (rxn_variant.cu)
#define RXN_DISPATCH(nv, FN, ...)                    \
  do {                                               \
    switch (nv) {                                    \
      case 4: FN<4>(__VA_ARGS__); break;             \
      case 8: FN<8>(__VA_ARGS__); break;             \
      case 16: FN<16>(__VA_ARGS__); break;           \
      default: FN<24>(__VA_ARGS__); break;           \
    }                                                \
    ++g_launches;                                    \
  } while (0)

// AoS/strided host image <-> SoA field.  tmp element (row r, cell c) at tmp[r*rs + c*cs].
__global__ void k_field_from_strided(double *dst, long long ld, long long ncells, int rows, const double *__restrict__ tmp,
                                     long long rs, long long cs) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  for (int r = 0; r < rows; ++r) dst[r * ld + c] = tmp[r * rs + c * cs];
}
__global__ void k_field_to_strided(const double *__restrict__ src, long long ld, long long ncells, int rows, double *tmp,
                                   long long rs, long long cs) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  for (int r = 0; r < rows; ++r) tmp[r * rs + c * cs] = src[r * ld + c];
}
__global__ void k_broadcast(double *dst, long long ld, long long ncells, int rows, const double *__restrict__ vals) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  for (int r = 0; r < rows; ++r) dst[r * ld + c] = vals[r];
}
// FP64 FMA throughput probe (roofline denominator measured on the box): 8 independent chains per thread, 4 rounds per loop
// trip so that loop control is 3 % of the issue slots (round 1's probe spent 20 % on it and read 33.8 TFLOP/s where the
// DFMA pipe delivers 36.6 = 98 % of 148 SMs x 64 lanes x 2 x 1.965 GHz, profiles/r02_ubench_tmem.txt)
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_fill(double *dst, long long n, double v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

int check_launch(RxnState *s, bool timed) {
  if (timed) {
    CU(cudaEventRecord(s->ev1, s->stream));
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  if (timed) CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
  return RXN_OK;
}

// Global-implicit entry points: the cells OR their flags into one device word (rxn_device.cuh: report_cell_flags).  The
// reference stops in these cases (activity-coefficient Newton diverged, a sorption iteration that never ends) or would go on
// with NaNs; here the call returns RXN_ERR_CELL_FAILED with the flags in rxn_last_error.
// host-buffer entry points: every l2g entry must name a cell of the state (device-pointer variants are not checked)
int check_l2g(const RxnState *s, const int32_t *l2g, int64_t nlocal) {
  if (!l2g) return RXN_OK;
  for (int64_t i = 0; i < nlocal; ++i)
    if (l2g[i] < 0 || l2g[i] >= s->ncells) return fail(RXN_ERR_INVALID, "l2g[%lld] = %d is outside the state's %lld cells", (long long)i, l2g[i], s->ncells);
  return RXN_OK;
}
int begin_cell_flags(RxnState *s) {
  if (!s->d_fail) { CU(cudaMalloc(&s->d_fail, sizeof(unsigned int))); s->S.fail = s->d_fail; }
  CU(cudaMemsetAsync(s->d_fail, 0, sizeof(unsigned int), s->stream));
  return RXN_OK;
}
int end_cell_flags(RxnState *s, const char *what) {           // after the stream has been synchronised
  unsigned int fl = 0;
  CU(cudaMemcpy(&fl, s->d_fail, sizeof fl, cudaMemcpyDeviceToHost));
  if (fl != 0)
    return fail(RXN_ERR_CELL_FAILED, "%s: at least one cell failed (flags 0x%x:%s%s%s%s)", what, fl,
                (fl & RXN_FLAG_ACT_DIVERGED) ? " activity-coefficient iteration diverged" : "",
                (fl & RXN_FLAG_CAPPED) ? " sorption iteration capped" : "", (fl & RXN_FLAG_NONFINITE) ? " non-finite result" : "",
                (fl & RXN_FLAG_LU_ZERO_ROW) ? " singular matrix" : "");
  return RXN_OK;
}

}  // namespace

extern "C" {

const char *rxn_version(void) { return "pflotran_b200 rxn 0.2 (sm_100a)"; }

int rxn_last_error(char *buf, int32_t len) {
  if (buf && len > 0) {
    strncpy(buf, g_err.c_str(), (size_t)len - 1);
    buf[len - 1] = 0;
  }
  return (int)g_err.size();
}

int64_t rxn_launch_count(void) { return g_launches; }

int rxn_tables_create(const RxnTablesDesc *d, int device, RxnTables **out) {
  if (!d || !out) return fail(RXN_ERR_INVALID, "null argument");
  *out = nullptr;
  PackResult R;
  int rc = pack_tables(d, R);
  if (rc != RXN_OK) return fail(rc, "%s", R.err.c_str());
  RxnTables *t = new RxnTables();
  t->h = R.h;
  DevTab &h = t->h;
  Packer &P = R.P;
  memcpy(t->rows, R.rows, sizeof t->rows);
  t->blob_bytes = (size_t)h.ndbl * 8 + (size_t)h.nint * 4;
  if (t->blob_bytes > 160 * 1024) {
    const size_t nb = t->blob_bytes;
    delete t;
    return fail(RXN_ERR_UNSUPPORTED, "chemistry tables (%zu bytes) exceed the shared-memory staging budget", nb);
  }
  std::vector<unsigned char> blob = blob_bytes(R);
  t->nvariant = variant_for(h.ncomp);

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { delete t; return fail(RXN_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e)); }
  if (device < 0 || device >= ndev) { delete t; return fail(RXN_ERR_INVALID, "device %d out of range (%d devices)", device, ndev); }
  t->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&t->d_blob, t->blob_bytes) != cudaSuccess ||
      cudaMemcpy(t->d_blob, blob.data(), t->blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    int rc2 = fail(RXN_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete t;
    return rc2;
  }
  rc = lane_kernel_build(h, P.d, P.i, device, &t->lane);
  if (rc != RXN_OK) { int rc2 = fail(rc, "lane plan: %s", t->lane.plan.err.c_str()); lane_kernel_free(&t->lane); cudaFree(t->d_blob); delete t; return rc2; }
  *out = t;
  return RXN_OK;
}

int rxn_tables_destroy(RxnTables *t) {
  if (!t) return RXN_OK;
  cudaSetDevice(t->device);
  if (t->d_blob) cudaFree(t->d_blob);
  lane_kernel_free(&t->lane);
  delete t;
  return RXN_OK;
}

int32_t rxn_field_rows(const RxnTables *t, int field) {
  if (!t || field < 0 || field >= RXN_F_COUNT) return -1;
  return t->rows[field];
}

static int alloc_field(RxnState *s, int f) {
  if (s->S.f[f] || s->t->rows[f] == 0) return RXN_OK;
  const size_t n = (size_t)s->t->rows[f] * s->ld;
  CU(cudaMalloc(&s->S.f[f], n * 8));
  double v = 0.0;
  if (f == RXN_F_PRI_ACT_COEF || f == RXN_F_SEC_ACT_COEF) v = 1.0;                 // reactive_transport_aux.F90:240-260
  if (f == RXN_F_FREE_SITE_CONC || f == RXN_F_EQIONX_REF_CATION_SORBED_CONC) v = 1.0e-9;   // :300, :335
  if (v == 0.0) CU(cudaMemsetAsync(s->S.f[f], 0, n * 8, s->stream));
  else { k_fill<<<nblocks((long long)n, 256), 256, 0, s->stream>>>(s->S.f[f], (long long)n, v); ++g_launches; }
  return RXN_OK;
}

int rxn_state_create(const RxnTables *t, int64_t ncells, RxnState **out) {
  if (!t || !out || ncells < 1) return fail(RXN_ERR_INVALID, "bad argument");
  *out = nullptr;
  CU(cudaSetDevice(t->device));
  RxnState *s = new RxnState();
  s->t = t;
  s->ncells = ncells;
  s->ld = (ncells + 31) / 32 * 32;
  memset(&s->S, 0, sizeof s->S);
  s->S.ld = s->ld; s->S.ncells = ncells;
  if (const char *e = getenv("RXN_REACT_KERNEL")) s->react_kernel = atoi(e);
  if (const char *e = getenv("RXN_GI_KERNEL")) s->gi_kernel = atoi(e);
  if (const char *e = getenv("RXN_FLUX_GENERIC")) s->flux_generic = atoi(e);
  if (const char *e = getenv("RXN_FLUX_ROWS")) s->flux_rows = atoi(e);
  int rc = RXN_OK;
  if (cudaStreamCreate(&s->stream) != cudaSuccess || cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess)
    rc = fail(RXN_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
  for (int f = 0; f < RXN_F_COUNT && rc == RXN_OK; ++f) {
    if (f == RXN_F_DTOTAL || f == RXN_F_DTOTAL_SORB_EQ) continue;   // materialised on demand
    rc = alloc_field(s, f);
  }
  if (rc == RXN_OK && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = fail(RXN_ERR_CUDA, "state init failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc != RXN_OK) { rxn_state_destroy(s); return rc; }
  *out = s;
  return RXN_OK;
}

int rxn_state_destroy(RxnState *s) {
  if (!s) return RXN_OK;
  cudaSetDevice(s->t->device);
  for (int f = 0; f < RXN_F_COUNT; ++f) if (s->S.f[f]) cudaFree(s->S.f[f]);
  if (s->d_active) cudaFree(s->d_active);
  if (s->d_counter) cudaFree(s->d_counter);
  if (s->d_fail) cudaFree(s->d_fail);
  for (int32_t *p : {s->d_prev_it, s->d_keys, s->d_iota, s->d_order}) if (p) cudaFree(p);
  if (s->d_sort_tmp) cudaFree(s->d_sort_tmp);
  if (s->h2d) cudaStreamDestroy(s->h2d);
  if (s->d2h) cudaStreamDestroy(s->d2h);
  if (s->stream2) cudaStreamDestroy(s->stream2);
  if (s->d_counters) cudaFree(s->d_counters);
  for (int c = 0; c < RxnState::NCHUNK; ++c) {
    if (s->ev_in[c]) cudaEventDestroy(s->ev_in[c]);
    if (s->ev_k[c]) cudaEventDestroy(s->ev_k[c]);
  }
  for (int k = 0; k < 4; ++k) if (s->scratch[k]) cudaFree(s->scratch[k]);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->tev0) cudaEventDestroy(s->tev0);
  if (s->tev1) cudaEventDestroy(s->tev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return RXN_OK;
}

int64_t rxn_state_ncells(const RxnState *s) { return s ? s->ncells : -1; }

int rxn_state_materialize(RxnState *s, int field) {
  if (!s || field < 0 || field >= RXN_F_COUNT) return fail(RXN_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(s->t->device));
  int rc = alloc_field(s, field);
  if (rc != RXN_OK) return rc;
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}

int rxn_state_upload(RxnState *s, int field, const double *host, int64_t rs, int64_t cs) {
  if (!s || !host || field < 0 || field >= RXN_F_COUNT) return fail(RXN_ERR_INVALID, "bad argument");
  const int rows = s->t->rows[field];
  if (rows == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  int rc = alloc_field(s, field);
  if (rc != RXN_OK) return rc;
  if (cs == 1) {
    CU(cudaMemcpy2DAsync(s->S.f[field], (size_t)s->ld * 8, host, (size_t)rs * 8, (size_t)s->ncells * 8, rows, cudaMemcpyHostToDevice, s->stream));
  } else {
    const size_t span = (size_t)(rows - 1) * rs + (size_t)(s->ncells - 1) * cs + 1;
    void *tmp;
    rc = ensure_scratch(s, 0, span * 8, &tmp);
    if (rc != RXN_OK) return rc;
    CU(cudaMemcpyAsync(tmp, host, span * 8, cudaMemcpyHostToDevice, s->stream));
    k_field_from_strided<<<nblocks(s->ncells, 256), 256, 0, s->stream>>>(s->S.f[field], s->ld, s->ncells, rows, (const double *)tmp, rs, cs);
    ++g_launches;
  }
  return check_launch(s, false);
}

int rxn_state_broadcast(RxnState *s, int field, const double *row_values) {
  if (!s || !row_values || field < 0 || field >= RXN_F_COUNT) return fail(RXN_ERR_INVALID, "bad argument");
  const int rows = s->t->rows[field];
  if (rows == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  int rc = alloc_field(s, field);
  if (rc != RXN_OK) return rc;
  void *tmp;
  rc = ensure_scratch(s, 1, (size_t)rows * 8, &tmp);
  if (rc != RXN_OK) return rc;
  CU(cudaMemcpyAsync(tmp, row_values, (size_t)rows * 8, cudaMemcpyHostToDevice, s->stream));
  k_broadcast<<<nblocks(s->ncells, 256), 256, 0, s->stream>>>(s->S.f[field], s->ld, s->ncells, rows, (const double *)tmp);
  ++g_launches;
  return check_launch(s, false);
}

int rxn_state_download(const RxnState *cs_, int field, double *host, int64_t rs, int64_t cs) {
  RxnState *s = const_cast<RxnState *>(cs_);
  if (!s || !host || field < 0 || field >= RXN_F_COUNT) return fail(RXN_ERR_INVALID, "bad argument");
  const int rows = s->t->rows[field];
  if (rows == 0) return RXN_OK;
  if (!s->S.f[field]) return fail(RXN_ERR_INVALID, "field %d is not materialised (call rxn_state_materialize first)", field);
  CU(cudaSetDevice(s->t->device));
  if (cs == 1) {
    CU(cudaMemcpy2DAsync(host, (size_t)rs * 8, s->S.f[field], (size_t)s->ld * 8, (size_t)s->ncells * 8, rows, cudaMemcpyDeviceToHost, s->stream));
  } else {
    const size_t span = (size_t)(rows - 1) * rs + (size_t)(s->ncells - 1) * cs + 1;
    void *tmp;
    int rc = ensure_scratch(s, 0, span * 8, &tmp);
    if (rc != RXN_OK) return rc;
    // strided destinations may interleave with caller data: start from the caller's bytes
    CU(cudaMemcpyAsync(tmp, host, span * 8, cudaMemcpyHostToDevice, s->stream));
    k_field_to_strided<<<nblocks(s->ncells, 256), 256, 0, s->stream>>>(s->S.f[field], s->ld, s->ncells, rows, (double *)tmp, rs, cs);
    ++g_launches;
    CU(cudaMemcpyAsync(host, tmp, span * 8, cudaMemcpyDeviceToHost, s->stream));
  }
  return check_launch(s, false);
}

int rxn_set_cell_scalars(RxnState *s, const double *den_kg, const double *sat, const double *temp, const double *pres,
                         const double *volume, const double *porosity, const double *soil_particle_density,
                         const uint8_t *active) {
  if (!s) return fail(RXN_ERR_INVALID, "null state");
  CU(cudaSetDevice(s->t->device));
  const double *src[7] = {den_kg, sat, temp, pres, volume, porosity, soil_particle_density};
  const int fld[7] = {RXN_F_DEN_KG, RXN_F_SAT, RXN_F_TEMP, RXN_F_PRES, RXN_F_VOLUME, RXN_F_POROSITY, RXN_F_SOIL_PARTICLE_DENSITY};
  for (int k = 0; k < 7; ++k)
    if (src[k]) CU(cudaMemcpyAsync(s->S.f[fld[k]], src[k], (size_t)s->ncells * 8, cudaMemcpyHostToDevice, s->stream));
  if (active) {
    if (!s->d_active) CU(cudaMalloc(&s->d_active, (size_t)s->ld));
    CU(cudaMemcpyAsync(s->d_active, active, (size_t)s->ncells, cudaMemcpyHostToDevice, s->stream));
    s->S.active = s->d_active;
  }
  return check_launch(s, false);
}

int rxn_set_react_kernel(RxnState *s, int which) {
  if (!s || which < 0 || which > 3) return fail(RXN_ERR_INVALID, "bad argument");
  s->react_kernel = which;
  return RXN_OK;
}

static bool tail_bound_info(const RxnTables *t) { return !t->lane.plan_tm.usable && t->h.naq > 16; }
int rxn_react_kernel_info(const RxnState *s, char *buf, int32_t len) {
  if (!s || !buf || len < 1) return fail(RXN_ERR_INVALID, "bad argument");
  const RxnTables *t = s->t;
  const bool no_dtotal = !s->S.f[RXN_F_DTOTAL] && !s->S.f[RXN_F_DTOTAL_SORB_EQ];
  const bool lane_ok = t->lane.plan.usable && no_dtotal;
  if (lane_ok && t->lane.plan_tm.usable && (s->react_kernel == 0 || s->react_kernel == 3))
    snprintf(buf, (size_t)len, "tensor-memory N=%d cells/CTA=%d warps/cell=%d threads=%d smem=%zu B plan=%zu B J in TMEM (spec %d, planA %d, planB %d terms)%s",
             t->lane.plan_tm.lt.N, t->lane.plan_tm.lt.CPB, t->lane.G_tm, 128 * t->lane.G_tm, t->lane.plan_tm.smem_bytes,
             t->lane.plan_tm.blob.size(), t->lane.plan_tm.terms_spec, t->lane.plan_tm.terms_A, t->lane.plan_tm.terms_B,
             getenv("RXN_NO_REACT_ORDER") ? "" : tail_bound_info(t) ? " work order: previous call's iteration counts, slowest first"
             : t->h.naq >= 8 ? " work order below 16 generations of resident cells: previous call's iteration counts, slowest first" : "");
  else if (lane_ok && (s->react_kernel == 0 || s->react_kernel == 3))
    snprintf(buf, (size_t)len, "resident-lane N=%d cells/CTA=%d lanes/cell=%d threads=%d smem=%zu B plan=%zu B (spec %d, planA %d, planB %d terms)%s",
             t->lane.plan.lt.N, t->lane.plan.lt.CPB, t->lane.G, ((t->lane.plan.lt.CPB * t->lane.G + 31) / 32) * 32, t->lane.plan.smem_bytes,
             t->lane.plan.blob.size(), t->lane.plan.terms_spec, t->lane.plan.terms_A, t->lane.plan.terms_B,
             getenv("RXN_NO_REACT_ORDER") ? "" : tail_bound_info(t) ? " work order: previous call's iteration counts, slowest first"
             : t->h.naq >= 8 ? " work order below 16 generations of resident cells: previous call's iteration counts, slowest first" : "");
  else
    snprintf(buf, (size_t)len, "thread-per-cell N<=%d (lane: %s)", t->nvariant,
             t->lane.plan.usable ? "DTOTAL materialised" : t->lane.plan.err.c_str());
  return RXN_OK;
}

// Work order of the on-chip RReact kernels.  The lanes take the items of a launch from a counter; a cell that needed many Newton
// iterations in one transport step needs them in the next one too, so the launch can be handed out sorted by the iteration counts
// of the previous call, slowest first (longest-processing-time-first on the persistent lanes): the long cells start while the SMs
// are full and the launch ends when the work does.  A stable 14-bit radix sort (cub) keeps cells with equal counts in index order.
// Results do not depend on the order (cells are independent).  What it costs: neighbouring lanes no longer hold neighbouring
// cells, so their 8-byte accesses stop sharing DRAM sectors.  Measured (profiles/bench_order_prediction.py, r02_av_* / r02_aw_*:
// every call ordered by the counts of a DIFFERENT noise realisation of the same cells - an imperfect prediction, as in a transport
// run - beside the exact prediction of a benchmark that repeats its inputs):
//   300A    5*10^4 / 10^5 / 3*10^5 / 10^6 / 2*10^6 cells: x1.06 / 1.10 / 1.03 / 0.95 / 0.92  (exact prediction: x1.14 / 1.29 / 1.21 / 1.16 / 1.14)
//   calcite 10^5 / 4*10^6: x0.95 / 0.91;   ascem (tail-bound) 2*10^5 / 10^6: x1.07 / 1.13  (exact: x1.00 / 1.58)
// Hence the policy (react_ordered): tail-bound chemistries always; chemistries with at least 8 primaries on batches below 16
// generations of resident cells (300A: 3*10^5 cells - a rank's share of a grid), where removing the tail pays even with the imperfect
// prediction; large batches and small chemistries are taken in index order.  The gain of an exact prediction on large batches
// (300A, 10^7 cells: 57.9 -> 65.7 M cell-updates/s, profiles/r02_au_bench_default.json: warps whose 32 cells finish in the same trip
// run their finish / load rounds at full width) is an artefact of repeated inputs and is not taken.
__global__ void k_iota(int32_t *p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}
static bool tail_bound_tables(const RxnTables *t) { return !t->lane.plan_tm.usable && t->h.naq > 16; }
static int react_order(RxnState *s, int64_t nlocal, cudaStream_t stream, const int32_t **order) {
  *order = nullptr;
  if (s->prev_n != nlocal || getenv("RXN_NO_REACT_ORDER")) return RXN_OK;
  size_t need = 0;
  CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, s->d_prev_it, s->d_keys, s->d_iota, s->d_order, (int)nlocal, 0, 14, stream));
  if (need > s->sort_tmp_bytes) {
    if (s->d_sort_tmp) CU(cudaFree(s->d_sort_tmp));
    s->d_sort_tmp = nullptr; s->sort_tmp_bytes = 0;
    CU(cudaMalloc(&s->d_sort_tmp, need));
    s->sort_tmp_bytes = need;
  }
  k_iota<<<nblocks(nlocal, 256), 256, 0, stream>>>(s->d_iota, nlocal);
  CU(cub::DeviceRadixSort::SortPairsDescending(s->d_sort_tmp, need, s->d_prev_it, s->d_keys, s->d_iota, s->d_order, (int)nlocal, 0, 14, stream));
  g_launches += 2;
  *order = s->d_order;
  return RXN_OK;
}
// room for the iteration counts of a launch of nlocal items (drops what was remembered when it has to grow)
static int react_order_reserve(RxnState *s, int64_t nlocal) {
  if (nlocal <= s->order_cap) return RXN_OK;
  s->prev_n = 0;
  for (int32_t **p : {&s->d_prev_it, &s->d_keys, &s->d_iota, &s->d_order}) { if (*p) CU(cudaFree(*p)); *p = nullptr; }
  s->order_cap = 0;
  for (int32_t **p : {&s->d_prev_it, &s->d_keys, &s->d_iota, &s->d_order}) CU(cudaMalloc(p, (size_t)nlocal * 4));
  s->order_cap = nlocal;
  return RXN_OK;
}
// see the policy above
static bool react_ordered(const RxnTables *t, int64_t nlocal) {
  if (tail_bound_tables(t)) return true;
  const LaneTab &klt = t->lane.plan_tm.usable ? t->lane.plan_tm.lt : t->lane.plan.lt;
  return t->h.naq >= 8 && nlocal < 16LL * t->lane.sm_count * klt.CPB;
}

static int launch_react(RxnState *s, double *d_xx, const int32_t *d_l2g, int64_t nlocal, double dt, int dt_mode,
                        int32_t *d_iters, int32_t *d_flags, long long cell0 = 0, cudaStream_t stream = nullptr,
                        unsigned long long *counter = nullptr) {
  const RxnTables *t = s->t;
  if (!stream) stream = s->stream;
  // the shared-memory kernels keep dtotal only as Newton scratch: states with DTOTAL materialised use thread-per-cell
  const bool no_dtotal = !s->S.f[RXN_F_DTOTAL] && !s->S.f[RXN_F_DTOTAL_SORB_EQ];
  const bool lane_ok = t->lane.plan.usable && no_dtotal;
  if (s->react_kernel == 2) return fail(RXN_ERR_UNSUPPORTED, "the cooperative kernel of round 1 has been removed (use 0, 1 or 3)");
  if (s->react_kernel == 3 && !lane_ok)
    return fail(RXN_ERR_UNSUPPORTED, "resident-lane kernel unavailable: %s", t->lane.plan.usable ? "DTOTAL is materialised" : t->lane.plan.err.c_str());
  const bool use_lane = lane_ok && (s->react_kernel == 0 || s->react_kernel == 3);
  if (use_lane) {
    const bool single_launch = counter == nullptr;            // not a chunk of the pipelined host-buffer call
    if (!counter) {
      if (!s->d_counter) CU(cudaMalloc(&s->d_counter, sizeof(unsigned long long)));
      counter = s->d_counter;
    }
    DevState S = s->S;
    // work order (react_order; policy: react_ordered): single launches only, never the chunks of the pipelined call
    const bool ordered = single_launch && d_iters != nullptr && nlocal <= 0x7fffffffLL && react_ordered(t, nlocal);
    if (ordered) {
      const int rcv = react_order_reserve(s, nlocal); if (rcv != RXN_OK) return rcv;
      const int rco = react_order(s, nlocal, stream, &S.order); if (rco != RXN_OK) return rco;
    }
    int rc = lane_launch_react(t->lane, t->h, t->d_blob, S, d_xx, d_l2g, nlocal, dt, dt_mode, d_iters, d_flags, counter, stream, cell0);
    if (single_launch) s->prev_n = 0;
    if (rc == RXN_OK && ordered) {
      CU(cudaMemcpyAsync(s->d_prev_it, d_iters, (size_t)nlocal * 4, cudaMemcpyDeviceToDevice, stream));
      s->prev_n = nlocal;
    }
    if (rc != RXN_OK) return fail(rc, "resident-lane kernel launch failed (N=%d CPB=%d): %s", t->lane.plan.lt.N, t->lane.plan.lt.CPB,
                                  cudaGetErrorString(cudaGetLastError()));
    ++g_launches;
    return RXN_OK;
  }
  {
    int threads = t->nvariant <= 8 ? 128 : 64;
    if (const char *e = getenv("RXN_TPC_BLOCK")) threads = std::max(32, atoi(e));
    unsigned grid = nblocks(nlocal, threads);
    if (const char *e = getenv("RXN_TPC_GRID")) grid = std::min<unsigned>(grid, (unsigned)std::max(1, atoi(e)));
    const LaunchCfg L{grid, threads, t->blob_bytes, stream};
    RXN_DISPATCH(t->nvariant, run_react, L, t->h, (const double *)t->d_blob, s->S, d_xx, d_l2g, (long long)nlocal, dt, dt_mode,
                 d_iters, d_flags);
  }
  return RXN_OK;
}

int rxn_react_batch_device(RxnState *s, double *d_xx, const int32_t *d_l2g, int64_t nlocal, double dt, int dt_mode,
                           int32_t *d_iters, int32_t *d_flags) {
  Nvtx nvtx_("RTReact");
  if (!s || !d_xx || nlocal < 0 || !(dt > 0.0)) return fail(RXN_ERR_INVALID, "bad argument");
  if (nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  CU(cudaEventRecord(s->ev0, s->stream));
  int rc = launch_react(s, d_xx, d_l2g, nlocal, dt, dt_mode, d_iters, d_flags);
  if (rc != RXN_OK) return rc;
  return check_launch(s, true);
}

int rxn_react_batch(RxnState *s, double *tran_xx, const int32_t *l2g, int64_t nlocal, double dt, int dt_mode,
                    int32_t *iters_out, int32_t *flags_out) {
  Nvtx nvtx_("RTReact");
  if (!s || !tran_xx || nlocal < 0 || !(dt > 0.0)) return fail(RXN_ERR_INVALID, "bad argument");
  if (nlocal == 0) return RXN_OK;
  { const int rcg = check_l2g(s, l2g, nlocal); if (rcg != RXN_OK) return rcg; }
  CU(cudaSetDevice(s->t->device));
  const int n = s->t->h.ncomp;
  void *d_xx, *d_l2g = nullptr, *d_it, *d_fl;
  int rc;
  if ((rc = ensure_scratch(s, 0, (size_t)nlocal * n * 8, &d_xx)) != RXN_OK) return rc;
  if ((rc = ensure_scratch(s, 2, (size_t)nlocal * 4, &d_it)) != RXN_OK) return rc;
  if ((rc = ensure_scratch(s, 3, (size_t)nlocal * 4, &d_fl)) != RXN_OK) return rc;
  if (l2g) {
    if ((rc = ensure_scratch(s, 1, (size_t)nlocal * 4, &d_l2g)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_l2g, l2g, (size_t)nlocal * 4, cudaMemcpyHostToDevice, s->stream));
  }
  const RxnTables *t = s->t;
  const bool lane = t->lane.plan.usable && !s->S.f[RXN_F_DTOTAL] && !s->S.f[RXN_F_DTOTAL_SORB_EQ] &&
                    (s->react_kernel == 0 || s->react_kernel == 3);
  // chemistries on the N = 24 shapes (more than 16 primaries) are not chunked: a launch ends with the tail of its slowest cells
  // (ascem: damped redox cells with thousands of Newton iterations, 81 % of a 100 000-cell launch,
  // profiles/r02_ac2_ascem_lane_g8.metrics.txt), every chunk would pay it again, and the copies are 2 % of the kernel time
  const bool tail_bound = tail_bound_tables(t);
  if (lane && nlocal >= 262144 && !tail_bound && !getenv("RXN_NO_PIPELINE")) {
    // resident-lane kernel on a large batch: NCHUNK chunks; chunk c+1 crosses PCIe while chunk c is solved and chunk
    // c-1 returns (full duplex), so the host-buffer call costs about the kernel time
    if (!s->h2d) {
      CU(cudaStreamCreate(&s->h2d)); CU(cudaStreamCreate(&s->d2h)); CU(cudaStreamCreate(&s->stream2));
      CU(cudaMalloc(&s->d_counters, RxnState::NCHUNK * sizeof(unsigned long long)));
      for (int c = 0; c < RxnState::NCHUNK; ++c) {
        CU(cudaEventCreateWithFlags(&s->ev_in[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_k[c], cudaEventDisableTiming));
      }
    }
    // an error inside the pipeline must not leave copies or kernels running on the caller's tran_xx / iters / flags:
    // drain the three streams before returning
    auto drain = [&]() { cudaStreamSynchronize(s->h2d); cudaStreamSynchronize(s->stream); cudaStreamSynchronize(s->stream2); cudaStreamSynchronize(s->d2h); };
#define CUP(call)                                                                              \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      drain();                                                                                  \
      return fail(RXN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    }                                                                                           \
  } while (0)
    CUP(cudaStreamSynchronize(s->stream));                       // the l2g copy above / earlier work on the scratch buffers
    // up to NCHUNK chunks of at least 131 072 cells (several generations of resident cells per launch)
    const int64_t nchunks = std::min<int64_t>(RxnState::NCHUNK, std::max<int64_t>(2, nlocal / 131072));
    const int64_t chunk = (((nlocal + nchunks - 1) / nchunks) + 63) / 64 * 64;
    int nch = 0;
    for (int64_t off = 0; off < nlocal; off += chunk, ++nch) {
      const int64_t len = std::min<int64_t>(chunk, nlocal - off);
      CUP(cudaMemcpyAsync((double *)d_xx + off * n, tran_xx + off * n, (size_t)len * n * 8, cudaMemcpyHostToDevice, s->h2d));
      CUP(cudaEventRecord(s->ev_in[nch], s->h2d));
    }
    CUP(cudaEventRecord(s->ev0, s->stream));
    int c = 0;
    int last_k[2] = {-1, -1};
    for (int64_t off = 0; off < nlocal; off += chunk, ++c) {
      const int64_t len = std::min<int64_t>(chunk, nlocal - off);
      cudaStream_t ks = (c & 1) ? s->stream2 : s->stream;       // chunk c+1 fills the SMs chunk c has drained
      CUP(cudaStreamWaitEvent(ks, s->ev_in[c], 0));
      rc = launch_react(s, (double *)d_xx + off * n, d_l2g ? (const int32_t *)d_l2g + off : nullptr, len, dt, dt_mode, (int32_t *)d_it + off,
                        (int32_t *)d_fl + off, d_l2g ? 0 : off, ks, s->d_counters + c);
      if (rc != RXN_OK) { drain(); return rc; }
      CUP(cudaEventRecord(s->ev_k[c], ks));
      last_k[c & 1] = c;
      CUP(cudaStreamWaitEvent(s->d2h, s->ev_k[c], 0));
      CUP(cudaMemcpyAsync(tran_xx + off * n, (double *)d_xx + off * n, (size_t)len * n * 8, cudaMemcpyDeviceToHost, s->d2h));
      if (iters_out) CUP(cudaMemcpyAsync(iters_out + off, (int32_t *)d_it + off, (size_t)len * 4, cudaMemcpyDeviceToHost, s->d2h));
      if (flags_out) CUP(cudaMemcpyAsync(flags_out + off, (int32_t *)d_fl + off, (size_t)len * 4, cudaMemcpyDeviceToHost, s->d2h));
    }
    if (last_k[1] >= 0) CUP(cudaStreamWaitEvent(s->stream, s->ev_k[last_k[1]], 0));   // ev1 after the kernels of both streams
    CUP(cudaEventRecord(s->ev1, s->stream));
    CUP(cudaGetLastError());
    CUP(cudaStreamSynchronize(s->d2h));
    CUP(cudaStreamSynchronize(s->stream2));
    CUP(cudaStreamSynchronize(s->stream));
    CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
#undef CUP
    return RXN_OK;
  }
  CU(cudaMemcpyAsync(d_xx, tran_xx, (size_t)nlocal * n * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaEventRecord(s->ev0, s->stream));
  rc = launch_react(s, (double *)d_xx, (const int32_t *)d_l2g, nlocal, dt, dt_mode, (int32_t *)d_it, (int32_t *)d_fl);
  if (rc != RXN_OK) return rc;
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaMemcpyAsync(tran_xx, d_xx, (size_t)nlocal * n * 8, cudaMemcpyDeviceToHost, s->stream));
  if (iters_out) CU(cudaMemcpyAsync(iters_out, d_it, (size_t)nlocal * 4, cudaMemcpyDeviceToHost, s->stream));
  if (flags_out) CU(cudaMemcpyAsync(flags_out, d_fl, (size_t)nlocal * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
  return RXN_OK;
}

// RTUpdateAuxVars / RTUpdateFixedAccumulation cell loops: the tensor-memory layout when the tables allow it (tm_gi_cell,
// rxn_tm_dev.cuh), else one thread per cell (rxn_device.cuh)
static int launch_update_auxvars(RxnState *s, const double *d_xx, int update_act_coefs) {
  const RxnTables *t = s->t;
  if (s->gi_kernel == 0 && tm_gi_usable(t->lane, update_act_coefs)) {
    const GiArgs a{GI_AUX, update_act_coefs, d_xx, 0, nullptr, nullptr, nullptr, 1.0};
    int rc = tm_launch_gi(t->lane, t->h, t->d_blob, s->S, nullptr, s->ncells, a, s->stream);
    if (rc != RXN_OK) return fail(rc, "tensor-memory auxvar kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ++g_launches;
    return RXN_OK;
  }
  const int threads = t->nvariant <= 8 ? 128 : 64;
  const LaunchCfg L{nblocks(s->ncells, threads), threads, t->blob_bytes, s->stream};
  RXN_DISPATCH(t->nvariant, run_update_auxvars, L, t->h, (const double *)t->d_blob, s->S, d_xx, update_act_coefs);
  return RXN_OK;
}

static int launch_fixed_accum(RxnState *s, const double *d_xx, const int32_t *d_l2g, int64_t nlocal, double *d_out) {
  const RxnTables *t = s->t;
  if (s->gi_kernel == 0 && tm_gi_usable(t->lane, 0)) {
    const GiArgs a{GI_AUX, 0, d_xx, 1, d_out, nullptr, nullptr, 1.0};
    int rc = tm_launch_gi(t->lane, t->h, t->d_blob, s->S, d_l2g, (long long)nlocal, a, s->stream);
    if (rc != RXN_OK) return fail(rc, "tensor-memory accumulation kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ++g_launches;
    return RXN_OK;
  }
  const int threads = t->nvariant <= 8 ? 128 : 64;
  const LaunchCfg L{nblocks(nlocal, threads), threads, t->blob_bytes, s->stream};
  RXN_DISPATCH(t->nvariant, run_fixed_accum, L, t->h, (const double *)t->d_blob, s->S, d_xx, (const int *)d_l2g, (long long)nlocal, d_out);
  return RXN_OK;
}

int rxn_update_auxvars_batch(RxnState *s, const double *xx_loc, int update_act_coefs) {
  Nvtx nvtx_("RTAuxVars");
  if (!s) return fail(RXN_ERR_INVALID, "null state");
  CU(cudaSetDevice(s->t->device));
  const RxnTables *t = s->t;
  void *d_xx = nullptr;
  if (xx_loc) {
    int rc = ensure_scratch(s, 0, (size_t)s->ncells * t->h.ncomp * 8, &d_xx);
    if (rc != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_xx, xx_loc, (size_t)s->ncells * t->h.ncomp * 8, cudaMemcpyHostToDevice, s->stream));
  }
  { const int rcf = begin_cell_flags(s); if (rcf != RXN_OK) return rcf; }
  CU(cudaEventRecord(s->ev0, s->stream));
  { const int rcu = launch_update_auxvars(s, (const double *)d_xx, update_act_coefs); if (rcu != RXN_OK) return rcu; }
  { const int rcl = check_launch(s, true); if (rcl != RXN_OK) return rcl; }
  return end_cell_flags(s, "RTUpdateAuxVars");
}

int rxn_fixed_accum_batch(RxnState *s, const double *xx, const int32_t *l2g, int64_t nlocal, double *accum_out) {
  Nvtx nvtx_("RTUpdateFixedAccumulation");
  if (!s || !accum_out || nlocal < 0) return fail(RXN_ERR_INVALID, "bad argument");
  if (nlocal == 0) return RXN_OK;
  { const int rcg = check_l2g(s, l2g, nlocal); if (rcg != RXN_OK) return rcg; }
  CU(cudaSetDevice(s->t->device));
  const RxnTables *t = s->t;
  const int n = t->h.ncomp;
  void *d_xx = nullptr, *d_l2g = nullptr, *d_out;
  int rc;
  if ((rc = ensure_scratch(s, 2, (size_t)nlocal * n * 8, &d_out)) != RXN_OK) return rc;
  if (xx) {
    if ((rc = ensure_scratch(s, 0, (size_t)nlocal * n * 8, &d_xx)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_xx, xx, (size_t)nlocal * n * 8, cudaMemcpyHostToDevice, s->stream));
  }
  if (l2g) {
    if ((rc = ensure_scratch(s, 1, (size_t)nlocal * 4, &d_l2g)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_l2g, l2g, (size_t)nlocal * 4, cudaMemcpyHostToDevice, s->stream));
  }
  CU(cudaMemsetAsync(d_out, 0, (size_t)nlocal * n * 8, s->stream));
  { const int rcf = begin_cell_flags(s); if (rcf != RXN_OK) return rcf; }
  CU(cudaEventRecord(s->ev0, s->stream));
  if ((rc = launch_fixed_accum(s, (const double *)d_xx, (const int32_t *)d_l2g, nlocal, (double *)d_out)) != RXN_OK) return rc;
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaMemcpyAsync(accum_out, d_out, (size_t)nlocal * n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
  return end_cell_flags(s, "RTUpdateFixedAccumulation");
}

static int launch_residual_jacobian(RxnState *s, const int32_t *d_l2g, int64_t nlocal, double dt, double *d_res, double *d_jac) {
  const RxnTables *t = s->t;
  if (s->gi_kernel == 0 && tm_gi_usable(t->lane, 0)) {
    const GiArgs a{GI_RJ, 0, nullptr, 0, nullptr, d_res, d_jac, dt};
    int rc = tm_launch_gi(t->lane, t->h, t->d_blob, s->S, d_l2g, (long long)nlocal, a, s->stream);
    if (rc != RXN_OK) return fail(rc, "tensor-memory residual/Jacobian kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ++g_launches;
  } else if (t->lane.plan_gi.usable && s->gi_kernel != 1) {
    int rc = lane_launch_gi(t->lane, t->h, t->d_blob, s->S, d_l2g, (long long)nlocal, dt, d_res, d_jac, s->stream);
    if (rc != RXN_OK) return fail(rc, "resident-lane residual/Jacobian kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ++g_launches;
  } else {
    const int threads = t->nvariant <= 8 ? 128 : 64;
    const LaunchCfg L{nblocks(nlocal, threads), threads, t->blob_bytes, s->stream};
    RXN_DISPATCH(t->nvariant, run_residual_jacobian, L, t->h, (const double *)t->d_blob, s->S, (const int *)d_l2g, (long long)nlocal,
                 dt, d_res, d_jac);
  }
  return RXN_OK;
}

// device-resident variant (PETSc VECCUDA / MATAIJCUSPARSE arrays): d_res / d_jac are device pointers the
// caller owns; blocks of inactive cells are left untouched
int rxn_residual_jacobian_blocks_batch_device(RxnState *s, const int32_t *d_l2g, int64_t nlocal, double dt, double *d_res,
                                              double *d_jac) {
  Nvtx nvtx_("RTResReaction+RTJacReaction+RTJacobianAccum");
  if (!s || nlocal < 0 || !(dt > 0.0) || (!d_res && !d_jac)) return fail(RXN_ERR_INVALID, "bad argument");
  if (nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  { const int rcf = begin_cell_flags(s); if (rcf != RXN_OK) return rcf; }
  CU(cudaEventRecord(s->ev0, s->stream));
  int rc = launch_residual_jacobian(s, d_l2g, nlocal, dt, d_res, d_jac);
  if (rc != RXN_OK) return rc;
  { const int rcl = check_launch(s, true); if (rcl != RXN_OK) return rcl; }
  return end_cell_flags(s, "RTResidual/RTJacobian blocks");
}

int rxn_update_auxvars_batch_device(RxnState *s, const double *d_xx_loc, int update_act_coefs) {
  Nvtx nvtx_("RTAuxVars");
  if (!s) return fail(RXN_ERR_INVALID, "null state");
  CU(cudaSetDevice(s->t->device));
  { const int rcf = begin_cell_flags(s); if (rcf != RXN_OK) return rcf; }
  CU(cudaEventRecord(s->ev0, s->stream));
  { const int rcu = launch_update_auxvars(s, d_xx_loc, update_act_coefs); if (rcu != RXN_OK) return rcu; }
  { const int rcl = check_launch(s, true); if (rcl != RXN_OK) return rcl; }
  return end_cell_flags(s, "RTUpdateAuxVars");
}

int rxn_residual_jacobian_blocks_batch(RxnState *s, const int32_t *l2g, int64_t nlocal, double dt, double *res_out,
                                       double *jac_out) {
  Nvtx nvtx_("RTResReaction+RTJacReaction+RTJacobianAccum");
  if (!s || nlocal < 0 || !(dt > 0.0) || (!res_out && !jac_out)) return fail(RXN_ERR_INVALID, "bad argument");
  if (nlocal == 0) return RXN_OK;
  { const int rcg = check_l2g(s, l2g, nlocal); if (rcg != RXN_OK) return rcg; }
  CU(cudaSetDevice(s->t->device));
  const RxnTables *t = s->t;
  const int n = t->h.ncomp;
  void *d_res = nullptr, *d_jac = nullptr, *d_l2g = nullptr;
  int rc;
  if (res_out) { if ((rc = ensure_scratch(s, 2, (size_t)nlocal * n * 8, &d_res)) != RXN_OK) return rc; CU(cudaMemsetAsync(d_res, 0, (size_t)nlocal * n * 8, s->stream)); }
  if (jac_out) { if ((rc = ensure_scratch(s, 0, (size_t)nlocal * n * n * 8, &d_jac)) != RXN_OK) return rc; CU(cudaMemsetAsync(d_jac, 0, (size_t)nlocal * n * n * 8, s->stream)); }
  if (l2g) {
    if ((rc = ensure_scratch(s, 1, (size_t)nlocal * 4, &d_l2g)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_l2g, l2g, (size_t)nlocal * 4, cudaMemcpyHostToDevice, s->stream));
  }
  { const int rcf = begin_cell_flags(s); if (rcf != RXN_OK) return rcf; }
  CU(cudaEventRecord(s->ev0, s->stream));
  if ((rc = launch_residual_jacobian(s, (const int32_t *)d_l2g, nlocal, dt, (double *)d_res, (double *)d_jac)) != RXN_OK) return rc;
  CU(cudaEventRecord(s->ev1, s->stream));
  if (res_out) CU(cudaMemcpyAsync(res_out, d_res, (size_t)nlocal * n * 8, cudaMemcpyDeviceToHost, s->stream));
  if (jac_out) CU(cudaMemcpyAsync(jac_out, d_jac, (size_t)nlocal * n * n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
  return end_cell_flags(s, "RTResidual/RTJacobian blocks");
}

int rxn_residual_blocks_batch(RxnState *s, const int32_t *l2g, int64_t nlocal, double dt, double *res_out) {
  if (!res_out) return fail(RXN_ERR_INVALID, "null res_out");
  return rxn_residual_jacobian_blocks_batch(s, l2g, nlocal, dt, res_out, nullptr);
}
int rxn_jacobian_blocks_batch(RxnState *s, const int32_t *l2g, int64_t nlocal, double dt, double *jac_out) {
  if (!jac_out) return fail(RXN_ERR_INVALID, "null jac_out");
  return rxn_residual_jacobian_blocks_batch(s, l2g, nlocal, dt, nullptr, jac_out);
}

int rxn_equilibrate_constraint_batch(RxnState *s, const int32_t *constraint_type, const double *constraint_conc, int64_t conc_stride,
                                     const int32_t *constraint_id, const double *free_ion_guess, int use_prev_soln_as_guess,
                                     int initialize_with_molality, const int32_t *l2g, int64_t nlocal, double *basis_molarity_out,
                                     int32_t *iters_out, int32_t *status_out) {
  Nvtx nvtx_("ReactionEquilibrateConstraint");
  if (!s || !constraint_type || !constraint_conc || !constraint_id || nlocal < 0) return fail(RXN_ERR_INVALID, "bad argument");
  const RxnTables *t = s->t;
  const int n = t->h.naq;
  if (conc_stride != 0 && conc_stride < n) return fail(RXN_ERR_INVALID, "conc_stride must be 0 or >= naqcomp");
  if (nlocal == 0) return RXN_OK;
  { const int rcg = check_l2g(s, l2g, nlocal); if (rcg != RXN_OK) return rcg; }
  for (int i = 0; i < n; ++i) {
    if (constraint_type[i] == RXN_CONSTRAINT_MINERAL && (constraint_id[i] < 1 || constraint_id[i] > t->h.mnrl.n))
      return fail(RXN_ERR_INVALID, "constraint %d: mineral id %d out of range", i + 1, constraint_id[i]);
    if (constraint_type[i] == RXN_CONSTRAINT_GAS && (constraint_id[i] < 1 || constraint_id[i] > t->h.gas.n))
      return fail(RXN_ERR_INVALID, "constraint %d: gas id %d out of range", i + 1, constraint_id[i]);
  }
  CU(cudaSetDevice(t->device));
  // scratch 0: [conc | basis]   scratch 1: l2g   scratch 2: [ctype | cid | iters | status]   scratch 3: guess
  const size_t nconc = conc_stride ? (size_t)nlocal * conc_stride : (size_t)n;
  void *d_dbl, *d_l2g = nullptr, *d_int, *d_guess = nullptr;
  int rc;
  if ((rc = ensure_scratch(s, 0, (nconc + (size_t)nlocal * n) * 8, &d_dbl)) != RXN_OK) return rc;
  if ((rc = ensure_scratch(s, 2, ((size_t)2 * n + 2 * (size_t)nlocal) * 4, &d_int)) != RXN_OK) return rc;
  double *d_conc = (double *)d_dbl, *d_basis = d_conc + nconc;
  int32_t *d_ctype = (int32_t *)d_int, *d_cid = d_ctype + n, *d_it = d_cid + n, *d_st = d_it + nlocal;
  CU(cudaMemcpyAsync(d_conc, constraint_conc, nconc * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(d_ctype, constraint_type, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(d_cid, constraint_id, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
  if (free_ion_guess) {
    if ((rc = ensure_scratch(s, 3, (size_t)n * 8, &d_guess)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_guess, free_ion_guess, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
  }
  if (l2g) {
    if ((rc = ensure_scratch(s, 1, (size_t)nlocal * 4, &d_l2g)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(d_l2g, l2g, (size_t)nlocal * 4, cudaMemcpyHostToDevice, s->stream));
  }
  CU(cudaEventRecord(s->ev0, s->stream));
  const int threads = t->nvariant <= 8 ? 128 : 64;
  const LaunchCfg L{nblocks(nlocal, threads), threads, t->blob_bytes, s->stream};
  RXN_DISPATCH(t->nvariant, run_equilibrate, L, t->h, (const double *)t->d_blob, s->S, (const int *)d_ctype, (const double *)d_conc,
               (long long)conc_stride, (const int *)d_cid, (const double *)d_guess, use_prev_soln_as_guess, initialize_with_molality,
               (const int *)d_l2g, (long long)nlocal, d_basis, (int *)d_it, (int *)d_st);
  CU(cudaEventRecord(s->ev1, s->stream));
  if (basis_molarity_out) CU(cudaMemcpyAsync(basis_molarity_out, d_basis, (size_t)nlocal * n * 8, cudaMemcpyDeviceToHost, s->stream));
  if (iters_out) CU(cudaMemcpyAsync(iters_out, d_it, (size_t)nlocal * 4, cudaMemcpyDeviceToHost, s->stream));
  if (status_out) CU(cudaMemcpyAsync(status_out, d_st, (size_t)nlocal * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
  return RXN_OK;
}

// Multirate sorbed totals of RUpdateKineticState (reaction.F90:5394-5408): S_r <- (S_r + k_r dt f_r S_eq) / (1 + k_r dt) for every rate r
// and component i - 2 x nrate x naq doubles of HBM traffic per cell and one quotient per element, no other state.  One thread
// per (cell, component) walks the rates, cells fastest: every load and store is a full 128-byte line of one row of the SoA field,
// S_eq is read once, and the loads of 4 rates are in flight before the first quotient.  The thread-per-cell kernel (which kept
// the whole cell context in local memory for the mineral rates) ran this loop at 39 % of the HBM peak (profiles/r02_q_kinstate_mr.json).
__global__ void __launch_bounds__(256)
k_kinmr_update(DevState S, const double *__restrict__ rate, const double *__restrict__ frac, int naq, int nrate, long long row0, double dt) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (cell >= S.ncells || (S.active && !S.active[cell])) return;
  double *f = S.f[RXN_F_KINMR_TOTAL_SORB];
  const double S0 = f[(row0 + i) * S.ld + cell];
  int r = 0;
  for (; r + 4 <= nrate; r += 4) {
    double *p[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { p[u] = f + (row0 + (long long)(r + u + 1) * naq + i) * S.ld + cell; v[u] = *p[u]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double kdt = rate[r + u] * dt;
      const double one_plus_kdt = 1.0 + kdt;
      *p[u] = (v[u] + kdt * frac[r + u] * S0) / one_plus_kdt;
    }
  }
  for (; r < nrate; ++r) {
    double *p = f + (row0 + (long long)(r + 1) * naq + i) * S.ld + cell;
    const double kdt = rate[r] * dt;
    const double one_plus_kdt = 1.0 + kdt;
    *p = (*p + kdt * frac[r] * S0) / one_plus_kdt;
  }
}

int rxn_update_kinetic_state_batch(RxnState *s, double dt) {
  Nvtx nvtx_("RTUpdateKineticState");
  if (!s || !(dt > 0.0)) return fail(RXN_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(s->t->device));
  const RxnTables *t = s->t;
  { const int rcf = begin_cell_flags(s); if (rcf != RXN_OK) return rcf; }
  CU(cudaEventRecord(s->ev0, s->stream));
  const int threads = 128;
  const LaunchCfg L{nblocks(s->ncells, threads), threads, t->blob_bytes, s->stream};
  const bool stream_mr = t->h.nmr > 0 && !getenv("RXN_KINMR_PER_CELL");
  if (t->h.nkin > 0 || t->h.nkinrxn > 0 || !stream_mr)       // mineral volume fractions, kinetic surface complexes (and the flags of their cells)
    RXN_DISPATCH(t->nvariant, run_update_kinetic_state, L, t->h, (const double *)t->d_blob, s->S, dt, stream_mr ? 1 : 0);
  if (stream_mr) {
    const double *bd = (const double *)t->d_blob;
    for (int ikr = 0; ikr < t->h.nmr; ++ikr) {
      const int nrate = t->lane.mr_nrate[ikr];
      if (nrate <= 0) continue;
      const dim3 grid(nblocks(s->ncells, 256), (unsigned)t->h.naq);
      k_kinmr_update<<<grid, 256, 0, s->stream>>>(s->S, bd + t->h.o_mr_rate + (size_t)ikr * t->h.mr_ld, bd + t->h.o_mr_frac + (size_t)ikr * t->h.mr_ld,
                                                  t->h.naq, nrate, (long long)ikr * (t->h.mr_ld + 1) * t->h.naq, dt);
      ++g_launches;
    }
  }
  { const int rcl = check_launch(s, true); if (rcl != RXN_OK) return rcl; }
  return end_cell_flags(s, "RTUpdateKineticState");
}

int rxn_timer_start(RxnState *s) {
  if (!s) return fail(RXN_ERR_INVALID, "null state");
  CU(cudaSetDevice(s->t->device));
  if (!s->tev0) { CU(cudaEventCreate(&s->tev0)); CU(cudaEventCreate(&s->tev1)); }
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaEventRecord(s->tev0, s->stream));
  return RXN_OK;
}
int rxn_timer_stop(RxnState *s, float *ms) {
  if (!s || !ms || !s->tev0) return fail(RXN_ERR_INVALID, "timer not started");
  CU(cudaSetDevice(s->t->device));
  CU(cudaEventRecord(s->tev1, s->stream));
  CU(cudaEventSynchronize(s->tev1));
  CU(cudaEventElapsedTime(ms, s->tev0, s->tev1));
  return RXN_OK;
}

int rxn_probe_fp64(RxnState *s, double *tflops) {
  if (!s || !tflops) return fail(RXN_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(s->t->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, s->t->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 17;
  void *out;
  int rc = ensure_scratch(s, 3, (size_t)blocks * threads * 8, &out);
  if (rc != RXN_OK) return rc;
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(a, s->stream));
    k_dfma_probe<<<blocks, threads, 0, s->stream>>>((double *)out, iters, 0.999999, 1.0e-9);
    ++g_launches;
    CU(cudaEventRecord(b, s->stream));
    CU(cudaEventSynchronize(b));
    float ms;
    CU(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  return RXN_OK;
}

// ------------------------------------------------------------------ flux side (rxn_flux.h)
int rxn_connset_create(RxnState *s, int64_t nconn, const int32_t *id_up, const int32_t *id_dn, const int32_t *ghost_to_local,
                       int64_t nlocal, const uint8_t *active, RxnConnSet **out) {
  if (!s || !out || nconn < 0 || nlocal < 0 || (nconn > 0 && (!id_up || !id_dn))) return fail(RXN_ERR_INVALID, "bad argument");
  *out = nullptr;
  if (s->t->h.nim > 0)
    return fail(RXN_ERR_UNSUPPORTED, "flux-side blocks are naqcomp x naqcomp: tables with immobile dofs (ncomp %d > naqcomp %d) are not handled by the flux entry points", s->t->h.ncomp, s->t->h.naq);
  CU(cudaSetDevice(s->t->device));
  RxnConnSet *c = new RxnConnSet();
  c->s = s;
  c->n = s->t->h.naq;
  c->device = s->t->device;
  if (!flux_rows_build(s->ncells, nlocal, nconn, id_up, id_dn, ghost_to_local, active, &c->R)) {
    const std::string e = c->R.err;
    delete c;
    return fail(RXN_ERR_INVALID, "connection set: %s", e.c_str());
  }
  auto up = [&](int32_t **d, const std::vector<int32_t> &v) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, std::max<size_t>(v.size(), 1) * 4);
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    return e;
  };
  cudaError_t e = up(&c->d_row_ptr, c->R.row_ptr);
  if (e == cudaSuccess) e = up(&c->d_col, c->R.col);
  if (e == cudaSuccess) e = up(&c->d_ent, c->R.ent);
  if (e == cudaSuccess) e = up(&c->d_l2g, c->R.l2g);
  {
    FluxCols C;
    flux_cols_build(c->R, &C);
    if (e == cudaSuccess) e = up(&c->d_col_ptr, C.col_ptr);
    if (e == cudaSuccess) e = up(&c->d_tgt_slot, C.tgt_slot);
    if (e == cudaSuccess) e = up(&c->d_tgt_ent, C.tgt_ent);
    if (e == cudaSuccess) e = up(&c->d_col_row, C.col_row);
  }
  if (e == cudaSuccess) e = cudaMalloc(&c->d_T, std::max<size_t>((size_t)2 * c->n * nconn, 1) * 8);
  if (e == cudaSuccess && c->n == 15) e = cudaMalloc(&c->d_Ta, std::max<size_t>((size_t)2 * c->n * nconn, 1) * 8);
  if (e != cudaSuccess) {
    rxn_connset_destroy(c);
    return fail(RXN_ERR_CUDA, "connection set upload failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return RXN_OK;
}

int rxn_connset_destroy(RxnConnSet *c) {
  if (!c) return RXN_OK;
  cudaSetDevice(c->device);
  cudaFree(c->d_row_ptr); cudaFree(c->d_col); cudaFree(c->d_ent); cudaFree(c->d_l2g); cudaFree(c->d_T);
  cudaFree(c->d_Ta);
  cudaFree(c->d_col_ptr); cudaFree(c->d_tgt_slot); cudaFree(c->d_tgt_ent); cudaFree(c->d_col_row);
  delete c;
  return RXN_OK;
}

int rxn_connset_structure(const RxnConnSet *c, int64_t *nnz_blocks, int32_t *row_ptr, int32_t *col) {
  if (!c) return fail(RXN_ERR_INVALID, "null connection set");
  if (nnz_blocks) *nnz_blocks = c->R.nnzb;
  if (row_ptr) memcpy(row_ptr, c->R.row_ptr.data(), c->R.row_ptr.size() * 4);
  if (col) memcpy(col, c->R.col.data(), c->R.col.size() * 4);
  return RXN_OK;
}

int rxn_connset_device_structure(const RxnConnSet *c, const int32_t **d_row_ptr, const int32_t **d_col) {
  if (!c) return fail(RXN_ERR_INVALID, "null connection set");
  if (d_row_ptr) *d_row_ptr = c->d_row_ptr;
  if (d_col) *d_col = c->d_col;
  return RXN_OK;
}

int rxn_connset_flux_coefs(RxnConnSet *c, const double *area, const double *velocity, const double *disp_over_dist,
                           const double *fraction_upwind, int use_upwinding) {
  Nvtx nvtx_("TFluxCoef");
  if (!c || !area || !velocity || !disp_over_dist || (!use_upwinding && !fraction_upwind)) return fail(RXN_ERR_INVALID, "bad argument");
  RxnState *s = c->s;
  const long long nc = c->R.nconn;
  const int n = c->n;
  if (nc == 0) { c->have_coefs = true; return RXN_OK; }
  CU(cudaSetDevice(s->t->device));
  void *tmp;
  int rc = ensure_scratch(s, 0, (size_t)nc * (n + 3) * 8, &tmp);
  if (rc != RXN_OK) return rc;
  double *d_area = (double *)tmp, *d_vel = d_area + nc, *d_fu = d_vel + nc, *d_disp = d_fu + nc;
  CU(cudaMemcpyAsync(d_area, area, (size_t)nc * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(d_vel, velocity, (size_t)nc * 8, cudaMemcpyHostToDevice, s->stream));
  if (fraction_upwind) CU(cudaMemcpyAsync(d_fu, fraction_upwind, (size_t)nc * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(d_disp, disp_over_dist, (size_t)nc * n * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaEventRecord(s->ev0, s->stream));
  k_flux_coefs<<<nblocks(nc, 256), 256, (size_t)256 * (n | 1) * 8, s->stream>>>(n, nc, d_area, d_vel, d_disp, d_fu, use_upwinding, c->d_T,
                                                                                c->d_T + (size_t)n * nc, c->d_Ta);
  ++g_launches;
  c->have_coefs = true;
  return check_launch(s, true);
}

static int flux_ready(RxnState *s, RxnConnSet *c, int field, const char *what) {
  if (!s || !c || c->s != s) return fail(RXN_ERR_INVALID, "bad argument (the connection set belongs to another state)");
  if (!c->have_coefs) return fail(RXN_ERR_INVALID, "rxn_connset_flux_coefs has not been called");
  if (!s->S.f[field]) return fail(RXN_ERR_INVALID, "%s is not materialised (rxn_state_materialize, then rxn_update_auxvars_batch)", what);
  return RXN_OK;
}

int rxn_flux_residual_batch_device(RxnState *s, RxnConnSet *c, double *d_res) {
  Nvtx nvtx_("RTResidualFlux");
  int rc = flux_ready(s, c, RXN_F_TOTAL, "total");
  if (rc != RXN_OK) return rc;
  if (!d_res) return fail(RXN_ERR_INVALID, "null output");
  if (c->R.nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  const int n = c->n;
  CU(cudaEventRecord(s->ev0, s->stream));
  k_flux_residual<<<nblocks(c->R.nlocal, 128), 128, (size_t)128 * (n | 1) * 8, s->stream>>>(
      n, c->R.nlocal, c->R.nconn, c->d_row_ptr, c->d_col, c->d_ent, c->d_l2g, s->S.f[RXN_F_TOTAL], s->ld, c->d_T,
      c->d_T + (size_t)n * c->R.nconn, d_res);
  ++g_launches;
  return check_launch(s, true);
}

int rxn_flux_jacobian_batch_device(RxnState *s, RxnConnSet *c, double *d_val) {
  Nvtx nvtx_("RTJacobianFlux");
  int rc = flux_ready(s, c, RXN_F_DTOTAL, "dtotal");
  if (rc != RXN_OK) return rc;
  if (!d_val) return fail(RXN_ERR_INVALID, "null output");
  if (c->R.nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  const int n = c->n;
  CU(cudaEventRecord(s->ev0, s->stream));
  const unsigned tiles = nblocks(c->R.nlocal, 32);
#define FLUX_JAC_T(N_, JC_)                                                                                                    \
  k_flux_jacobian_t<N_, JC_><<<tiles * (N_ / JC_), 32 * FLUX_NW, 0, s->stream>>>(c->R.nlocal, c->R.nconn, c->d_row_ptr, c->d_col, c->d_ent, \
                                                                          c->d_l2g, s->S.f[RXN_F_DTOTAL], s->ld, c->d_T,       \
                                                                          c->d_T + (size_t)n * c->R.nconn, d_val)
#ifndef FLUX_JC15
#define FLUX_JC15 5
#endif
  if (n == 15 && !s->flux_generic && s->flux_rows == 0) {
    // by block columns: every dtotal sector read once, blocks written from registers (rxn_flux.cuh)
    const long long quads = (c->R.nghosted + 3) / 4;
    k_flux_jacobian_cols<15><<<nblocks(quads, 4), 128, 0, s->stream>>>(c->R.nghosted, c->R.nconn, c->d_col_ptr, c->d_tgt_slot, c->d_tgt_ent,
                                                                    c->d_col_row, c->d_row_ptr, c->d_ent, s->S.f[RXN_F_DTOTAL], s->ld, c->d_Ta, d_val);
  } else if (n == 15 && !s->flux_generic) FLUX_JAC_T(15, FLUX_JC15);
  else if (n == 4 && !s->flux_generic) FLUX_JAC_T(4, 4);
  else if (n == 3 && !s->flux_generic) FLUX_JAC_T(3, 3);
  else {
    // column chunks of about 80 elements of a block per CTA (rxn_flux.cuh)
    int nchunk = (n * n + 79) / 80;
    const int jc = std::min((n + nchunk - 1) / nchunk, (int)FLUX_JC);
    nchunk = (n + jc - 1) / jc;
    const size_t smem = (size_t)32 * ((jc * n) | 1) * 8;
    k_flux_jacobian<<<dim3(tiles, nchunk), 128, smem, s->stream>>>(n, jc, c->R.nlocal, c->R.nconn, c->d_row_ptr, c->d_col, c->d_ent, c->d_l2g,
                                                                   s->S.f[RXN_F_DTOTAL], s->ld, c->d_T, c->d_T + (size_t)n * c->R.nconn, d_val);
  }
#undef FLUX_JAC_T
  ++g_launches;
  return check_launch(s, true);
}

int rxn_flux_residual_batch(RxnState *s, RxnConnSet *c, double *res_out) {
  Nvtx nvtx_("RTResidualFlux");
  if (!s || !c || !res_out) return fail(RXN_ERR_INVALID, "bad argument");
  if (c->R.nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  const size_t bytes = (size_t)c->R.nlocal * c->n * 8;
  void *d;
  int rc = ensure_scratch(s, 1, bytes, &d);
  if (rc != RXN_OK) return rc;
  if ((rc = rxn_flux_residual_batch_device(s, c, (double *)d)) != RXN_OK) return rc;
  CU(cudaMemcpyAsync(res_out, d, bytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}

int rxn_flux_jacobian_batch(RxnState *s, RxnConnSet *c, double *val_out) {
  Nvtx nvtx_("RTJacobianFlux");
  if (!s || !c || !val_out) return fail(RXN_ERR_INVALID, "bad argument");
  if (c->R.nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  const size_t bytes = (size_t)c->R.nnzb * c->n * c->n * 8;
  void *d;
  int rc = ensure_scratch(s, 2, bytes, &d);
  if (rc != RXN_OK) return rc;
  if ((rc = rxn_flux_jacobian_batch_device(s, c, (double *)d)) != RXN_OK) return rc;
  CU(cudaMemcpyAsync(val_out, d, bytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}

float rxn_last_kernel_ms(const RxnState *s) { return s ? s->last_ms : -1.f; }

// ------------------------------------------------------------------ boundary conditions / source-sinks (rxn_flux.h)
int rxn_couplerset_create(RxnState *s, int kind, int64_t nconn, const int32_t *id_dn, const int32_t *ghost_to_local, int64_t nlocal,
                          const uint8_t *active, RxnCouplerSet **out) {
  if (!s || !out || nconn < 0 || nlocal < 0 || (nconn > 0 && !id_dn) || (kind != RXN_COUPLER_BOUNDARY && kind != RXN_COUPLER_SRC_SINK))
    return fail(RXN_ERR_INVALID, "bad argument");
  *out = nullptr;
  if (s->t->h.nim > 0)
    return fail(RXN_ERR_UNSUPPORTED, "flux-side blocks are naqcomp x naqcomp: tables with immobile dofs (ncomp %d > naqcomp %d) are not handled by the coupler entry points", s->t->h.ncomp, s->t->h.naq);
  CU(cudaSetDevice(s->t->device));
  RxnCouplerSet *b = new RxnCouplerSet();
  b->s = s; b->kind = kind; b->n = s->t->h.naq; b->device = s->t->device;
  if (!coupler_rows_build(s->ncells, nlocal, nconn, id_dn, ghost_to_local, active, &b->R)) {
    const std::string e = b->R.err;
    delete b;
    return fail(RXN_ERR_INVALID, "coupler set: %s", e.c_str());
  }
  auto up = [&](int32_t **d, const std::vector<int32_t> &v) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, std::max<size_t>(v.size(), 1) * 4);
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    return e;
  };
  cudaError_t e = up(&b->d_row, b->R.row);
  if (e == cudaSuccess) e = up(&b->d_own, b->R.own);
  if (e == cudaSuccess) e = up(&b->d_row_ptr, b->R.row_ptr);
  if (e == cudaSuccess) e = up(&b->d_conn, b->R.conn);
  const size_t nb = std::max<size_t>((size_t)b->n * nconn, 1) * 8;
  if (e == cudaSuccess) e = cudaMalloc(&b->d_ext, nb);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_cx, nb);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_cc, nb);
  if (e != cudaSuccess) {
    rxn_couplerset_destroy(b);
    return fail(RXN_ERR_CUDA, "coupler set upload failed: %s", cudaGetErrorString(e));
  }
  *out = b;
  return RXN_OK;
}

int rxn_couplerset_destroy(RxnCouplerSet *b) {
  if (!b) return RXN_OK;
  cudaSetDevice(b->device);
  cudaFree(b->d_row); cudaFree(b->d_own); cudaFree(b->d_row_ptr); cudaFree(b->d_conn);
  cudaFree(b->d_ext); cudaFree(b->d_cx); cudaFree(b->d_cc);
  delete b;
  return RXN_OK;
}

int rxn_couplerset_bc_coefs(RxnCouplerSet *b, const double *area, const double *velocity, const double *disp_over_dist,
                            int use_upwinding) {
  Nvtx nvtx_("TFluxCoef (boundary)");
  if (!b || !area || !velocity || !disp_over_dist) return fail(RXN_ERR_INVALID, "bad argument");
  if (b->kind != RXN_COUPLER_BOUNDARY) return fail(RXN_ERR_INVALID, "not a boundary coupler set");
  RxnState *s = b->s;
  const long long nc = b->R.nconn;
  const int n = b->n;
  if (nc == 0) { b->have_coefs = true; return RXN_OK; }
  CU(cudaSetDevice(s->t->device));
  void *tmp;
  int rc = ensure_scratch(s, 0, (size_t)nc * (n + 3) * 8, &tmp);
  if (rc != RXN_OK) return rc;
  double *d_area = (double *)tmp, *d_vel = d_area + nc, *d_fu = d_vel + nc, *d_disp = d_fu + nc;
  CU(cudaMemcpyAsync(d_area, area, (size_t)nc * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(d_vel, velocity, (size_t)nc * 8, cudaMemcpyHostToDevice, s->stream));
  k_fill<<<nblocks(nc, 256), 256, 0, s->stream>>>(d_fu, nc, 0.5);                // fraction upwind of a boundary face (:2372)
  CU(cudaMemcpyAsync(d_disp, disp_over_dist, (size_t)nc * n * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaEventRecord(s->ev0, s->stream));
  k_flux_coefs<<<nblocks(nc, 256), 256, (size_t)256 * (n | 1) * 8, s->stream>>>(n, nc, d_area, d_vel, d_disp, d_fu, use_upwinding, b->d_cx,
                                                                                b->d_cc);
  g_launches += 2;
  b->have_coefs = true;
  return check_launch(s, true);
}

int rxn_couplerset_ss_coefs(RxnCouplerSet *b, const double *qsrc, const int32_t *tran_src_sink_type) {
  Nvtx nvtx_("TSrcSinkCoef");
  if (!b || !qsrc || !tran_src_sink_type) return fail(RXN_ERR_INVALID, "bad argument");
  if (b->kind != RXN_COUPLER_SRC_SINK) return fail(RXN_ERR_INVALID, "not a source/sink coupler set");
  RxnState *s = b->s;
  const long long nc = b->R.nconn;
  if (nc == 0) { b->have_coefs = true; return RXN_OK; }
  CU(cudaSetDevice(s->t->device));
  void *tmp;
  int rc = ensure_scratch(s, 0, (size_t)nc * 12, &tmp);
  if (rc != RXN_OK) return rc;
  double *d_q = (double *)tmp;
  int32_t *d_t = (int32_t *)(d_q + nc);
  CU(cudaMemcpyAsync(d_q, qsrc, (size_t)nc * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(d_t, tran_src_sink_type, (size_t)nc * 4, cudaMemcpyHostToDevice, s->stream));
  CU(cudaEventRecord(s->ev0, s->stream));
  k_coupler_ss_coefs<<<nblocks(nc, 256), 256, 0, s->stream>>>(b->n, nc, d_q, d_t, b->d_cx, b->d_cc);
  ++g_launches;
  b->have_coefs = true;
  return check_launch(s, true);
}

int rxn_couplerset_set_totals(RxnCouplerSet *b, const double *total) {
  if (!b || !total) return fail(RXN_ERR_INVALID, "bad argument");
  RxnState *s = b->s;
  const long long nc = b->R.nconn;
  if (nc == 0) { b->have_totals = true; return RXN_OK; }
  CU(cudaSetDevice(s->t->device));
  void *tmp;
  int rc = ensure_scratch(s, 0, (size_t)nc * b->n * 8, &tmp);
  if (rc != RXN_OK) return rc;
  CU(cudaMemcpyAsync(tmp, total, (size_t)nc * b->n * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaEventRecord(s->ev0, s->stream));
  k_coupler_totals<<<nblocks(nc * b->n, 256), 256, 0, s->stream>>>(b->n, nc, (const double *)tmp, 0, b->d_ext);
  ++g_launches;
  b->have_totals = true;
  return check_launch(s, true);
}

int rxn_couplerset_totals_from_state(RxnCouplerSet *b, const RxnState *bc) {
  if (!b || !bc) return fail(RXN_ERR_INVALID, "bad argument");
  RxnState *s = b->s;
  const long long nc = b->R.nconn;
  if (bc->t->device != s->t->device) return fail(RXN_ERR_INVALID, "the boundary state lives on another device");
  if (bc->t->h.naq != b->n) return fail(RXN_ERR_INVALID, "the boundary state has %d components, the coupler set %d", bc->t->h.naq, b->n);
  if (bc->ncells < nc) return fail(RXN_ERR_INVALID, "the boundary state has %lld cells for %lld connections", (long long)bc->ncells, nc);
  if (nc == 0) { b->have_totals = true; return RXN_OK; }
  CU(cudaSetDevice(s->t->device));
  CU(cudaStreamSynchronize(bc->stream));                         // the boundary state's update runs on its own stream
  CU(cudaEventRecord(s->ev0, s->stream));
  k_coupler_totals<<<nblocks(nc * b->n, 256), 256, 0, s->stream>>>(b->n, nc, bc->S.f[RXN_F_TOTAL], bc->ld, b->d_ext);
  ++g_launches;
  b->have_totals = true;
  return check_launch(s, true);
}

static int coupler_ready(RxnState *s, RxnCouplerSet *b, int field, const char *what) {
  if (!s || !b || b->s != s) return fail(RXN_ERR_INVALID, "bad argument (the coupler set belongs to another state)");
  if (!b->have_coefs) return fail(RXN_ERR_INVALID, "rxn_couplerset_bc_coefs / rxn_couplerset_ss_coefs has not been called");
  if (!s->S.f[field]) return fail(RXN_ERR_INVALID, "%s is not materialised (rxn_state_materialize, then rxn_update_auxvars_batch)", what);
  return RXN_OK;
}

int rxn_coupler_residual_batch_device(RxnState *s, RxnCouplerSet *b, double *d_res, double *d_flux_out) {
  Nvtx nvtx_(b && b->kind == RXN_COUPLER_SRC_SINK ? "RTResidual (source/sink)" : "RTResidualFlux (boundary)");
  int rc = coupler_ready(s, b, RXN_F_TOTAL, "total");
  if (rc != RXN_OK) return rc;
  if (!b->have_totals) return fail(RXN_ERR_INVALID, "the external totals have not been set (rxn_couplerset_set_totals / _totals_from_state)");
  if (!d_res) return fail(RXN_ERR_INVALID, "null residual");
  if (b->R.nrows == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  CU(cudaEventRecord(s->ev0, s->stream));
  k_coupler_residual<<<nblocks(b->R.nrows * b->n, 128), 128, 0, s->stream>>>(
      b->n, b->R.nrows, b->R.nconn, b->kind == RXN_COUPLER_BOUNDARY ? -1.0 : 1.0, b->d_row, b->d_own, b->d_row_ptr, b->d_conn,
      s->S.f[RXN_F_TOTAL], s->ld, b->d_ext, b->d_cx, b->d_cc, d_res, d_flux_out);
  ++g_launches;
  return check_launch(s, true);
}

int rxn_coupler_jacobian_batch_device(RxnState *s, RxnConnSet *c, RxnCouplerSet *b, double *d_val) {
  Nvtx nvtx_(b && b->kind == RXN_COUPLER_SRC_SINK ? "RTJacobianSS" : "RTJacobianFluxBC");
  int rc = coupler_ready(s, b, RXN_F_DTOTAL, "dtotal");
  if (rc != RXN_OK) return rc;
  if (c && (c->s != s || c->R.nlocal != b->R.nlocal)) return fail(RXN_ERR_INVALID, "the connection set and the coupler set describe different grids");
  if (!d_val) return fail(RXN_ERR_INVALID, "null matrix values");
  if (b->R.nrows == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  CU(cudaEventRecord(s->ev0, s->stream));
  k_coupler_jacobian<<<nblocks(b->R.nrows * b->n * b->n, 128), 128, 0, s->stream>>>(
      b->n, b->R.nrows, b->R.nconn, b->kind == RXN_COUPLER_BOUNDARY ? -1.0 : 1.0, b->d_row, b->d_own, b->d_row_ptr, b->d_conn,
      c ? c->d_row_ptr : nullptr, s->S.f[RXN_F_DTOTAL], s->ld, b->d_cc, d_val);
  ++g_launches;
  return check_launch(s, true);
}

int rxn_coupler_residual_batch(RxnState *s, RxnCouplerSet *b, double *res_inout, double *flux_out) {
  if (!s || !b || !res_inout) return fail(RXN_ERR_INVALID, "bad argument");
  if (b->R.nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  const size_t bytes = (size_t)b->R.nlocal * b->n * 8, fbytes = (size_t)b->R.nconn * b->n * 8;
  void *d, *df = nullptr;
  int rc = ensure_scratch(s, 1, bytes, &d);
  if (rc != RXN_OK) return rc;
  if (flux_out && fbytes) {
    if ((rc = ensure_scratch(s, 2, fbytes, &df)) != RXN_OK) return rc;
    CU(cudaMemcpyAsync(df, flux_out, fbytes, cudaMemcpyHostToDevice, s->stream));   // rows of skipped connections keep the caller's values
  }
  CU(cudaMemcpyAsync(d, res_inout, bytes, cudaMemcpyHostToDevice, s->stream));
  if ((rc = rxn_coupler_residual_batch_device(s, b, (double *)d, (double *)df)) != RXN_OK) return rc;
  CU(cudaMemcpyAsync(res_inout, d, bytes, cudaMemcpyDeviceToHost, s->stream));
  if (df) CU(cudaMemcpyAsync(flux_out, df, fbytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}

int rxn_coupler_jacobian_batch(RxnState *s, RxnConnSet *c, RxnCouplerSet *b, double *val_inout) {
  if (!s || !b || !val_inout) return fail(RXN_ERR_INVALID, "bad argument");
  if (b->R.nlocal == 0) return RXN_OK;
  CU(cudaSetDevice(s->t->device));
  const size_t bytes = (size_t)(c ? c->R.nnzb : b->R.nlocal) * b->n * b->n * 8;
  void *d;
  int rc = ensure_scratch(s, 0, bytes, &d);
  if (rc != RXN_OK) return rc;
  CU(cudaMemcpyAsync(d, val_inout, bytes, cudaMemcpyHostToDevice, s->stream));
  if ((rc = rxn_coupler_jacobian_batch_device(s, c, b, (double *)d)) != RXN_OK) return rc;
  CU(cudaMemcpyAsync(val_inout, d, bytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}

int rxn_state_device_ptr(RxnState *s, int field, double **d_ptr, int64_t *ld) {
  if (!s || field < 0 || field >= RXN_F_COUNT || !d_ptr) return fail(RXN_ERR_INVALID, "bad argument");
  *d_ptr = s->S.f[field];
  if (ld) *ld = s->ld;
  return RXN_OK;
}
int rxn_device_alloc(RxnState *s, int64_t bytes, void **d_ptr) {
  if (!s || !d_ptr || bytes <= 0) return fail(RXN_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(s->t->device));
  CU(cudaMalloc(d_ptr, (size_t)bytes));
  return RXN_OK;
}
int rxn_device_free(RxnState *s, void *d_ptr) {
  if (!s) return fail(RXN_ERR_INVALID, "null state");
  CU(cudaSetDevice(s->t->device));
  CU(cudaFree(d_ptr));
  return RXN_OK;
}
int rxn_device_copy(RxnState *s, void *dst, const void *src, int64_t bytes, int kind) {
  if (!s || !dst || !src || bytes < 0) return fail(RXN_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(s->t->device));
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  CU(cudaMemcpyAsync(dst, src, (size_t)bytes, k, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}
int rxn_device_sync(RxnState *s) {
  if (!s) return fail(RXN_ERR_INVALID, "null state");
  CU(cudaSetDevice(s->t->device));
  CU(cudaStreamSynchronize(s->stream));
  return RXN_OK;
}
int rxn_host_alloc(int64_t bytes, void **h_ptr) {
  if (!h_ptr || bytes <= 0) return fail(RXN_ERR_INVALID, "bad argument");
  CU(cudaMallocHost(h_ptr, (size_t)bytes));
  return RXN_OK;
}
int rxn_host_free(void *h_ptr) {
  CU(cudaFreeHost(h_ptr));
  return RXN_OK;
}

}  // extern "C"
