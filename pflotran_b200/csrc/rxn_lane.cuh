// rxn_lane.cuh — host interface of the resident-lane RReact kernel (plan: rxn_lane.h,
// device code: rxn_lane_dev.cuh, one translation unit per shape: rxn_lane_variant.cu).
#pragma once
#include <cuda_runtime.h>

#include "rxn_lane.h"

namespace rxn {

// shapes compiled into the library: X(N, CPB, G) = matrix dimension, resident cells per CTA, lanes per cell;
// per N in order of preference - the first shape whose per-cell state fits in shared memory is used
// (measured on B200, 300A chemistry: G = 2 > G = 1 > G = 4; the 22-primary / 164-complex ascem chemistry holds only 16 cells
// per SM, where more lanes per cell win (for N = 8, 192 cells per SM, they lose: G = 1 / 2 / 4 = 172 / 161 / 117 M cell-updates/s on the
// scco2_brine chemistry, profiles/r02_ag_*.json): G = 8 > 4 > 2 and 16 no better, profiles/r02_ab_ascem_g*.json, r02_ad_ascem_*.json; keep in sync with LANE_SHAPES in the Makefile)
#define RXN_LANE_SHAPES(X) \
  X(4, 448, 1) X(4, 256, 1) X(4, 128, 1) \
  X(8, 192, 1) X(8, 128, 1) X(8, 64, 1) \
  X(12, 96, 2) X(12, 64, 2) X(12, 64, 1) X(12, 64, 4) \
  X(15, 64, 2) X(15, 60, 2) X(15, 48, 2) X(15, 64, 4) X(15, 64, 1) \
  X(16, 64, 2) X(16, 48, 2) \
  X(24, 32, 2) X(24, 28, 2) X(24, 24, 2) X(24, 20, 8) X(24, 16, 8) X(24, 16, 2) X(24, 28, 4) X(24, 28, 1)

// tensor-memory kernel shapes (rxn_tm_dev.cuh): X(N, QUADS, G) = matrix dimension (<= 15), 32-cell quads per CTA, member warps
// per cell; per N the first shape whose vectors fit in shared memory is used (keep in sync with TM_SHAPES in the Makefile)
#define RXN_TM_SHAPES(X) \
  X(15, 4, 2) X(15, 3, 2) X(15, 2, 2) X(15, 4, 4) X(15, 3, 4) X(15, 2, 4) X(15, 4, 3) X(15, 3, 3) X(15, 4, 1)

struct LaneKernel {
  LanePlan plan_tm;      // tensor-memory kernel (preferred when usable)
  int G_tm = 0, quads_tm = 0;
  double *d_blob_tm = nullptr;
  LanePlan plan;
  int G = 1;             // lanes per cell of the selected shape
  LanePlan plan_gi;      // global-implicit residual/Jacobian kernel: same streams, activity coefficients from the state
  double *d_blob_gi = nullptr;
  int G_gi = 1;
  double *d_blob = nullptr;
  int sm_count = 0;
  std::vector<double> mr_rate, mr_frac;   // host copies for the per-launch K1 sums
  std::vector<int> mr_nrate;
  int mr_ld = 0;
};

template <int N, int CPB, int G>
int lane_launch_variant(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob, const double *blob,
                        const DevState &S, double *tran_xx, const int32_t *l2g, long long nlocal, double dt, int dt_mode,
                        int32_t *iters, int32_t *flags, unsigned long long *counter, long long cell0, cudaStream_t stream);

template <int N, int QUADS, int G>
int tm_launch_variant(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob, const double *blob,
                      const DevState &S, double *tran_xx, const int32_t *l2g, long long nlocal, double dt, int dt_mode, int32_t *iters,
                      int32_t *flags, unsigned long long *counter, long long cell0, cudaStream_t stream);

template <int N, int QUADS, int G>
int tm_launch_gi_variant(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob, const double *blob,
                         const DevState &S, const int32_t *l2g, long long nlocal, const GiArgs &a, cudaStream_t stream);

template <int N, int CPB, int G>
int lane_launch_gi_variant(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob, const double *blob,
                           const DevState &S, const int32_t *l2g, long long nlocal, double dt, double *res_out, double *jac_out,
                           cudaStream_t stream);

int lane_kernel_build(const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi, int device, LaneKernel *k);
void lane_kernel_free(LaneKernel *k);
int lane_launch_react(LaneKernel &k, const DevTab &h, const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                      long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter,
                      cudaStream_t stream, long long cell0 = 0);

// global-implicit pass on the tensor-memory layout: usable when the tensor-memory plan is, except for an activity update with
// activity coefficients switched off in the tables (one class per species: nothing to compute them from)
bool tm_gi_usable(const LaneKernel &k, int update_act);
int tm_launch_gi(LaneKernel &k, const DevTab &h, const double *blob, const DevState &S, const int32_t *l2g, long long nlocal,
                 const GiArgs &a, cudaStream_t stream);

int lane_launch_gi(LaneKernel &k, const DevTab &h, const double *blob, const DevState &S, const int32_t *l2g, long long nlocal, double dt,
                   double *res_out, double *jac_out, cudaStream_t stream);

}  // namespace rxn
