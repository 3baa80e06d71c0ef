// rxn_kernels.cuh — batched kernels: one thread per cell, tables staged in shared memory.
// Each kernel replaces one per-rank cell loop of src/pflotran/reactive_transport.F90 (the
// per-cell bodies, with citations, are in rxn_device.cuh).  Instantiated per naq bound N in
// rxn_variant.cu (one translation unit per N so the build parallelises).
#pragma once
#include <cuda_runtime.h>
#include "rxn_tab.h"

namespace rxn {

struct LaunchCfg {
  unsigned grid;
  int block;
  size_t smem;
  cudaStream_t stream;
};

// host launchers (defined in rxn_variant.cu for N = 4, 8, 16, 24)
template <int N> void run_react(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, double *tran_xx,
                                const int *l2g, long long nlocal, double dt, int dt_mode, int *iters, int *flags);
template <int N> void run_update_auxvars(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S,
                                         const double *xx_loc, int update_act_coefs);
template <int N> void run_fixed_accum(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, const double *xx,
                                      const int *l2g, long long nlocal, double *accum_out);
template <int N> void run_residual_jacobian(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S,
                                            const int *l2g, long long nlocal, double dt, double *res_out, double *jac_out);
template <int N> void run_update_kinetic_state(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, double dt, int skip_mr);

template <int N> void run_equilibrate(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, const int *ctype,
                                      const double *conc, long long conc_stride, const int *cid, const double *guess, int use_prev,
                                      int init_molal, const int *l2g, long long nlocal, double *basis_out, int *iters, int *status);

}  // namespace rxn
