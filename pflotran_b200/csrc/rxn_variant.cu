// rxn_variant.cu — kernel definitions for one compile-time bound RXN_N on naqcomp.
// Compiled once per RXN_N in {4, 8, 16, 24} (see Makefile).
#ifndef RXN_N
#error "compile with -DRXN_N=<bound>"
#endif
#include "rxn_kernels.cuh"
#include "rxn_device.cuh"

namespace rxn {

extern __shared__ double rxn_smem[];

// stage the table blob ([ndbl doubles][nint int32]) into shared memory
__device__ __forceinline__ Tab stage_tables(const DevTab &tab, const double *__restrict__ blob) {
  const int words = tab.ndbl + (tab.nint + 1) / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) rxn_smem[w] = blob[w];
  __syncthreads();
  Tab T;
  T.d = rxn_smem;
  T.i = reinterpret_cast<const int *>(rxn_smem + tab.ndbl);
  T.h = &tab;
  return T;
}

template <int N>
__global__ void __launch_bounds__(128)
k_react(const __grid_constant__ DevTab tab, const double *__restrict__ blob, DevState S, double *tran_xx,
        const int *__restrict__ l2g, long long nlocal, double dt, int dt_mode, int *iters, int *flags) {
  Tab T = stage_tables(tab, blob);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += (long long)gridDim.x * blockDim.x)
    cell_react<N>(T, S, i, tran_xx, l2g, dt, dt_mode, iters, flags);
}

template <int N>
__global__ void __launch_bounds__(128)
k_update_auxvars(const __grid_constant__ DevTab tab, const double *__restrict__ blob, DevState S,
                 const double *__restrict__ xx_loc, int update_act_coefs) {
  Tab T = stage_tables(tab, blob);
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < S.ncells) cell_update_auxvars<N>(T, S, cell, xx_loc, update_act_coefs);
}

template <int N>
__global__ void __launch_bounds__(128)
k_fixed_accum(const __grid_constant__ DevTab tab, const double *__restrict__ blob, DevState S,
              const double *__restrict__ xx, const int *__restrict__ l2g, long long nlocal, double *accum_out) {
  Tab T = stage_tables(tab, blob);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nlocal) cell_fixed_accum<N>(T, S, i, xx, l2g, accum_out);
}

template <int N>
__global__ void __launch_bounds__(128)
k_residual_jacobian(const __grid_constant__ DevTab tab, const double *__restrict__ blob, DevState S,
                    const int *__restrict__ l2g, long long nlocal, double dt, double *res_out, double *jac_out) {
  Tab T = stage_tables(tab, blob);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nlocal) cell_residual_jacobian<N>(T, S, i, l2g, dt, res_out, jac_out);
}

template <int N>
__global__ void __launch_bounds__(128)
k_update_kinetic_state(const __grid_constant__ DevTab tab, const double *__restrict__ blob, DevState S, double dt, int skip_mr) {
  Tab T = stage_tables(tab, blob);
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < S.ncells) cell_update_kinetic_state<N>(T, S, cell, dt, skip_mr != 0);
}

// ReactionEquilibrateConstraint applied cell by cell (reaction.F90:1308-2046; condition_control.F90:725-741)
template <int N>
__global__ void __launch_bounds__(128)
k_equilibrate(const __grid_constant__ DevTab tab, const double *__restrict__ blob, DevState S, const int *__restrict__ ctype,
              const double *__restrict__ conc, long long conc_stride, const int *__restrict__ cid, const double *__restrict__ guess,
              int use_prev, int init_molal, const int *__restrict__ l2g, long long nlocal, double *basis_out, int *iters, int *status) {
  Tab T = stage_tables(tab, blob);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const long long cell = l2g ? l2g[i] : i;
  if (S.active && !S.active[cell]) {
    if (iters) iters[i] = 0;
    if (status) status[i] = RXN_EQ_OK;
    return;
  }
  int nit = 0;
  const int rc = cell_equilibrate<N>(T, S, cell, ctype, conc + i * conc_stride, cid, guess, use_prev, init_molal,
                                     basis_out ? basis_out + i * tab.naq : nullptr, &nit);
  if (iters) iters[i] = nit;
  if (status) status[i] = rc;
}

template <class K>
static void set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <> void run_react<RXN_N>(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, double *tran_xx,
                                  const int *l2g, long long nlocal, double dt, int dt_mode, int *iters, int *flags) {
  set_smem(k_react<RXN_N>, L.smem);
  k_react<RXN_N><<<L.grid, L.block, L.smem, L.stream>>>(tab, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags);
}
template <> void run_update_auxvars<RXN_N>(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S,
                                           const double *xx_loc, int update_act_coefs) {
  set_smem(k_update_auxvars<RXN_N>, L.smem);
  k_update_auxvars<RXN_N><<<L.grid, L.block, L.smem, L.stream>>>(tab, blob, S, xx_loc, update_act_coefs);
}
template <> void run_fixed_accum<RXN_N>(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, const double *xx,
                                        const int *l2g, long long nlocal, double *accum_out) {
  set_smem(k_fixed_accum<RXN_N>, L.smem);
  k_fixed_accum<RXN_N><<<L.grid, L.block, L.smem, L.stream>>>(tab, blob, S, xx, l2g, nlocal, accum_out);
}
template <> void run_residual_jacobian<RXN_N>(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S,
                                              const int *l2g, long long nlocal, double dt, double *res_out, double *jac_out) {
  set_smem(k_residual_jacobian<RXN_N>, L.smem);
  k_residual_jacobian<RXN_N><<<L.grid, L.block, L.smem, L.stream>>>(tab, blob, S, l2g, nlocal, dt, res_out, jac_out);
}
template <> void run_update_kinetic_state<RXN_N>(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, double dt, int skip_mr) {
  set_smem(k_update_kinetic_state<RXN_N>, L.smem);
  k_update_kinetic_state<RXN_N><<<L.grid, L.block, L.smem, L.stream>>>(tab, blob, S, dt, skip_mr);
}

template <> void run_equilibrate<RXN_N>(LaunchCfg L, const DevTab &tab, const double *blob, const DevState &S, const int *ctype,
                                        const double *conc, long long conc_stride, const int *cid, const double *guess, int use_prev,
                                        int init_molal, const int *l2g, long long nlocal, double *basis_out, int *iters, int *status) {
  set_smem(k_equilibrate<RXN_N>, L.smem);
  k_equilibrate<RXN_N><<<L.grid, L.block, L.smem, L.stream>>>(tab, blob, S, ctype, conc, conc_stride, cid, guess, use_prev, init_molal,
                                                           l2g, nlocal, basis_out, iters, status);
}

}  // namespace rxn
