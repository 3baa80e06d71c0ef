// rxn_small.cu — the register RReact kernel for small chemistries (design: rxn_small.h, device code: rxn_small_dev.cuh).
#include <algorithm>

#include "rxn_small.cuh"
#include "rxn_small_dev.cuh"

namespace rxn {
namespace small {

// One thread = one cell; the tables are a kernel parameter (constant bank), nothing is staged in shared memory.
__global__ void __launch_bounds__(128)
k_react_small(const __grid_constant__ SmallTab T, DevState S, double *tran_xx, const int32_t *__restrict__ l2g, long long nlocal, double dt,
              int dt_mode, int32_t *iters, int32_t *flags, long long cell0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const long long cell = l2g ? l2g[i] : i + cell0;                // cell0: first cell of this chunk of the batch
  if (S.active && !S.active[cell]) {                              // imat <= 0 (reactive_transport.F90:1699)
    if (iters) iters[i] = 0;
    if (flags) flags[i] = RXN_FLAG_INACTIVE;
    return;
  }
  small_react_cell<SMALL_N>(T, S, i, cell, tran_xx, dt, dt_mode, iters, flags);
}

}  // namespace small

int small_launch_react(const SmallPlan &p, const DevState &S, double *tran_xx, const int32_t *l2g, long long nlocal, double dt, int dt_mode,
                       int32_t *iters, int32_t *flags, cudaStream_t stream, long long cell0) {
  if (!p.usable) return RXN_ERR_UNSUPPORTED;
  if (nlocal <= 0) return RXN_OK;
  const unsigned grid = (unsigned)((nlocal + 127) / 128);
  small::k_react_small<<<grid, 128, 0, stream>>>(p.st, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, cell0);
  return RXN_OK;
}

}  // namespace rxn
