// rxn_lane_variant.cu — one instantiation (LANE_N x LANE_CPB) of the resident-lane RReact kernel
// (compiled once per shape, see Makefile; device code in rxn_lane_dev.cuh, design in rxn_lane.h).
#if !defined(LANE_N) || !defined(LANE_CPB)
#error "compile with -DLANE_N=<matrix dimension> -DLANE_CPB=<resident cells per CTA>"
#endif
#include <algorithm>

#include "rxn_lane.cuh"
#include "rxn_lane_dev.cuh"

namespace rxn {
namespace lane {

// Persistent lanes.  Every lane of a CTA owns one column of the shared-memory arrays; a lane without a
// cell takes the next item from the global counter (one atomicAdd per warp and round), loads it, and
// from then on makes one trip through the Newton loop per iteration of the outer loop together with
// the other lanes of its warp, whatever Newton iteration each of them is in.
template <int N, int CPB>
__global__ void __launch_bounds__(((CPB + 31) / 32) * 32, 1)
k_react_lane(const __grid_constant__ LaneTab lt, const __grid_constant__ DevTab h, const double *__restrict__ pblob,
             const double *__restrict__ blob, DevState S, double *tran_xx, const int32_t *__restrict__ l2g, long long nlocal,
             double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter) {
  const int words = lt.blob_dbl + lt.blob_int / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) tsm[w] = pblob[w];
  __syncthreads();
  const int t = threadIdx.x, ln = t & 31;
  const double *bd = blob;
  const int *bi = reinterpret_cast<const int *>(blob + h.ndbl);
  Lane<N> c;
  lane_bind<N, CPB>(lt, c, t);
  bool active = false, exhausted = t >= lt.cells;
  int pending = 0;                                             // exit status waiting for its closing pass
  const double inv_dt = 1.0 / dt;
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (;;) {                                                 // hand out work to the idle lanes of this warp
      const bool want = !active && !exhausted;
      const unsigned wm = __ballot_sync(0xffffffffu, want);
      if (wm == 0u) break;
      const int leader = __ffs(wm) - 1;
      unsigned long long base = 0;
      if (ln == leader) base = atomicAdd(counter, (unsigned long long)__popc(wm));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (want) {
        const long long i = (long long)base + __popc(wm & ((1u << ln) - 1u));
        if (i >= nlocal) {
          exhausted = true;
        } else {
          const long long cell = l2g ? l2g[i] : i;
          if (S.active && !S.active[cell]) {                   // imat <= 0 (reactive_transport.F90:1699)
            if (iters) iters[i] = 0;
            if (flags) flags[i] = RXN_FLAG_INACTIVE;
          } else {
            lane_load<N, CPB>(lt, c, S, bd, bi, h, i, cell, tran_xx, dt);
            active = true;
            pending = 0;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;
    if (active) {
      bool recompute;
      const int st = lane_trip<N, CPB>(lt, c, S, dt, inv_dt, dt_mode, pending != 0, recompute);
      if (pending != 0) {
        lane_finish<N, CPB>(lt, c, S, h, tran_xx, iters, flags, pending);
        active = false;
      } else if (st != 0) {
        if (recompute) {
          pending = st;
        } else {
          lane_finish<N, CPB>(lt, c, S, h, tran_xx, iters, flags, st);
          active = false;
        }
      }
    }
  }
}

}  // namespace lane

template <>
int lane_launch_variant<LANE_N, LANE_CPB>(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob,
                                          const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g, long long nlocal,
                                          double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter,
                                          cudaStream_t stream) {
  auto kern = lane::k_react_lane<LANE_N, LANE_CPB>;
  constexpr int threads = ((LANE_CPB + 31) / 32) * 32;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return RXN_ERR_CUDA;
  int bps = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem_bytes) != cudaSuccess || bps < 1) bps = 1;
  const long long want = (nlocal + LANE_CPB - 1) / LANE_CPB;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * bps));
  kern<<<grid, threads, smem_bytes, stream>>>(lt, h, pblob, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, counter);
  return RXN_OK;
}

}  // namespace rxn
