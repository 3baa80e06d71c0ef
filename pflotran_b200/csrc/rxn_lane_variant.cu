// rxn_lane_variant.cu — one instantiation (LANE_N x LANE_CPB x LANE_G) of the resident-lane RReact kernel
// (compiled once per shape, see Makefile; device code in rxn_lane_dev.cuh, design in rxn_lane.h).
#if !defined(LANE_N) || !defined(LANE_CPB) || !defined(LANE_G)
#error "compile with -DLANE_N=<matrix dimension> -DLANE_CPB=<resident cells per CTA> -DLANE_G=<lanes per cell>"
#endif
#include <algorithm>

#include "rxn_lane.cuh"
#include "rxn_lane_dev.cuh"

namespace rxn {
namespace lane {

// Persistent lane groups.  Every group of G lanes owns one column of the shared-memory arrays; a group
// without a cell takes the next item from the global counter (one atomicAdd per warp and round), loads
// it, and from then on makes one trip through the Newton loop per iteration of the outer loop together
// with the other groups of its warp, whatever Newton iteration each of them is in.
template <int N, int CPB, int G>
__global__ void __launch_bounds__(((CPB * G + 31) / 32) * 32, 1)
k_react_lane(const __grid_constant__ LaneTab lt, const __grid_constant__ DevTab h, const double *__restrict__ pblob,
             const double *__restrict__ blob, DevState S, double *tran_xx, const int32_t *__restrict__ l2g, long long nlocal,
             double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter) {
  const int words = lt.blob_dbl + lt.blob_int / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) tsm[w] = pblob[w];
  __syncthreads();
  const int t = threadIdx.x, ln = t & 31;
  const int s = t / G, l = t % G;
  const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (ln & ~(G - 1)));
  const unsigned leaders = G == 1 ? 0xffffffffu : G == 2 ? 0x55555555u : G == 4 ? 0x11111111u : G == 8 ? 0x01010101u : 0x00010001u;
  const double *bd = blob;
  const int *bi = reinterpret_cast<const int *>(blob + h.ndbl);
  Lane<N, G> c;
  lane_bind<N, CPB, G>(lt, c, s, l, gm);
  bool active = false, exhausted = s >= lt.cells;              // uniform over a group
  int pending = 0;                                             // exit status waiting for its closing pass
  long long next_item = -1;                                    // item reserved (and prefetched into L2) for this group
  const double inv_dt = 1.0 / dt;
  // one-cell-ahead reservation + L2 prefetch: only when every group gets several cells anyway (small batches
  // would lose balance: a reserved item cannot be taken over by an idle group)
#ifndef LANE_NO_PREFETCH
  const bool kPrefetch = nlocal >= 4LL * gridDim.x * CPB;
#else
  const bool kPrefetch = false;
#endif
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (;;) {                                                 // hand out work to the idle groups of this warp
      // a group without a cell first takes its reserved item; every idle or reservation-less group then draws from the counter
      const bool idle = !active && !exhausted;
      long long i = -1;
      if (idle && next_item >= 0) { i = next_item; next_item = -1; }
      const bool want = !exhausted && ((idle && i < 0) || (kPrefetch && next_item == -1 && (active || i >= 0)));
      const unsigned wm = __ballot_sync(0xffffffffu, want) & leaders;
      if (wm != 0u) {
        const int leader = __ffs(wm) - 1;
        unsigned long long base = 0;
        if (ln == leader) base = atomicAdd(counter, (unsigned long long)__popc(wm));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (want) {
          const long long got = (long long)base + __popc(wm & ((1u << (ln & ~(G - 1))) - 1u));
          if (got >= nlocal) {
            if (idle && i < 0) exhausted = true;               // nothing left for a group that has no cell
            else next_item = -2;                                // no further reservation attempts
          } else if (idle && i < 0) {
            i = got;
          } else {
            next_item = got;
            const long long pc = l2g ? l2g[got] : got;
            if (!(S.active && !S.active[pc])) lane_prefetch<N, CPB, G>(lt, l, S, h, got, pc, tran_xx);
          }
        }
      }
      if (idle && i >= 0) {
        const long long cell = l2g ? l2g[i] : i;
        if (S.active && !S.active[cell]) {                     // imat <= 0 (reactive_transport.F90:1699)
          if (l == 0) {
            if (iters) iters[i] = 0;
            if (flags) flags[i] = RXN_FLAG_INACTIVE;
          }
        } else {
          lane_load<N, CPB, G>(lt, c, S, bd, bi, h, i, cell, tran_xx, dt);
          active = true;
          pending = 0;
        }
      }
      // another round while some group is still without a cell (inactive cell drawn, or reservation just consumed)
      const bool again = !active && !exhausted;
      if (!__any_sync(0xffffffffu, again || (kPrefetch && !exhausted && next_item == -1 && active))) break;
    }
    if (!__any_sync(0xffffffffu, active)) break;
    if (active) {
      bool recompute;
      const int st = lane_trip<N, CPB, G>(lt, c, S, dt, inv_dt, dt_mode, pending != 0, recompute);
      if (pending != 0) {
        lane_finish<N, CPB, G>(lt, c, S, h, tran_xx, iters, flags, pending);
        active = false;
      } else if (st != 0) {
        if (recompute) {
          pending = st;
        } else {
          lane_finish<N, CPB, G>(lt, c, S, h, tran_xx, iters, flags, st);
          active = false;
        }
      }
    }
  }
}

}  // namespace lane

template <>
int lane_launch_variant<LANE_N, LANE_CPB, LANE_G>(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob,
                                                  const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                                                  long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags,
                                                  unsigned long long *counter, cudaStream_t stream) {
  auto kern = lane::k_react_lane<LANE_N, LANE_CPB, LANE_G>;
  constexpr int threads = ((LANE_CPB * LANE_G + 31) / 32) * 32;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return RXN_ERR_CUDA;
  int bps = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem_bytes) != cudaSuccess || bps < 1) bps = 1;
  const long long want = (nlocal + LANE_CPB - 1) / LANE_CPB;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * bps));
  kern<<<grid, threads, smem_bytes, stream>>>(lt, h, pblob, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, counter);
  return RXN_OK;
}

}  // namespace rxn
