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
             double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter, long long cell0) {
  const int words = lt.blob_dbl + lt.blob_int / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) tsm[w] = pblob[w];
  __syncthreads();
  const int t = threadIdx.x, ln = t & 31;
#ifndef LANE_ADJACENT
  constexpr int W = 32 / G;                                    // cell columns per warp; lane = l*W + column
  const int col = ln % W, l = ln / W, s = (t >> 5) * W + col;
  const unsigned gm = (G == 1 ? 1u : G == 2 ? 0x00010001u : G == 4 ? 0x01010101u : G == 8 ? 0x11111111u : 0x55555555u) << col;
  const unsigned below = (1u << col) - 1u;                     // leaders of the groups before this one
#else
  const int s = t / G, l = t % G;
  const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (ln & ~(G - 1)));
  const unsigned below = (1u << (ln & ~(G - 1))) - 1u;
#endif
#ifndef LANE_ADJACENT
  const unsigned leaders = G == 1 ? 0xffffffffu : (1u << W) - 1u;
#else
  const unsigned leaders = G == 1 ? 0xffffffffu : G == 2 ? 0x55555555u : G == 4 ? 0x11111111u : G == 8 ? 0x01010101u : 0x00010001u;
#endif
  const double *bd = blob;
  const int *bi = reinterpret_cast<const int *>(blob + h.ndbl);
  Lane<N, G> c;
  lane_bind<N, CPB, G>(lt, c, s, l, gm);
  bool active = false, exhausted = s >= lt.cells;              // uniform over a group
  int pending = 0;                                             // exit status waiting for its closing pass
  const double inv_dt = 1.0 / dt;
#pragma unroll 1
  for (;;) {
    bool fresh = false;                                        // this group took a cell in this round of the outer loop
#pragma unroll 1
    for (;;) {                                                 // hand out work to the idle groups of this warp
      const bool want = !active && !exhausted;
      const unsigned wm = __ballot_sync(0xffffffffu, want) & leaders;
      if (wm == 0u) break;
      const int leader = __ffs(wm) - 1;
      unsigned long long base = 0;
      if (ln == leader) base = atomicAdd(counter, (unsigned long long)__popc(wm));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (want) {
        long long i = (long long)base + __popc(wm & below);
        if (i >= nlocal) {
          exhausted = true;
        } else {
          if (S.order) i = S.order[i];                         // slowest cells of the previous call first (rxn_b200.cu: react_order)
          const long long cell = l2g ? l2g[i] : i + cell0;    // cell0: first cell of this chunk of the batch
          if (S.active && !S.active[cell]) {                   // imat <= 0 (reactive_transport.F90:1699)
            if (l == 0) {
              if (iters) iters[i] = 0;
              if (flags) flags[i] = RXN_FLAG_INACTIVE;
            }
          } else {
            lane_load<N, CPB, G>(lt, c, S, bd, bi, h, i, cell, tran_xx, dt);
            active = true;
            fresh = true;
            pending = 0;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;
    // the long arrays of the cells just taken: all 32 lanes of the warp, one cell after the other
    for (unsigned fm = lt.coop_io ? (__ballot_sync(0xffffffffu, fresh) & leaders) : 0u; fm != 0u;) {
      int slot[3], cnt = 0;
      long long cell[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int g = fm != 0u ? __ffs(fm) - 1 : 0;
        slot[q] = __shfl_sync(0xffffffffu, c.s, g);
        cell[q] = __shfl_sync(0xffffffffu, c.cell, g);
        if (fm != 0u) { cnt = q + 1; fm &= fm - 1u; }
      }
      lane_coop_in_sm<N, CPB>(lt, S, slot, cell, cnt, ln, 32);
      if (lt.nmr > 0)
        for (int q = 0; q < cnt; ++q) lane_coop_in_mr<N, CPB>(lt, S, h, bd, bi, slot[q], cell[q], dt, ln, 32);
    }
    __syncwarp();
    bool fin = false;
    if (active) {
      bool recompute;
      const int st = lane_trip<N, CPB, G>(lt, c, S, dt, inv_dt, dt_mode, pending != 0, recompute);
      if (pending != 0) {
        lane_finish<N, CPB, G>(lt, c, S, h, tran_xx, iters, flags, pending);
        active = false; fin = true;
      } else if (st != 0) {
        if (recompute) {
          pending = st;
        } else {
          lane_finish<N, CPB, G>(lt, c, S, h, tran_xx, iters, flags, st);
          active = false; fin = true;
        }
      }
    }
    __syncwarp();
    for (unsigned fm = lt.coop_io ? (__ballot_sync(0xffffffffu, fin) & leaders) : 0u; fm != 0u; fm &= fm - 1u) {
      const int g = __ffs(fm) - 1;
      const int slot = __shfl_sync(0xffffffffu, c.s, g);
      const long long cell = __shfl_sync(0xffffffffu, c.cell, g);
      lane_coop_out<N, CPB>(lt, S, slot, cell, ln, 32);
    }
    __syncwarp();
  }
}

// Global-implicit residual / Jacobian blocks: no Newton loop, every lane group takes cells base+s, base+s+grid*CPB, ...
template <int N, int CPB, int G>
__global__ void __launch_bounds__(((CPB * G + 31) / 32) * 32, 1)
k_gi_lane(const __grid_constant__ LaneTab lt, const __grid_constant__ DevTab h, const double *__restrict__ pblob,
          const double *__restrict__ blob, DevState S, const int32_t *__restrict__ l2g, long long nlocal, double dt, double *res_out,
          double *jac_out) {
  const int words = lt.blob_dbl + lt.blob_int / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) tsm[w] = pblob[w];
  __syncthreads();
  const int t = threadIdx.x, ln = t & 31;
#ifndef LANE_ADJACENT
  constexpr int W = 32 / G;                                    // cell columns per warp; lane = l*W + column
  const int col = ln % W, l = ln / W, s = (t >> 5) * W + col;
  const unsigned gm = (G == 1 ? 1u : G == 2 ? 0x00010001u : G == 4 ? 0x01010101u : G == 8 ? 0x11111111u : 0x55555555u) << col;
#else
  const int s = t / G, l = t % G;
  const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (ln & ~(G - 1)));
#endif
  const double *bd = blob;
  const int *bi = reinterpret_cast<const int *>(blob + h.ndbl);
  Lane<N, G> c;
  lane_bind<N, CPB, G>(lt, c, s, l, gm);
#pragma unroll 1
  for (long long base = (long long)blockIdx.x * CPB; base < nlocal; base += (long long)gridDim.x * CPB) {
    const long long item = base + s;
    if (s < lt.cells && item < nlocal) {
      const long long cell = l2g ? l2g[item] : item;
      if (!(S.active && !S.active[cell])) lane_gi_cell<N, CPB, G>(lt, c, S, bd, bi, h, item, cell, dt, res_out, jac_out);
    }
    __syncwarp();
  }
}

}  // namespace lane

template <>
int lane_launch_gi_variant<LANE_N, LANE_CPB, LANE_G>(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob,
                                                     const double *blob, const DevState &S, const int32_t *l2g, long long nlocal, double dt,
                                                     double *res_out, double *jac_out, cudaStream_t stream) {
  auto kern = lane::k_gi_lane<LANE_N, LANE_CPB, LANE_G>;
  constexpr int threads = ((LANE_CPB * LANE_G + 31) / 32) * 32;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return RXN_ERR_CUDA;
  int bps = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem_bytes) != cudaSuccess || bps < 1) bps = 1;
  const long long want = (nlocal + LANE_CPB - 1) / LANE_CPB;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * bps));
  kern<<<grid, threads, smem_bytes, stream>>>(lt, h, pblob, blob, S, l2g, nlocal, dt, res_out, jac_out);
  return RXN_OK;
}

template <>
int lane_launch_variant<LANE_N, LANE_CPB, LANE_G>(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob,
                                                  const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                                                  long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags,
                                                  unsigned long long *counter, long long cell0, cudaStream_t stream) {
  auto kern = lane::k_react_lane<LANE_N, LANE_CPB, LANE_G>;
  constexpr int threads = ((LANE_CPB * LANE_G + 31) / 32) * 32;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return RXN_ERR_CUDA;
  int bps = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem_bytes) != cudaSuccess || bps < 1) bps = 1;
  const long long want = (nlocal + LANE_CPB - 1) / LANE_CPB;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * bps));
  kern<<<grid, threads, smem_bytes, stream>>>(lt, h, pblob, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, counter, cell0);
  return RXN_OK;
}

}  // namespace rxn
