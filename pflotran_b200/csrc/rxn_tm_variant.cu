// rxn_tm_variant.cu — one instantiation (TM_N x TM_QUADS x TM_G) of the tensor-memory resident RReact kernel
// (device code and design: rxn_tm_dev.cuh; plan: rxn_lane.h with tmG > 0).
#if !defined(TM_N) || !defined(TM_QUADS) || !defined(TM_G)
#error "compile with -DTM_N=<matrix dimension <= 15> -DTM_QUADS=<32-cell quads per CTA, 1..4> -DTM_G=<member warps per cell>"
#endif
#include <algorithm>

#include "rxn_lane.cuh"
#include "rxn_tm_dev.cuh"

namespace rxn {
namespace tmk {

// CTA = 4 G warps; warp w is member w / 4 of quad w % 4 (the TMEM lane quarter its tcgen05.ld/st can reach).  Quads beyond
// QUADS (chemistries whose vectors do not fit 128 cells in shared memory) idle.  Every lane is persistent: a lane without a
// cell takes the next item from the global counter (member 0 of the quad asks, one atomicAdd per warp and round, and
// publishes the items through the exchange slots); each round of the outer loop is one trip of the whole warp through the
// Newton loop, whatever Newton iteration each of its 32 cells is in.
template <int N, int QUADS, int G>
__global__ void __launch_bounds__(128 * G, 1)
k_react_tm(const __grid_constant__ LaneTab lt, const __grid_constant__ DevTab h, const double *__restrict__ pblob,
           const double *__restrict__ blob, DevState S, double *tran_xx, const int32_t *__restrict__ l2g, long long nlocal, double dt,
           int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter, long long cell0) {
  constexpr int CPB = 32 * QUADS;
#ifndef TM_ALLOC_SLOT
#define TM_ALLOC_SLOT 0            /* experiment hook: which word receives the tcgen05.alloc result (compute-sanitizer synccheck reports it as a barrier) */
#endif
  __shared__ unsigned tmem_base_w[4];
  __shared__ double mr_kk[TM_MR_KK];
  const int words = lt.blob_dbl + lt.blob_int / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) tsm[w] = pblob[w];
  tm_mr_kk_fill(mr_kk, threadIdx.x, blockDim.x, blob, h, dt);
  const double *kk = h.nmr * h.mr_ld <= TM_MR_KK ? mr_kk : nullptr;
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"((unsigned long long)__cvta_generic_to_shared(&tmem_base_w[TM_ALLOC_SLOT])));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const unsigned tmem_base = tmem_base_w[TM_ALLOC_SLOT];
  const int quad = warp & 3, l = warp >> 2;
  if (quad < QUADS) {
    const double *bd = blob;
    const int *bi = reinterpret_cast<const int *>(blob + h.ndbl);
    Ctx<N, G> c;
    tm_bind<N, CPB, G>(lt, c, quad * 32 + ln, l, quad, tmem_base);
    tm_init_column<N, CPB, G>(lt, c);
    bool has = false, exhausted = false, closing = false;      // identical in the G members of a cell
    int pending = 0;                                            // exit status waiting for its closing pass
    const double inv_dt = 1.0 / dt;
#pragma unroll 1
    for (;;) {
      bool fresh = false;                                       // this lane took a cell in this round
#pragma unroll 1
      for (;;) {                                                // hand out work to the idle lanes of this warp
        const bool want = !has && !exhausted;
        const unsigned wm = __ballot_sync(0xffffffffu, want);
        if (wm == 0u) break;
        double mine = -1.0;
        if (l == 0) {
          const int leader = __ffs(wm) - 1;
          unsigned long long base = 0;
          if (ln == leader) base = atomicAdd(counter, (unsigned long long)__popc(wm));
          base = __shfl_sync(0xffffffffu, base, leader);
          mine = (double)(base + (unsigned long long)__popc(wm & ((1u << ln) - 1u)));
        }
        double o[G];
        grp_gather<CPB, G>(c, mine, o);                         // item numbers are exact in a double
        if (want) {
          long long i = (long long)o[0];
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
              tm_load<N, CPB, G>(lt, c, S, bd, bi, h, i, cell, tran_xx, dt);
              has = true; closing = false; pending = 0; fresh = true;
            }
          }
        }
        __syncwarp();
      }
      if (!__any_sync(0xffffffffu, has)) break;
      if (lt.nmr > 0) {                                         // the multirate sorbed totals of the cells just taken, one cell after the other
        for (unsigned fm = __ballot_sync(0xffffffffu, fresh); fm != 0u; fm &= fm - 1u) {
          const int g = __ffs(fm) - 1;
          const int slot = __shfl_sync(0xffffffffu, c.s, g);
          const long long cell = __shfl_sync(0xffffffffu, c.cell, g);
          tm_coop_in_mr<N, CPB, G>(lt, S, h, bd, bi, l, slot, cell, dt, ln, 32, kk);
        }
      }
      grp_sync<G>(c);                                           // the loads of every member are in shared memory
      int st;
      bool recompute;
      tm_trip<N, CPB, G>(lt, c, S, dt, inv_dt, dt_mode, has && !closing, has && closing, st, recompute);
      bool fin = false;
      int status = 0;
      if (has) {
        if (closing) { fin = true; status = pending; }
        else if (st != 0) {
          if (recompute) { pending = st; closing = true; }
          else { fin = true; status = st; }
        }
      }
      __syncwarp();
      if (__any_sync(0xffffffffu, fin)) {
        tm_finish<N, CPB, G>(lt, c, S, h, tran_xx, iters, flags, fin, status);
        if (fin) { has = false; closing = false; }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
}

// Global-implicit pass (tm_gi_cell): every lane takes the cells base + column, base + column + grid * CPB, ... - one pass per
// cell, so the rounds of a CTA are uniform and no work counter is needed.
template <int N, int QUADS, int G>
__global__ void __launch_bounds__(128 * G, 1)
k_gi_tm(const __grid_constant__ LaneTab lt, const __grid_constant__ DevTab h, const double *__restrict__ pblob,
        const double *__restrict__ blob, DevState S, const int32_t *__restrict__ l2g, long long nlocal, const __grid_constant__ GiArgs a) {
  constexpr int CPB = 32 * QUADS;
  __shared__ unsigned tmem_base_w[4];
  __shared__ double mr_kk[TM_MR_KK];
  const int words = lt.blob_dbl + lt.blob_int / 2;
  for (int w = threadIdx.x; w < words; w += blockDim.x) tsm[w] = pblob[w];
  tm_mr_kk_fill(mr_kk, threadIdx.x, blockDim.x, blob, h, a.dt);
  const double *kk = h.nmr * h.mr_ld <= TM_MR_KK ? mr_kk : nullptr;
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"((unsigned long long)__cvta_generic_to_shared(&tmem_base_w[0])));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const unsigned tmem_base = tmem_base_w[0];
  const int quad = warp & 3, l = warp >> 2;
  if (quad < QUADS) {
    const double *bd = blob;
    const int *bi = reinterpret_cast<const int *>(blob + h.ndbl);
    Ctx<N, G> c;
    tm_bind<N, CPB, G>(lt, c, quad * 32 + ln, l, quad, tmem_base);
    tm_init_column<N, CPB, G>(lt, c);
#pragma unroll 1
    for (long long base = (long long)blockIdx.x * CPB; base < nlocal; base += (long long)gridDim.x * CPB) {
      const long long item = base + c.s;
      bool on = item < nlocal;
      const long long itc = on ? item : nlocal - 1;             // a lane beyond the batch walks a valid cell with its stores off
      const long long cell = l2g ? l2g[itc] : itc;
      if (S.active && !S.active[cell]) on = false;              // imat <= 0: cycle
      tm_gi_cell<N, CPB, G>(lt, c, S, h, bd, bi, a, itc, cell, on, kk);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
}

}  // namespace tmk

template <>
int tm_launch_gi_variant<TM_N, TM_QUADS, TM_G>(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob,
                                               const double *blob, const DevState &S, const int32_t *l2g, long long nlocal,
                                               const GiArgs &a, cudaStream_t stream) {
  auto kern = tmk::k_gi_tm<TM_N, TM_QUADS, TM_G>;
  constexpr int threads = 128 * TM_G;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return RXN_ERR_CUDA;
  const long long want = (nlocal + 32 * TM_QUADS - 1) / (32 * TM_QUADS);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count));
  kern<<<grid, threads, smem_bytes, stream>>>(lt, h, pblob, blob, S, l2g, nlocal, a);
  return RXN_OK;
}

template <>
int tm_launch_variant<TM_N, TM_QUADS, TM_G>(const LaneTab &lt, size_t smem_bytes, int sm_count, const DevTab &h, const double *pblob,
                                            const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g, long long nlocal,
                                            double dt, int dt_mode, int32_t *iters, int32_t *flags, unsigned long long *counter,
                                            long long cell0, cudaStream_t stream) {
  auto kern = tmk::k_react_tm<TM_N, TM_QUADS, TM_G>;
  constexpr int threads = 128 * TM_G;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return RXN_ERR_CUDA;
  const long long want = (nlocal + 32 * TM_QUADS - 1) / (32 * TM_QUADS);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count));   // one CTA per SM owns its TMEM
  kern<<<grid, threads, smem_bytes, stream>>>(lt, h, pblob, blob, S, tran_xx, l2g, nlocal, dt, dt_mode, iters, flags, counter, cell0);
  return RXN_OK;
}

}  // namespace rxn
