// rxn_flux.cuh — CUDA kernels of the flux side (rxn_flux.h has the design and the reference citations).
//
// All three kernels are HBM bound.  Cell-fastest SoA reads (total, dtotal, coefficients) are coalesced with one lane
// per row cell; the outputs are AoS (PETSc Vec: components of a cell contiguous; MATBAIJ: the n x n block of a slot
// contiguous), so every tile is transposed through shared memory and written with one lane per element.
#pragma once
#include <cuda_runtime.h>
#include "rxn_flux.h"

namespace rxn {

// TFluxCoef for every connection.  disp: AoS [connection][component] as patch%internal_tran_coefs(:,1,conn);
// T_up/T_dn: SoA [component][connection].
__global__ void __launch_bounds__(256) k_flux_coefs(int n, long long nconn, const double *__restrict__ area,
                                                   const double *__restrict__ velocity, const double *__restrict__ disp,
                                                   const double *__restrict__ fraction_upwind, int use_upwinding,
                                                   double *__restrict__ T_up, double *__restrict__ T_dn, double *__restrict__ T_aos = nullptr) {
  extern __shared__ double sh[];               // [256][n | 1]
  const int ldp = n | 1;
  const long long c0 = (long long)blockIdx.x * 256;
  const long long span = min((long long)256, nconn - c0) * n;
  for (long long k = threadIdx.x; k < span; k += 256) sh[(k / n) * ldp + k % n] = disp[c0 * n + k];   // coalesced AoS read
  __syncthreads();
  const long long c = c0 + threadIdx.x;
  if (c < nconn) {
    const double q = velocity[c], a = area[c], fu = use_upwinding ? 0.0 : fraction_upwind[c];
    for (int i = 0; i < n; ++i) {
      double tu, td;
      flux_coef(q, sh[threadIdx.x * ldp + i], a, fu, use_upwinding, &tu, &td);
      T_up[(long long)i * nconn + c] = tu;
      T_dn[(long long)i * nconn + c] = td;
    }
  }
  if (T_aos) {
    // the same coefficients as [connection][up | dn][component] (what the Jacobian's column walk gathers: 4 sectors per entry instead
    // of n), written with one lane per element of the CTA's contiguous range; the element is re-evaluated (same operations, same bits)
    const long long span2 = min((long long)256, nconn - c0) * 2 * n;
    for (long long k = threadIdx.x; k < span2; k += 256) {
      const int cl = (int)(k / (2 * n)), r = (int)(k % (2 * n)), i = r % n;
      const long long cc = c0 + cl;
      double tu, td;
      flux_coef(velocity[cc], sh[cl * ldp + i], area[cc], use_upwinding ? 0.0 : fraction_upwind[cc], use_upwinding, &tu, &td);
      T_aos[c0 * 2 * n + k] = r < n ? tu : td;
    }
  }
}

// Flux residual of every local row: one warp per tile of 32 rows, lane = row.  r: AoS [nlocal][n].
// (Keeping the row's entries in registers across the whole component loop measured slower: 0.52 against 0.41 ms per 10^6 rows.)
enum { FLUX_Q = 6, FLUX_JC = 8 };
#ifndef FLUX_RC
#define FLUX_RC 2
#endif

// dtotal is read up to 7 times (the cell itself and its neighbours' rows) while the matrix streams through L2 once:
// loads ask L2 to keep the line (evict_last), stores are streaming (st.global.cs)
__device__ __forceinline__ unsigned long long flux_keep_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double flux_ld_keep(const double *p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}

__global__ void __launch_bounds__(128) k_flux_residual(int n, long long nlocal, long long nconn, const int32_t *__restrict__ row_ptr,
                                                      const int32_t *__restrict__ col, const int32_t *__restrict__ ent,
                                                      const int32_t *__restrict__ l2g, const double *__restrict__ total, long long ld,
                                                      const double *__restrict__ T_up, const double *__restrict__ T_dn,
                                                      double *__restrict__ r) {
  extern __shared__ double sh[];               // [4 warps][32][n | 1]
  const int ldp = n | 1, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *tile = sh + (size_t)w * 32 * ldp;
  const long long row0 = ((long long)blockIdx.x * 4 + w) * 32;
  if (row0 >= nlocal) return;
  const long long row = row0 + lane;
  if (row < nlocal) {
    const int s0 = row_ptr[row], s1 = row_ptr[row + 1];
    const int32_t own = l2g[row];
    // FLUX_RC components per walk of the row: the entry / neighbour loads are shared and 3*FLUX_RC value loads are independent;
    // each component's sum is formed in connection order as flux_row_residual (rxn_flux.h) does
    int i = 0;
    for (; i + FLUX_RC <= n; i += FLUX_RC) {
      double acc[FLUX_RC], t_own[FLUX_RC];
#pragma unroll
      for (int u = 0; u < FLUX_RC; ++u) { acc[u] = 0.0; t_own[u] = total[(long long)(i + u) * ld + own]; }
      for (int s = s0 + 1; s < s1; ++s) {
        const int32_t e = ent[s], c = e >> 1, nb = col[s];
#pragma unroll
        for (int u = 0; u < FLUX_RC; ++u) {
          const double t_nb = total[(long long)(i + u) * ld + nb];
          const double tu = T_up[(long long)(i + u) * nconn + c], td = T_dn[(long long)(i + u) * nconn + c];
          acc[u] = (e & 1) ? fl_add(acc[u], -flux_res(tu, t_nb, td, t_own[u])) : fl_add(acc[u], flux_res(tu, t_own[u], td, t_nb));
        }
      }
#pragma unroll
      for (int u = 0; u < FLUX_RC; ++u) tile[lane * ldp + i + u] = acc[u];
    }
    for (; i < n; ++i)
      tile[lane * ldp + i] = flux_row_residual(ent, col, s0, s1, own, total + (long long)i * ld, T_up + (long long)i * nconn,
                                               T_dn + (long long)i * nconn);
  }
  __syncwarp();
  const long long span = min((long long)32, nlocal - row0) * n;
  for (long long k = lane; k < span; k += 32) r[row0 * n + k] = tile[(k / n) * ldp + k % n];
}

// Flux Jacobian in block CSR.  val: [nnzb][n*n], blocks column-major.
// grid = (tiles of 32 rows, column chunks of the block): a CTA of 4 warps owns jc columns (cw = jc*n elements) of every block of
// its 32 rows.  Slot by slot (0 = diagonal block) the chunk is computed with lane = row — coalesced cell-fastest reads of
// dtotal of the row cell / its neighbour, the jc loads of a block row issued together, the coefficient of (slot, i) loaded
// once — staged in shared memory ([32][cw | 1], conflict free both ways) and written with lane = element as contiguous runs of
// cw doubles per block, streaming (the matrix is written once and not read here: keep L2 for the neighbours' dtotal).
__global__ void __launch_bounds__(128) k_flux_jacobian(int n, int jc, long long nlocal, long long nconn,
                                                      const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                      const int32_t *__restrict__ ent, const int32_t *__restrict__ l2g,
                                                      const double *__restrict__ dtotal, long long ld,
                                                      const double *__restrict__ T_up, const double *__restrict__ T_dn,
                                                      double *__restrict__ val) {
  extern __shared__ double sh[];               // [32][cw | 1]
  __shared__ int sh_s0[32], sh_ns[32];
  const int nn = n * n, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j0 = blockIdx.y * jc, jn = min(jc, n - j0);           // columns j0 .. j0+jn-1 of every block
  const int cw = jn * n, ldp = (jc * n) | 1;
  const long long row0 = (long long)blockIdx.x * 32;
  const long long row = row0 + lane;
  int s0 = 0, s1 = 0;
  int32_t own = 0;
  if (row < nlocal) { s0 = row_ptr[row]; s1 = row_ptr[row + 1]; own = l2g[row]; }
  const int ns = s1 - s0, deg = ns - 1;
  if (w == 0) { sh_s0[lane] = s0; sh_ns[lane] = ns; }
  int nslot = ns;                                                 // slots of the longest row of the tile (CTA uniform)
  for (int o = 16; o; o >>= 1) nslot = max(nslot, __shfl_xor_sync(0xffffffffu, nslot, o));
  const int rows_here = (int)min((long long)32, nlocal - row0);
  const int lsh = cw > 16 ? 5 : cw > 8 ? 4 : cw > 4 ? 3 : 2;      // log2 of the lanes that write one row's run
  int32_t en[FLUX_Q];
#pragma unroll
  for (int q = 0; q < FLUX_Q; ++q) en[q] = q < deg ? ent[s0 + 1 + q] : 0;
  const double *D0 = dtotal + (long long)j0 * n * ld;             // element (i, j0 + jj) of cell c at D0[(jj*n + i)*ld + c]
  for (int k = 0; k < nslot; ++k) {
    if (k < ns) {
      if (k == 0) {
        for (int i = w; i < n; i += 4) {
          const double *Tu_i = T_up + (long long)i * nconn, *Td_i = T_dn + (long long)i * nconn;
          double d[FLUX_JC], sc[FLUX_Q];
#pragma unroll
          for (int jj = 0; jj < FLUX_JC; ++jj) d[jj] = jj < jn ? D0[(long long)(jj * n + i) * ld + own] : 0.0;
#pragma unroll
          for (int q = 0; q < FLUX_Q; ++q) sc[q] = q < deg ? ((en[q] & 1) ? -Td_i[en[q] >> 1] : Tu_i[en[q] >> 1]) : 0.0;
#pragma unroll
          for (int jj = 0; jj < FLUX_JC; ++jj)
            if (jj < jn) {
              double a = 0.0;
#pragma unroll
              for (int q = 0; q < FLUX_Q; ++q) if (q < deg) a = fl_add(a, fl_mul(d[jj], sc[q]));
              for (int s = s0 + 1 + FLUX_Q; s < s1; ++s) {
                const int32_t e = ent[s];
                a = fl_add(a, fl_mul(d[jj], (e & 1) ? -Td_i[e >> 1] : Tu_i[e >> 1]));
              }
              sh[lane * ldp + jj * n + i] = a;
            }
        }
      } else {
        const int32_t e = k <= FLUX_Q ? 0 : ent[s0 + k];
        int32_t ek = e;
#pragma unroll
        for (int q = 0; q < FLUX_Q; ++q) if (k == q + 1) ek = en[q];
        const int32_t nb = col[s0 + k], c = ek >> 1;
        for (int i = w; i < n; i += 4) {
          const double so = (ek & 1) ? -T_up[(long long)i * nconn + c] : T_dn[(long long)i * nconn + c];
          double d[FLUX_JC];
#pragma unroll
          for (int jj = 0; jj < FLUX_JC; ++jj) d[jj] = jj < jn ? D0[(long long)(jj * n + i) * ld + nb] : 0.0;
#pragma unroll
          for (int jj = 0; jj < FLUX_JC; ++jj) if (jj < jn) sh[lane * ldp + jj * n + i] = fl_mul(d[jj], so);
        }
      }
    }
    __syncthreads();
    // lpr lanes per row (a power of two >= min(cw, 32)): 32/lpr rows per warp pass
    for (int rr = w * (32 >> lsh) + (lane >> lsh); rr < rows_here; rr += 4 * (32 >> lsh))
      if (k < sh_ns[rr]) {
        double *dst = val + (long long)(sh_s0[rr] + k) * nn + j0 * n;
        const double *src = sh + rr * ldp;
        for (int e = lane & ((1 << lsh) - 1); e < cw; e += 1 << lsh) __stcs(dst + e, src[e]);
      }
    __syncthreads();
  }
}

// Same kernel with the block size known at compile time (the BASELINE chemistries: N = 15 in chunks of 5 columns, N = 4 and
// N = 3 whole): every index is an immediate, the loops are unrolled, and the staging tile is double buffered so that one
// barrier per slot suffices and the loads of slot k+1 are in flight while slot k is written out.
#ifndef FLUX_MINB
#define FLUX_MINB 4
#endif
// FLUX_MODE (experiments only, profiles/): 1 = no matrix stores, 2 = no dtotal loads
#ifndef FLUX_MODE
#define FLUX_MODE 0
#endif
#if FLUX_MODE == 2
#define FLUX_LD(p, pol) ((double)((p) - (const double *)0) * 1e-30)
#else
#define FLUX_LD(p, pol) flux_ld_keep(p, pol)
#endif
#if FLUX_MODE == 1
#define FLUX_ST(p, v) do { if ((v) == 1.2345e-300) __stcs(p, v); } while (0)
#else
#define FLUX_ST(p, v) __stcs(p, v)
#endif
#ifndef FLUX_DU
#define FLUX_DU 2
#endif
#ifndef FLUX_EARLY
#define FLUX_EARLY 0
#endif
#ifndef FLUX_NW
#define FLUX_NW 4
#endif
template <int N, int JC>
__global__ void __launch_bounds__(32 * FLUX_NW, FLUX_MINB) k_flux_jacobian_t(long long nlocal, long long nconn, const int32_t *__restrict__ row_ptr,
                                                        const int32_t *__restrict__ col, const int32_t *__restrict__ ent,
                                                        const int32_t *__restrict__ l2g, const double *__restrict__ dtotal, long long ld,
                                                        const double *__restrict__ T_up, const double *__restrict__ T_dn,
                                                        double *__restrict__ val) {
  static_assert(N % JC == 0, "column chunks must tile the block");
  constexpr int NN = N * N, CW = JC * N, LDP = CW | 1, NW = FLUX_NW;   // NW warps per CTA
  constexpr int LSH = CW > 16 ? 5 : CW > 8 ? 4 : CW > 4 ? 3 : 2, LPR = 1 << LSH, RPP = 32 >> LSH;   // lanes per row run, rows per warp pass
  __shared__ double sh[2][32 * LDP];
  __shared__ int sh_s0[32], sh_ns[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NCH = N / JC;                                      // the chunks of a tile are adjacent blocks: their partial
  const int j0 = (int)(blockIdx.x % NCH) * JC;                     // sectors at the chunk seams meet in L2
  const long long row0 = (long long)(blockIdx.x / NCH) * 32;
  const unsigned long long keep = flux_keep_policy();
  const long long row = row0 + lane;
  int s0 = 0, s1 = 0;
  int32_t own = 0;
  if (row < nlocal) { s0 = row_ptr[row]; s1 = row_ptr[row + 1]; own = l2g[row]; }
  const int ns = s1 - s0, deg = ns - 1;
  if (w == 0) { sh_s0[lane] = s0; sh_ns[lane] = ns; }
  int nslot = ns;
  for (int o = 16; o; o >>= 1) nslot = max(nslot, __shfl_xor_sync(0xffffffffu, nslot, o));
  const int rows_here = (int)min((long long)32, nlocal - row0);
  int32_t en[FLUX_Q];
#pragma unroll
  for (int q = 0; q < FLUX_Q; ++q) en[q] = q < deg ? ent[s0 + 1 + q] : 0;
  const long long sj = (long long)N * ld;                          // stride of one block column in dtotal
  const double *D0 = dtotal + (long long)j0 * sj;
  constexpr int NI = (N + NW - 1) / NW;
  int32_t nbr[FLUX_Q];
#pragma unroll
  for (int q = 0; q < FLUX_Q; ++q) nbr[q] = q < deg ? col[s0 + 1 + q] : own;
  double d[NI][JC], so[NI];
  // loads of off-diagonal slot k (all NI*JC + NI requests of the thread back to back); the signed coefficient of the
  // neighbour's side is applied in finish(): -(D T) == D (-T) exactly
  double sgn = 1.0;
  auto issue = [&](int k) {
    if (k >= ns) return;
    int32_t ek = k <= FLUX_Q ? 0 : ent[s0 + k], nb = k <= FLUX_Q ? own : col[s0 + k];
#pragma unroll
    for (int q = 0; q < FLUX_Q; ++q) if (k == q + 1) { ek = en[q]; nb = nbr[q]; }
    const double *To = ((ek & 1) ? T_up : T_dn) + (ek >> 1);
    sgn = (ek & 1) ? -1.0 : 1.0;
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) {
      const int i = w + NW * ii;
      if (i < N) {
        const double *p = D0 + (long long)i * ld + nb;
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) d[ii][jj] = FLUX_LD(p + jj * sj, keep);
        so[ii] = To[(long long)i * nconn];
      }
    }
  };
  auto finish = [&](int k) {
    if (k >= ns) return;
    double *tile = sh[k & 1] + lane * LDP;
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) {
      const int i = w + NW * ii;
      if (i < N) {
        const double sv = sgn * so[ii];
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) tile[jj * N + i] = fl_mul(d[ii][jj], sv);
      }
    }
  };
  // rows of the tile this lane helps to write (RPP rows per warp pass): slot count and element offset of slot 0 in registers
  constexpr int NP = (32 + NW * RPP - 1) / (NW * RPP);
  int wns[NP];
  long long wbase[NP];
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NP; ++t) {
    const int rr = w * RPP + (lane >> LSH) + t * NW * RPP;
    wns[t] = rr < rows_here ? sh_ns[rr] : 0;
    wbase[t] = (long long)sh_s0[rr < 32 ? rr : 0] * NN + j0 * N + (lane & (LPR - 1));
  }
  auto write_out = [&](int k) {
    const double *stage = sh[k & 1] + (lane & (LPR - 1));
#pragma unroll
    for (int t = 0; t < NP; ++t)
      if (k < wns[t]) {
        const int rr = w * RPP + (lane >> LSH) + t * NW * RPP;
        double *dst = val + wbase[t] + (long long)k * NN;
        const double *src = stage + rr * LDP;
#pragma unroll
        for (int u = 0; u < (CW + LPR - 1) / LPR; ++u)
          if ((lane & (LPR - 1)) + u * LPR < CW) FLUX_ST(dst + u * LPR, src[u * LPR]);
      }
  };
  if (FLUX_EARLY) issue(1);
  // slot 0, the diagonal block: two block rows at a time, their 2*(JC + FLUX_Q) loads issued together
  if (ns > 0) {
    double *tile = sh[0] + lane * LDP;
#pragma unroll
    for (int i2 = 0; i2 < NI; i2 += FLUX_DU) {
      double dd[FLUX_DU][JC], sc[FLUX_DU][FLUX_Q];
#pragma unroll
      for (int u = 0; u < FLUX_DU; ++u) {
        const int i = w + NW * (i2 + u);
        if (i2 + u < NI && i < N) {
          const double *Tu_i = T_up + (long long)i * nconn, *Td_i = T_dn + (long long)i * nconn, *p = D0 + (long long)i * ld + own;
#pragma unroll
          for (int jj = 0; jj < JC; ++jj) dd[u][jj] = FLUX_LD(p + jj * sj, keep);
#pragma unroll
          for (int q = 0; q < FLUX_Q; ++q) sc[u][q] = q < deg ? ((en[q] & 1) ? -Td_i[en[q] >> 1] : Tu_i[en[q] >> 1]) : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < FLUX_DU; ++u) {
        const int i = w + NW * (i2 + u);
        if (i2 + u < NI && i < N) {
          const double *Tu_i = T_up + (long long)i * nconn, *Td_i = T_dn + (long long)i * nconn;
#pragma unroll
          for (int jj = 0; jj < JC; ++jj) {
            double a = 0.0;
#pragma unroll
            for (int q = 0; q < FLUX_Q; ++q) if (q < deg) a = fl_add(a, fl_mul(dd[u][jj], sc[u][q]));
            for (int s = s0 + 1 + FLUX_Q; s < s1; ++s) {
              const int32_t e = ent[s];
              a = fl_add(a, fl_mul(dd[u][jj], (e & 1) ? -Td_i[e >> 1] : Tu_i[e >> 1]));
            }
            tile[jj * N + i] = a;
          }
        }
      }
    }
  }
  // off-diagonal slots, software pipelined: the loads of slot k+1 are in flight while slot k crosses the barrier and is written
  if (!FLUX_EARLY) issue(1);
  __syncthreads();
  write_out(0);
  for (int k = 1; k < nslot; ++k) {
    finish(k);
    issue(k + 1);
    __syncthreads();
    write_out(k);
  }
}

// Flux Jacobian by block COLUMNS (FluxCols, rxn_flux.h): a warp owns 4 consecutive ghosted cells.  Lane = element of the
// n x n block, e = lane + LW p with LW = the largest multiple of n that fits a warp (n = 15: 30 lanes, 8 passes), so that a
// lane's block row i = e mod n is the same in every pass: ONE coefficient per lane and block.  The 4 cells' dtotal values of an
// element share one 32-byte sector of the cell-fastest field: each lane loads them with two 16-byte loads - every dtotal
// sector crosses the memory system once - and keeps them in registers (NP x 4 doubles); then every block of the 4 block columns
// is a product by the slot's coefficient (diagonal: the sum over the row's own connections, in connection order) written as NP
// coalesced runs of LW doubles, straight from registers: no shared-memory staging, no barrier, no neighbour reads.  Entries
// are taken in chunks of 8 - lane u reads the descriptor of entry u, then every lane fetches its coefficient of all 8 entries
// back to back (one round trip to L2 per chunk, not per block).  Same products and sums as the row walk (flux_row_jac_*).
template <int N>
__global__ void __launch_bounds__(128, 4) k_flux_jacobian_cols(long long nghosted, long long nconn, const int32_t *__restrict__ col_ptr,
                                                              const int32_t *__restrict__ tgt_slot, const int32_t *__restrict__ tgt_ent,
                                                              const int32_t *__restrict__ col_row, const int32_t *__restrict__ row_ptr,
                                                              const int32_t *__restrict__ ent, const double *__restrict__ dtotal, long long ld,
                                                              const double *__restrict__ T_aos, double *__restrict__ val) {
  constexpr int NN = N * N, LW = (32 / N) * N, NP = (NN + LW - 1) / LW;
  static_assert(N <= 32, "one block row per lane");
  const int lane = threadIdx.x & 31;
  const long long c0 = 4 * ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (c0 >= nghosted) return;
  int cp[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) cp[q] = col_ptr[min(c0 + q, nghosted)];
  if (cp[4] == cp[0]) return;                                     // no block column here (ghost cells away from the local rows)
  const bool work = lane < LW;
  const double *Trow = T_aos + lane % N;                          // this lane's block row in the [connection][up | dn][component] coefficients
  double dq[NP][4];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int e = lane + LW * p;                                  // e = j*N + i
    dq[p][0] = dq[p][1] = dq[p][2] = dq[p][3] = 0.0;
    if (work && e < NN) {
      const double2 *src = reinterpret_cast<const double2 *>(dtotal + (long long)e * ld + c0);   // ld is a multiple of 32: aligned
      const double2 a = __ldg(src), b = __ldg(src + 1);
      dq[p][0] = a.x; dq[p][1] = a.y; dq[p][2] = b.x; dq[p][3] = b.y;
    }
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    int diag_slot = -1;
#pragma unroll 1
    for (int t0 = cp[cc]; t0 < cp[cc + 1]; t0 += 8) {
      int my_ent = -2, my_slot = 0;
      if (lane < 8 && t0 + lane < cp[cc + 1]) { my_ent = tgt_ent[t0 + lane]; my_slot = tgt_slot[t0 + lane]; }
      double cf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int en = __shfl_sync(0xffffffffu, my_ent, u);
        cf[u] = 0.0;
        // row is dn (odd entry): -Jup of this cell = -(D T_up); row is up: +Jdn = D T_dn.  Slot of (connection, side) = 2 c + side
        if (en >= 0 && work) cf[u] = ((en & 1) ? -1.0 : 1.0) * __ldg(Trow + (long long)(en ^ 1) * N);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int en = __shfl_sync(0xffffffffu, my_ent, u);
        const int slot = __shfl_sync(0xffffffffu, my_slot, u);
        if (en == -1) diag_slot = slot;
        if (en >= 0) {                                            // off-diagonal block of the slot's row: +Jdn / -Jup of this cell
          double *dst = val + (long long)slot * NN + lane;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (work && lane + LW * p < NN) __stcs(dst + LW * p, fl_mul(dq[p][cc], cf[u]));
        }
      }
    }
    if (diag_slot >= 0) {                                         // diagonal block: the row's connections in connection order
      const int row = col_row[c0 + cc];
      const int s0 = row_ptr[row] + 1, s1 = row_ptr[row + 1];
      double a[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) a[p] = 0.0;
#pragma unroll 1
      for (int sb = s0; sb < s1; sb += 8) {
        int es_mine = -2;
        if (lane < 8 && sb + lane < s1) es_mine = ent[sb + lane];
        double sc[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int es = __shfl_sync(0xffffffffu, es_mine, v);
          sc[v] = 0.0;
          if (es >= 0 && work) sc[v] = ((es & 1) ? -1.0 : 1.0) * __ldg(Trow + (long long)es * N);   // own side: up for an even entry, dn for an odd one
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          if (sb + v < s1) {
#pragma unroll
            for (int p = 0; p < NP; ++p) a[p] = fl_add(a[p], fl_mul(dq[p][cc], sc[v]));
          }
        }
      }
      double *dst = val + (long long)diag_slot * NN + lane;
#pragma unroll
      for (int p = 0; p < NP; ++p)
        if (work && lane + LW * p < NN) __stcs(dst + LW * p, a[p]);
    }
  }
}

// ---- coupler connections (boundary conditions, source/sinks; rxn_flux.h) ---------------------------------------------
// The sets are small (a grid's faces / wells): one thread per (row, component) resp. (row, block element); the row's
// connections are added in connection order onto the values the interior kernels left in r / in the diagonal block.

// TSrcSinkCoef per connection, coefficients repeated per component so that both kinds share the residual/Jacobian kernels
__global__ void __launch_bounds__(256) k_coupler_ss_coefs(int n, long long nconn, const double *__restrict__ qsrc,
                                                         const int32_t *__restrict__ type, double *__restrict__ c_ext,
                                                         double *__restrict__ c_cell) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= nconn) return;
  double tin, tout;
  ss_coef(qsrc[c], type[c], &tin, &tout);
  for (int i = 0; i < n; ++i) { c_ext[(long long)i * nconn + c] = tout; c_cell[(long long)i * nconn + c] = tin; }
}

// AoS [connection][component] -> SoA [component][connection] (external totals handed over by the host or taken from the
// TOTAL field of a boundary state: src_ld > 0 means src is SoA [component][src_ld] already and is copied row by row)
__global__ void __launch_bounds__(256) k_coupler_totals(int n, long long nconn, const double *__restrict__ src, long long src_ld,
                                                       double *__restrict__ ext) {
  const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
  if (k >= nconn * n) return;
  const long long c = k / n;
  const int i = (int)(k % n);
  ext[(long long)i * nconn + c] = src_ld > 0 ? src[(long long)i * src_ld + c] : src[k];
}

__global__ void __launch_bounds__(128) k_coupler_residual(int n, long long nrows, long long nconn, double sgn,
                                                         const int32_t *__restrict__ row, const int32_t *__restrict__ own,
                                                         const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ conn,
                                                         const double *__restrict__ total, long long ld,
                                                         const double *__restrict__ ext, const double *__restrict__ c_ext,
                                                         const double *__restrict__ c_cell, double *__restrict__ r,
                                                         double *__restrict__ flux_out) {
  const long long k = (long long)blockIdx.x * 128 + threadIdx.x;
  if (k >= nrows * n) return;
  const long long q = k / n;
  const int i = (int)(k % n);
  const int s0 = row_ptr[q], s1 = row_ptr[q + 1];
  const double t_own = total[(long long)i * ld + own[q]];
  const double *ext_i = ext + (long long)i * nconn, *cx_i = c_ext + (long long)i * nconn, *cc_i = c_cell + (long long)i * nconn;
  double *rp = r + (long long)row[q] * n + i;
  *rp = coupler_row_residual(*rp, conn, s0, s1, sgn, ext_i, cx_i, cc_i, t_own);
  if (flux_out)                                                 // patch%boundary_tran_fluxes = -Res / patch%ss_tran_fluxes = Res
    for (int s = s0; s < s1; ++s) {
      const int32_t c = conn[s];
      flux_out[(long long)c * n + i] = fl_mul(sgn, coupler_res(cx_i[c], ext_i[c], cc_i[c], t_own));
    }
}

// diag: base of the block array, blk[q] = index of the diagonal block of row q in it (block-CSR slot row_ptr[row] of the
// connection set, or the local row itself for a plain [nlocal][n*n] array)
__global__ void __launch_bounds__(128) k_coupler_jacobian(int n, long long nrows, long long nconn, double sgn,
                                                         const int32_t *__restrict__ row, const int32_t *__restrict__ own,
                                                         const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ conn,
                                                         const int32_t *__restrict__ csr_row_ptr, const double *__restrict__ dtotal,
                                                         long long ld, const double *__restrict__ c_cell, double *__restrict__ val) {
  const int nn = n * n;
  const long long k = (long long)blockIdx.x * 128 + threadIdx.x;
  if (k >= nrows * nn) return;
  const long long q = k / nn;
  const int e = (int)(k % nn), i = e % n;                       // e = j*n + i (column-major block)
  const long long blk = csr_row_ptr ? (long long)csr_row_ptr[row[q]] : (long long)row[q];
  double *vp = val + blk * nn + e;
  *vp = coupler_row_jac(*vp, conn, row_ptr[q], row_ptr[q + 1], sgn, c_cell + (long long)i * nconn, dtotal[(long long)e * ld + own[q]]);
}

}  // namespace rxn
