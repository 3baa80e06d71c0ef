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
                                                   double *__restrict__ T_up, double *__restrict__ T_dn) {
  extern __shared__ double sh[];               // [256][n | 1]
  const int ldp = n | 1;
  const long long c0 = (long long)blockIdx.x * 256;
  const long long span = min((long long)256, nconn - c0) * n;
  for (long long k = threadIdx.x; k < span; k += 256) sh[(k / n) * ldp + k % n] = disp[c0 * n + k];   // coalesced AoS read
  __syncthreads();
  const long long c = c0 + threadIdx.x;
  if (c >= nconn) return;
  const double q = velocity[c], a = area[c], fu = use_upwinding ? 0.0 : fraction_upwind[c];
  for (int i = 0; i < n; ++i) {
    double tu, td;
    flux_coef(q, sh[threadIdx.x * ldp + i], a, fu, use_upwinding, &tu, &td);
    T_up[(long long)i * nconn + c] = tu;
    T_dn[(long long)i * nconn + c] = td;
  }
}

// Flux residual of every local row: one warp per tile of 32 rows, lane = row.  r: AoS [nlocal][n].
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
    for (int i = 0; i < n; ++i)
      tile[lane * ldp + i] = flux_row_residual(ent, col, s0, s1, own, total + (long long)i * ld, T_up + (long long)i * nconn,
                                               T_dn + (long long)i * nconn);
  }
  __syncwarp();
  const long long span = min((long long)32, nlocal - row0) * n;
  for (long long k = lane; k < span; k += 32) r[row0 * n + k] = tile[(k / n) * ldp + k % n];
}

// Flux Jacobian in block CSR.  One CTA (8 warps) per tile of 32 rows; slot by slot (0 = diagonal block) the tile's
// blocks are computed with lane = row (coalesced dtotal reads of the row cell / its neighbour), staged in shared
// memory and written out with lane = element of the block.  val: [nnzb][n*n], blocks column-major.
__global__ void __launch_bounds__(256) k_flux_jacobian(int n, long long nlocal, long long nconn, int maxdeg,
                                                      const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                      const int32_t *__restrict__ ent, const int32_t *__restrict__ l2g,
                                                      const double *__restrict__ dtotal, long long ld,
                                                      const double *__restrict__ T_up, const double *__restrict__ T_dn,
                                                      double *__restrict__ val) {
  extern __shared__ double sh[];               // [32][n*n | 1]
  const int nn = n * n, ldp = nn | 1, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long row0 = (long long)blockIdx.x * 32;
  const long long row = row0 + lane;
  int s0 = 0, s1 = 0;
  int32_t own = 0;
  if (row < nlocal) { s0 = row_ptr[row]; s1 = row_ptr[row + 1]; own = l2g[row]; }
  // number of slots any row of the tile has (uniform across the CTA)
  int nslot = s1 - s0;
  for (int o = 16; o; o >>= 1) nslot = max(nslot, __shfl_xor_sync(0xffffffffu, nslot, o));
  const int rows_here = (int)min((long long)32, nlocal - row0);
  for (int k = 0; k < nslot; ++k) {
    if (k < s1 - s0) {
      if (k == 0) {
        for (int e = w; e < nn; e += 8) {
          const int i = e % n;
          sh[lane * ldp + e] = flux_row_jac_diag(ent, s0, s1, dtotal[(long long)e * ld + own], T_up + (long long)i * nconn,
                                                 T_dn + (long long)i * nconn);
        }
      } else {
        const int32_t en = ent[s0 + k], nb = col[s0 + k];
        for (int e = w; e < nn; e += 8) {
          const int i = e % n;
          sh[lane * ldp + e] = flux_row_jac_off(en, dtotal[(long long)e * ld + nb], T_up + (long long)i * nconn,
                                                T_dn + (long long)i * nconn);
        }
      }
    }
    __syncthreads();
    for (int rr = w; rr < rows_here; rr += 8) {
      const int t0 = __shfl_sync(0xffffffffu, s0, rr), t1 = __shfl_sync(0xffffffffu, s1, rr);
      if (k < t1 - t0) {
        double *dst = val + (long long)(t0 + k) * nn;
        for (int e = lane; e < nn; e += 32) dst[e] = sh[rr * ldp + e];
      }
    }
    __syncthreads();
  }
}

}  // namespace rxn
