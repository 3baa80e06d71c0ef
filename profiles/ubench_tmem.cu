// ubench_tmem.cu — microbenchmark behind DESIGN 4.3 (round 2): can tensor memory hold the per-cell Newton matrix?
// Measures, for one CTA per SM with NW warps, the throughput and latency of tcgen05.ld / tcgen05.st in the
// 32x32b shape (each thread reads/writes consecutive 32-bit columns of ITS OWN TMEM lane: TMEM used as a
// 512-word per-thread scratchpad for the 128 threads of lane quarter warp%4), next to ld.shared / st.shared of
// the same volume and a DFMA peak probe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem ubench_tmem.cu && ./ubench_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};\n"
      :
      : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31]), "r"(taddr));
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t &a, uint32_t &b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// mode 0: LDTM x32 throughput (independent loads, one wait per load)   mode 1: STTM x32 throughput
// mode 2: LDTM x2 dependent chain (latency)                            mode 3: load-FMA-store row update (the LU inner step)
// mode 4: the same row update in shared memory (LDS.128 / STS.128)     mode 5: DFMA peak
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_bench(int iters, unsigned long long *cycles, double *sink) {
  extern __shared__ __align__(16) double sm[];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nthreads = blockDim.x;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"((uint64_t)__cvta_generic_to_shared(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // this warp's lane quarter; warps w and w+4 share a quarter and split the 512 columns
  const int nshare = (nthreads / 32 + 3) / 4;
  const int cols_per = 512 / nshare;
  const uint32_t tbase = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols_per);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = lane + i;
  for (int c = 0; c < cols_per; c += 32) tmem_st32(tbase + c, r);
  tmem_wait_st();
  if (MODE == 4) for (int i = threadIdx.x; i < 16 * 1024; i += nthreads) sm[i] = 1.0 + i;
  __syncthreads();
  double acc0 = 1.0, acc1 = 2.0, acc2 = 3.0, acc3 = 4.0;
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
      tmem_ld32(tbase + (it * 32) % cols_per, r);
      tmem_wait_ld();
      acc0 += __hiloint2double(r[1], r[0]);
    }
  } else if (MODE == 1) {
    for (int it = 0; it < iters; ++it) {
      r[0] = it;
      tmem_st32(tbase + (it * 32) % cols_per, r);
      tmem_wait_st();
    }
  } else if (MODE == 2) {
    uint32_t a = 0, b = 0;
    for (int it = 0; it < iters; ++it) {
      tmem_ld2(tbase + (a & 30), a, b);
      tmem_wait_ld();
    }
    acc0 += a + b;
  } else if (MODE == 3) {
    // row update a[0..15] -= l * p[0..15]: 16 doubles = 32 columns per row
    double p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = 1e-9 * (j + 1);
    for (int it = 0; it < iters; ++it) {
      const uint32_t ta = tbase + (it * 32) % cols_per;
      tmem_ld32(ta, r);
      tmem_wait_ld();
      const double l = __hiloint2double(r[1], r[0]) * 1e-3;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double a = __hiloint2double(r[2 * j + 1], r[2 * j]);
        a = fma(-l, p[j], a);
        r[2 * j] = __double2loint(a);
        r[2 * j + 1] = __double2hiint(a);
      }
      tmem_st32(ta, r);
      tmem_wait_st();
    }
    acc0 += __hiloint2double(r[1], r[0]);
  } else if (MODE == 4) {
    double p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = 1e-9 * (j + 1);
    double2 *s2 = reinterpret_cast<double2 *>(sm);
    const int rows = 1024 / nthreads;                            // 128 KB of shared memory per CTA
    for (int it = 0; it < iters; ++it) {
      const int row = it % rows;
      double2 a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = s2[(row * 8 + q) * nthreads + threadIdx.x];
      const double l = a[0].x * 1e-3;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        a[q].x = fma(-l, p[2 * q], a[q].x);
        a[q].y = fma(-l, p[2 * q + 1], a[q].y);
        s2[(row * 8 + q) * nthreads + threadIdx.x] = a[q];
      }
    }
    acc0 += s2[threadIdx.x].x;
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc0 = fma(acc0, 1.0000001, 1e-9); acc1 = fma(acc1, 1.0000001, 1e-9);
        acc2 = fma(acc2, 1.0000001, 1e-9); acc3 = fma(acc3, 1.0000001, 1e-9);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * nthreads + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
}

template <int MODE>
int run(const char *name, int nw, int iters, double bytes_per_iter_per_thread, double flops_per_iter_per_thread) {
  int nsm = 148;
  unsigned long long *cyc;
  double *sink;
  CK(cudaMalloc(&cyc, nsm * sizeof(unsigned long long)));
  CK(cudaMalloc(&sink, (size_t)nsm * 512 * sizeof(double)));
  CK(cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const size_t sm_bytes = MODE == 4 ? (size_t)128 * 1024 : 1024;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_bench<MODE><<<nsm, nw * 32, sm_bytes>>>(iters / 10, cyc, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_bench<MODE><<<nsm, nw * 32, sm_bytes>>>(iters, cyc, sink);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[148];
  CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
  double c = 0;
  for (int i = 0; i < nsm; ++i) c += (double)h[i];
  c /= nsm;
  const double per_iter = c / iters;
  printf("%-44s nw=%2d  %8.1f cyc/iter/warp  %8.1f B/clk/SM  %8.2f TFLOP/s chip  (%.3f ms)\n", name, nw, per_iter,
         bytes_per_iter_per_thread * nw * 32 / per_iter, flops_per_iter_per_thread * nw * 32 * nsm * iters / (ms * 1e-3) * 1e-12, ms);
  cudaFree(cyc); cudaFree(sink);
  return 0;
}

int main() {
  for (int nw : {4, 8, 16}) {
    run<0>("LDTM 32x32b.x32 (128 B/thread) + wait", nw, 20000, 128, 0);
    run<1>("STTM 32x32b.x32 (128 B/thread) + wait", nw, 20000, 128, 0);
    run<2>("LDTM x2 dependent chain (latency)", nw, 20000, 8, 0);
    run<3>("TMEM row update ld+16 DFMA+st (256 B/thread)", nw, 20000, 256, 32);
    run<4>("smem row update 8x(LDS.128+2 DFMA+STS.128)", nw, 20000, 256, 32);
    run<5>("DFMA peak (32 per iteration)", nw, 20000, 0, 64);
  }
  return 0;
}
