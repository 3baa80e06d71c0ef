// rxn_tm_dev.cuh — device code of the TENSOR-MEMORY resident RReact kernel (sm_100a).
//
// Why.  The resident-lane kernel (rxn_lane_dev.cuh) keeps the Newton system [J | b] of a cell (N x (N+1) doubles, 1.9 KB for
// the 300A chemistry) in shared memory: 64 cells per SM, one warp per scheduler, and every LU row update pays 16 B of
// shared-memory load + 16 B of store per DFMA pair (measured: 125 B/clk/SM = 4.5 TFLOP/s chip-wide for the row update,
// profiles/r02_ubench_tmem.txt).  Blackwell's tensor memory is 256 KB per SM of 128 lanes x 512 32-bit columns that the
// 32x32b shape of tcgen05.ld / tcgen05.st exposes as a 512-word private scratchpad per thread (lane = 32 (warp % 4) + lane id):
// exactly 16 x 16 doubles per lane, i.e. one cell's [J | b] with N <= 15, at 300-660 B/clk/SM measured.  So here
//   * a CELL is a TMEM LANE; the CTA holds 128 cells (4 quads of 32) and J never touches shared memory;
//   * a cell is worked on by G member WARPS (warp = quad + 4 member): thread (quad q, member l, lane t) is member l of
//     cell 32 q + t.  Warps q, q+4, q+8, .. address the same TMEM lanes, so every member reads and writes its cell's J
//     directly; the members meet at named barrier 1+q and exchange scalars through shared memory;
//   * every warp is UNIFORM: its 32 lanes are 32 different cells doing the same member's share of the work, so table reads
//     are broadcasts, TMEM addresses are warp-uniform (tcgen05.ld/st are .sync.aligned) and the per-cell vectors
//     (elem[e][cell], cell-fastest, as in the resident-lane kernel) are conflict free;
//   * control flow is CONVERGENT: lanes are persistent (a lane whose cell finished takes the next one from the global
//     counter), cells of one warp are in different Newton iterations / closing passes, and everything cell-specific is a
//     predicate, never a branch around a TMEM access or a barrier.  A lane without a cell keeps iterating on its last
//     (converged) cell; nothing it does is visible outside its own shared-memory column and TMEM lane.
// With the vectors at 1.5 KB per cell the SM holds 128 cells and 8 (G = 2) or 16 (G = 4) warps instead of 64 cells / 4 warps.
//
// Reference routines restated (file:line at each site): as rxn_lane_dev.cuh.  The arithmetic of every routine is that of the
// resident-lane kernel (same term streams, ln-m Jacobian, REASSOC notes of rxn_lane_dev.cuh); what differs is where J lives,
// the LU (right-looking on TMEM rows, k fully unrolled so that register rows are indexed at compile time, row swaps done by
// select on column slices because a TMEM address cannot differ between the lanes of a warp) and the predicated control flow.
//
// The same source is compiled for the host by the CPU-only test harness (RXN_TM_HOST: one "warp" = one lane = one host
// thread per member, TMEM = a plain array) and checked against the oracle in the CPU-only suite.
#pragma once
#include "rxn_lane.h"

#ifndef RXN_TM_HOST
#include <cuda_runtime.h>
#define TM_DEV static __device__ __forceinline__
#define TM_COLD static __device__ __noinline__
#else
#include <pthread.h>
#define TM_DEV static inline
#define TM_COLD static inline
#ifndef RXN_LANE_HOST
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
#endif
#endif

namespace rxn {
namespace tmk {

#ifndef RXN_LOG_TO_LN
#define RXN_LOG_TO_LN 2.30258509299           /* pflotran_constants.F90:48 (truncated on purpose) */
#define RXN_IDEAL_GAS_CONSTANT 8.31446        /* pflotran_constants.F90:53 */
#endif

#ifndef RXN_TM_HOST
extern __shared__ __align__(16) double tsm[];
#else
static double *tsm = nullptr;                 // one emulated cell at a time
static double *tmh = nullptr;                 // its tensor-memory lane: 256 doubles
struct HostGroup { pthread_barrier_t bar; };
static HostGroup *g_hg = nullptr;
#endif

#ifndef TD
#define TD(lt, o) (tsm[(o)])
#define TI(lt, o) (reinterpret_cast<const int *>(tsm)[2 * (lt).blob_dbl + (o)])
#define TI4(lt, o4) (reinterpret_cast<const int4 *>(tsm)[((lt).blob_dbl >> 1) + (o4)])    /* o4 in units of 4 ints */
#define TD2(lt, o2) (reinterpret_cast<const double2 *>(tsm)[(o2)])                          /* o2 in units of 2 doubles */
#define GSL(S, field, row, cell) ((S).f[field][(long long)(row) * (S).ld + (cell)])
#endif

constexpr int TM_LD = 16;                     // doubles per TMEM row: [J_i0 .. J_i,N-1 | b_i] padded to 16

// ---------------------------------------------------------------------------------------------
// per-thread context: registers.  R = rows of the Newton system owned by this member.
template <int N, int G>
struct Ctx {
  static constexpr int R = (N + G - 1) / G;
  int l;                  // member (warp) of the cell's group
  int s;                  // cell column within the CTA = TMEM lane
  int par;                // parity of the exchange slots
  int bar;                // named barrier of the quad
  unsigned tb;            // TMEM address of column 0 of this thread's lane quarter
  int vm, vlna, vlng, vsm, vtot, vscr, vsc, vfree, vmnrl, vr0, vseq, vlk, vres, vx;   // double index of element 0 of each slot
  double fix[R];          // fixed accumulation of the owned rows l, l+G, ... (reaction.F90:3370-3400)
  double den_kg_per_L, psv, psvd, v_t, volume, porosity, soil_density, temp, ln_act_h2o, den_kg;
  long long item, cell;
  int iter, flags;
};

template <int N, int CPB, int G>
TM_DEV void tm_bind(const LaneTab &lt, Ctx<N, G> &c, int s, int l, int quad, unsigned tmem_base) {
  c.s = s; c.l = l; c.par = 0; c.bar = 1 + quad;
  c.tb = tmem_base + ((unsigned)(quad * 32) << 16);
  const int v = lt.o_vec + s;
  c.vm = v + lt.s_m * CPB; c.vlna = v + lt.s_lna * CPB; c.vlng = v + lt.s_lng * CPB; c.vsm = v + lt.s_sm * CPB;
  c.vtot = v + lt.s_tot * CPB; c.vscr = v + lt.s_scr * CPB; c.vsc = v + lt.s_sc * CPB; c.vfree = v + lt.s_free * CPB;
  c.vmnrl = v + lt.s_mnrl * CPB; c.vr0 = v + lt.s_r0 * CPB; c.vseq = v + lt.s_seq * CPB; c.vlk = v + lt.s_lk * CPB;
  c.vres = v + lt.s_res * CPB; c.vx = v + lt.s_x * CPB;
}

// ---------------------------------------------------------------------------------------------
// tensor memory as per-thread scratch: NW consecutive 32-bit columns of this thread's lane <-> NW/2 doubles.
// The load and its tcgen05.wait::ld are ONE asm statement: the compiler must not schedule a use of the destination
// registers between them.
#ifndef RXN_TM_HOST
template <int ND> struct TmIo;
#define TM_R2(a, i) "%" #a ", %" #i
template <> struct TmIo<1> {
  TM_DEV void ld(unsigned ta, double (&d)[1]) {
    unsigned r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;\n" : "=r"(r0), "=r"(r1) : "r"(ta) : "memory");
    d[0] = __hiloint2double(r1, r0);
  }
  TM_DEV void st(unsigned ta, const double (&d)[1]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%2], {%0, %1};\n" ::"r"(__double2loint(d[0])), "r"(__double2hiint(d[0])), "r"(ta) : "memory");
  }
};
template <> struct TmIo<4> {
  TM_DEV void ld(unsigned ta, double (&d)[4]) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(ta) : "memory");
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = __hiloint2double(r[2 * j + 1], r[2 * j]);
  }
  TM_DEV void st(unsigned ta, const double (&d)[4]) {
    unsigned r[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { r[2 * j] = __double2loint(d[j]); r[2 * j + 1] = __double2hiint(d[j]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};\n"
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(ta) : "memory");
  }
};
template <> struct TmIo<8> {
  TM_DEV void ld(unsigned ta, double (&d)[8]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(ta) : "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = __hiloint2double(r[2 * j + 1], r[2 * j]);
  }
  TM_DEV void st(unsigned ta, const double (&d)[8]) {
    unsigned r[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { r[2 * j] = __double2loint(d[j]); r[2 * j + 1] = __double2hiint(d[j]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};\n"
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                   "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(ta) : "memory");
  }
};
template <> struct TmIo<16> {
  TM_DEV void ld(unsigned ta, double (&d)[16]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(ta) : "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = __hiloint2double(r[2 * j + 1], r[2 * j]);
  }
  TM_DEV void st(unsigned ta, const double (&d)[16]) {
    unsigned r[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) { r[2 * j] = __double2loint(d[j]); r[2 * j + 1] = __double2hiint(d[j]); }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};\n"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
          "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
          "r"(r[31]), "r"(ta) : "memory");
  }
};
// ND doubles starting at double index `e` of row `i`
template <int ND> TM_DEV void tm_ld(unsigned tb, int i, int e, double (&d)[ND]) { TmIo<ND>::ld(tb + (unsigned)(2 * (TM_LD * i + e)), d); }
template <int ND> TM_DEV void tm_st(unsigned tb, int i, int e, const double (&d)[ND]) { TmIo<ND>::st(tb + (unsigned)(2 * (TM_LD * i + e)), d); }
TM_DEV void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
TM_DEV bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
TM_DEV void warp_converge() { __syncwarp(); }
#else
template <int ND> TM_DEV void tm_ld(unsigned tb, int i, int e, double (&d)[ND]) {
  for (int j = 0; j < ND; ++j) d[j] = tmh[tb / 2 + TM_LD * i + e + j];
}
template <int ND> TM_DEV void tm_st(unsigned tb, int i, int e, const double (&d)[ND]) {
  for (int j = 0; j < ND; ++j) tmh[tb / 2 + TM_LD * i + e + j] = d[j];
}
TM_DEV void tm_wait_st() {}
TM_DEV bool warp_any(bool p) { return p; }
TM_DEV void warp_converge() {}
#endif
// one double at 32-bit column `col` (plan-B closers carry columns)
TM_DEV void tm_st_col(unsigned tb, int col, double v) {
#ifndef RXN_TM_HOST
  double d[1] = {v};
  TmIo<1>::st(tb + (unsigned)col, d);
#else
  tmh[(tb + col) / 2] = v;
#endif
}
TM_DEV double tm_ld_el(unsigned tb, int i, int j) { double d[1]; tm_ld<1>(tb, i, j, d); return d[0]; }
TM_DEV void tm_st_el(unsigned tb, int i, int j, double v) { double d[1] = {v}; tm_st<1>(tb, i, j, d); }

// ---------------------------------------------------------------------------------------------
// group primitives: the G member warps of a quad.  A barrier is the quad's named barrier; a reduction is one store per
// member into the exchange slots, one barrier, G loads.  The slots alternate between two sets, so a reduction needs no
// second barrier: a member can only write set p again after passing the barrier of the reduction in between, which every
// member reaches only after it has read set p.  The combination order is the xor butterfly of the resident-lane kernel.
template <int G, class C> TM_DEV void grp_sync(const C &c) {
#ifndef RXN_TM_HOST
  if (G > 1) asm volatile("bar.sync %0, %1;\n" ::"r"(c.bar), "n"(32 * G) : "memory");
  else __syncwarp();
#else
  (void)c;
  if (G > 1) pthread_barrier_wait(&g_hg->bar);
#endif
}
template <int CPB, int G, class C> TM_DEV void grp_gather2(C &c, double v0, double v1, double (&o0)[G], double (&o1)[G]) {
  if (G == 1) { o0[0] = v0; o1[0] = v1; return; }
  const int base = c.vx + c.par * (2 * G) * CPB;
  tsm[base + c.l * CPB] = v0;
  tsm[base + (G + c.l) * CPB] = v1;
  c.par ^= 1;
  grp_sync<G>(c);
#pragma unroll
  for (int g = 0; g < G; ++g) { o0[g] = tsm[base + g * CPB]; o1[g] = tsm[base + (G + g) * CPB]; }
}
template <int CPB, int G, class C> TM_DEV void grp_gather(C &c, double v, double (&o)[G]) {
  if (G == 1) { o[0] = v; return; }
  const int base = c.vx + c.par * (2 * G) * CPB;
  tsm[base + c.l * CPB] = v;
  c.par ^= 1;
  grp_sync<G>(c);
#pragma unroll
  for (int g = 0; g < G; ++g) o[g] = tsm[base + g * CPB];
}
template <int CPB, int G, class C> TM_DEV double grp_sum(C &c, double v) {
  double o[G];
  grp_gather<CPB, G>(c, v, o);
  if (G == 1) return o[0];
  if (G == 2) return o[0] + o[G - 1];
  if (G == 3) return (o[0] + o[1]) + o[G - 1];
  return (o[0] + o[2 % G]) + (o[1 % G] + o[3 % G]);
}
template <int CPB, int G, class C> TM_DEV double grp_max(C &c, double v) {
  double o[G];
  grp_gather<CPB, G>(c, v, o);
  double m = o[0];
#pragma unroll
  for (int g = 1; g < G; ++g) m = fmax(m, o[g]);
  return m;
}
template <int CPB, int G, class C> TM_DEV double grp_min(C &c, double v) {
  double o[G];
  grp_gather<CPB, G>(c, v, o);
  double m = o[0];
#pragma unroll
  for (int g = 1; g < G; ++g) m = fmin(m, o[g]);
  return m;
}
template <int CPB, int G, class C> TM_DEV bool grp_any(C &c, bool p) {
  if (G == 1) return p;
  return grp_max<CPB, G>(c, p ? 1.0 : 0.0) != 0.0;
}
// pivot of ludcmp over the group: maximum value, ties -> largest index ("last maximum")
template <int CPB, int G, class C> TM_DEV void grp_argmax_last(C &c, double &best, int &bidx) {
  if (G == 1) return;
  double ob[G], oi[G];
  grp_gather2<CPB, G>(c, best, (double)bidx, ob, oi);
  best = ob[0]; bidx = (int)oi[0];
#pragma unroll
  for (int g = 1; g < G; ++g) {
    const int i = (int)oi[g];
    if (ob[g] > best || (ob[g] == best && i > bidx)) { best = ob[g]; bidx = i; }
  }
}

#ifndef RXN_TM_HOST
TM_DEV long long tm_bits(double x) { return __double_as_longlong(x); }
TM_DEV double tm_from_bits(long long b) { return __longlong_as_double(b); }
#else
TM_DEV long long tm_bits(double x) { long long b; memcpy(&b, &x, 8); return b; }
TM_DEV double tm_from_bits(long long b) { double x; memcpy(&x, &b, 8); return x; }
#endif
TM_DEV long long tm_max_ll(long long a, long long b) { return a > b ? a : b; }

// x / d with r = 1/d precomputed (one Newton correction: the quotient the division unit returns, bar double rounding)
TM_DEV double tm_div(double x, double d, double r) {
#ifndef RXN_TM_HOST
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
#else
  (void)r;
  return x / d;
#endif
}

TM_COLD double c_exp(double x) { return exp(x); }
TM_COLD double c_log(double x) { return log(x); }
TM_COLD double c_pow_slow(double x, double y) { return pow(x, y); }
TM_DEV double c_pow(double x, double y) { return y == 1.0 ? x : c_pow_slow(x, y); }   // pow(x, 1) == x exactly

// ---------------------------------------------------------------------------------------------
// RActivityCoefficients, LAG algorithm — reaction.F90:3994-4050 (as lane_act_coefs; stores predicated by `on`)
template <int N, int CPB, int G>
TM_DEV void tm_act_coefs(const LaneTab &lt, Ctx<N, G> &c, bool on) {
  const int n = lt.n, ncplx = lt.ncplx;
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, psum = 0.0;
#pragma unroll 1
  for (int i = c.l; i < n; i += G) {
    const double mm = tsm[c.vm + i * CPB];
    p0 = fma(mm, TD(lt, lt.d_pz2 + i), p0);
    if (lt.use_act_h2o && i + 1 != lt.h2o_aq_id) psum += mm;
  }
  const int k4 = ncplx & ~3;
#pragma unroll 1
  for (int k = 4 * c.l; k < k4; k += 4 * G) {                  // REASSOC: partial sums, Z^2 premultiplied
    const double s0 = tsm[c.vsm + k * CPB], s1 = tsm[c.vsm + (k + 1) * CPB], s2 = tsm[c.vsm + (k + 2) * CPB],
                 s3 = tsm[c.vsm + (k + 3) * CPB];
    p0 = fma(s0, TD(lt, lt.d_cz2 + k), p0); p1 = fma(s1, TD(lt, lt.d_cz2 + k + 1), p1);
    p2 = fma(s2, TD(lt, lt.d_cz2 + k + 2), p2); p3 = fma(s3, TD(lt, lt.d_cz2 + k + 3), p3);
    if (lt.use_act_h2o) psum += (s0 + s1) + (s2 + s3);
  }
#pragma unroll 1
  for (int k = k4 + c.l; k < ncplx; k += G) {
    const double s0 = tsm[c.vsm + k * CPB];
    p1 = fma(s0, TD(lt, lt.d_cz2 + k), p1);
    if (lt.use_act_h2o) psum += s0;
  }
  const double I = 0.5 * grp_sum<CPB, G>(c, (p0 + p1) + (p2 + p3));
  const double sqrt_I = sqrt(I);
  if (c.l == 0 && on) tsm[c.vlng] = 0.0;
#pragma unroll 1
  for (int q = 1 + c.l; q < lt.ncls; q += G) {
    const double v = (-TD(lt, lt.d_cls_z2 + q) * sqrt_I * lt.debyeA / (1.0 + TD(lt, lt.d_cls_a0 + q) * lt.debyeB * sqrt_I) + lt.debyeBdot * I) * RXN_LOG_TO_LN;
    if (on) tsm[c.vlng + q * CPB] = v;
  }
  if (lt.use_act_h2o) {                                        // :4043-4050
    const double a = 1.0 - 0.017 * grp_sum<CPB, G>(c, psum);
    const double la = (a > 0.0) ? c_log(a) : 0.0;
    if (on) {
      c.ln_act_h2o = la;
      if (c.l == 0) tsm[c.vlna + (lt.n + 1) * CPB] = la;
    }
  }
  grp_sync<G>(c);
}

// ---------------------------------------------------------------------------------------------
// term streams (rxn_lane.h): 4 accumulators advance together, one {coef[4], offset[4]} record per step
#ifndef TM_UNROLL
#define TM_UNROLL 2
#endif
TM_DEV void tm_run_group(const LaneTab &lt, int s, int c0, int o0, int nsteps, double &a0, double &a1, double &a2, double &a3) {
  constexpr int U = TM_UNROLL;
  const char *cell8 = reinterpret_cast<const char *>(tsm + s);   // a gather address is this + the record's byte offset
#pragma unroll U
  for (int q = 0; q < nsteps; ++q) {
    const double2 ca = TD2(lt, (c0 + 1 + q) * 2), cb = TD2(lt, (c0 + 1 + q) * 2 + 1);
    const int4 of = TI4(lt, o0 + q);
    a0 = fma(ca.x, *reinterpret_cast<const double *>(cell8 + of.x), a0);
    a1 = fma(ca.y, *reinterpret_cast<const double *>(cell8 + of.y), a1);
    a2 = fma(cb.x, *reinterpret_cast<const double *>(cell8 + of.z), a2);
    a3 = fma(cb.y, *reinterpret_cast<const double *>(cell8 + of.w), a3);
  }
}

// ln a_i = ln m_i + ln gamma_i, then sec_molal_k = exp(lnQK_k - ln gamma_k)   (RTotal, reaction.F90:4090-4122)
template <int N, int CPB, int G>
TM_DEV void tm_speciate(const LaneTab &lt, Ctx<N, G> &c) {
  const int n = lt.n, s = c.s;
#pragma unroll 2
  for (int i = c.l; i < n; i += G)
    tsm[c.vlna + i * CPB] = log(tsm[c.vm + i * CPB]) + tsm[c.vlng + TI(lt, lt.i_pcls + i) * CPB];
  grp_sync<G>(c);
#pragma unroll 1
  for (int g = c.l; g < lt.spec.ng; g += G) {
    const int4 hd = TI4(lt, (lt.spec.g0 >> 2) + 2 * g), h2 = TI4(lt, (lt.spec.g0 >> 2) + 2 * g + 1);
    const int c0 = hd.x, o0 = hd.y, nsteps = hd.z, cb = h2.x >> 2;
    const int4 m0 = TI4(lt, cb), m1 = TI4(lt, cb + 1), m2 = TI4(lt, cb + 2), m3 = TI4(lt, cb + 3);
    double a0, a1, a2, a3;
    if (lt.percell_logK) {
      a0 = tsm[c.vlk + (m0.z < 0 ? 0 : m0.z) * CPB]; a1 = tsm[c.vlk + (m1.z < 0 ? 0 : m1.z) * CPB];
      a2 = tsm[c.vlk + (m2.z < 0 ? 0 : m2.z) * CPB]; a3 = tsm[c.vlk + (m3.z < 0 ? 0 : m3.z) * CPB];
    } else {
      const double2 ia = TD2(lt, c0 * 2), ib = TD2(lt, c0 * 2 + 1);
      a0 = ia.x; a1 = ia.y; a2 = ib.x; a3 = ib.y;
    }
    tm_run_group(lt, s, c0, o0, nsteps, a0, a1, a2, a3);
    // REASSOC: exp(lnQK)/gamma_k -> exp(lnQK - ln gamma_k)
    const double e0 = exp(a0 - tsm[m0.y + s]), e1 = exp(a1 - tsm[m1.y + s]), e2 = exp(a2 - tsm[m2.y + s]),
                 e3 = exp(a3 - tsm[m3.y + s]);
    tsm[m0.x + s] = e0; tsm[m1.x + s] = e1; tsm[m2.x + s] = e2; tsm[m3.x + s] = e3;
  }
  grp_sync<G>(c);
}

// plan A: tot_i <- sum_k nu_ik sm_k
template <int N, int CPB, int G>
TM_DEV void tm_planA(const LaneTab &lt, Ctx<N, G> &c) {
  const LaneStream S = lt.planA;
  const int s = c.s;
#pragma unroll 1
  for (int g = c.l; g < S.ng; g += G) {
    const int4 hd = TI4(lt, (S.g0 >> 2) + 2 * g), h2 = TI4(lt, (S.g0 >> 2) + 2 * g + 1);
    const int c0 = hd.x, o0 = hd.y, nsteps = hd.z, mode = hd.w, cb = h2.x >> 2;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    tm_run_group(lt, s, c0, o0, nsteps, a0, a1, a2, a3);
    const int4 d = TI4(lt, cb);
    if (mode == LANE_WIDE) tsm[d.x + s] = (a0 + a1) + (a2 + a3);
    else { tsm[d.x + s] = a0; tsm[d.y + s] = a1; tsm[d.z + s] = a2; tsm[d.w + s] = a3; }
  }
}

// plan B: Jln_ij = Jln_ji <- (sum_k nu_ik nu_jk sm_k) dp, diagonal + m_i dp (REASSOC: (1 + D_ii/m_i) m_i), stored to TMEM
TM_DEV void tm_planB_close(const LaneTab &lt, unsigned tb, int s, const int4 d, double a, double dp) {
  double v = a * dp;
  if (d.z >= 0) v = fma(tsm[d.z + s], dp, v);
  tm_st_col(tb, d.x, v);
  if (d.y != d.x) tm_st_col(tb, d.y, v);
}
template <int N, int CPB, int G>
TM_DEV void tm_planB(const LaneTab &lt, Ctx<N, G> &c, double dp) {
  const LaneStream S = lt.planB;
  const int s = c.s;
  warp_converge();
#pragma unroll 1
  for (int g = c.l; g < S.ng; g += G) {
    const int4 hd = TI4(lt, (S.g0 >> 2) + 2 * g), h2 = TI4(lt, (S.g0 >> 2) + 2 * g + 1);
    const int c0 = hd.x, o0 = hd.y, nsteps = hd.z, mode = hd.w, cb = h2.x >> 2;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    tm_run_group(lt, s, c0, o0, nsteps, a0, a1, a2, a3);
    if (mode == LANE_WIDE) {
      tm_planB_close(lt, c.tb, s, TI4(lt, cb), (a0 + a1) + (a2 + a3), dp);
    } else {
      tm_planB_close(lt, c.tb, s, TI4(lt, cb), a0, dp);
      tm_planB_close(lt, c.tb, s, TI4(lt, cb + 1), a1, dp);
      tm_planB_close(lt, c.tb, s, TI4(lt, cb + 2), a2, dp);
      tm_planB_close(lt, c.tb, s, TI4(lt, cb + 3), a3, dp);
    }
  }
  tm_wait_st();
}

// ---------------------------------------------------------------------------------------------
// RTotalSorbEqSurfCplx1 — reaction_surf_complex.F90:658-934, one surface complexation reaction (as lane_srf_rxn).
//   target_i (tsm[tb + i*CPB]) += total sorbed of primary i
//   addJ: Jln(i, j) += fac * d(total_sorb_i)/d ln m_j, read-modify-write on the TMEM rows of this member
//   on: this lane's stores that outlive the call (free-site warm start, per-complex concentrations) are kept
// All lanes of the warp walk the same code; the free-site iteration runs until every lane of the warp is done
// (the members of a cell see identical values, so the trip count is the same in all of them).
template <int N, int CPB, int G>
TM_DEV void tm_srf_rxn(const LaneTab &lt, Ctx<N, G> &c, const DevState &S, int irxn, double fac, bool addJ, bool store_conc, bool on,
                       int tb) {
  const int n = lt.n;
  const double tol = 1.0e-12;
  const int c0 = TI(lt, lt.i_rxn_cptr + irxn), c1 = TI(lt, lt.i_rxn_cptr + irxn + 1);
  const int nlk0 = lt.d_nlk + lt.ncplx + lt.nkin;
  double free_site_conc = tsm[c.vfree + irxn * CPB];
  double site_density;
  const int surf_type = TI(lt, lt.i_rxn_surf_type + irxn);
  const double dens = TD(lt, lt.d_rxn_density + irxn);
  if (surf_type == RXN_MINERAL_SURFACE) site_density = dens * tsm[c.vmnrl + (TI(lt, lt.i_rxn_to_surf + irxn) - 1) * CPB];
  else if (surf_type == RXN_ROCK_SURFACE) site_density = dens * c.soil_density * (1.0 - c.porosity);
  else site_density = dens;
  const bool live = !(site_density < 1.0e-40);                // :749: the reaction is skipped for this cell
  if (!live) site_density = 1.0;
  const int stoich_flag = TI(lt, lt.i_rxn_flag + irxn);
  bool one_more = false, done = false;
  int num_iterations = 0;
  double damping_factor = 1.0;
  grp_sync<G>(c);                                              // every member has read the warm-start value
#pragma unroll 1
  for (;;) {                                                  // :760-829
    num_iterations = num_iterations + 1;
    const double ln_free_site = c_log(free_site_conc);
    double part = 0.0, part2 = 0.0;
#pragma unroll 1
    for (int j = c0 + c.l; j < c1; j += G) {
      const int icplx = TI(lt, lt.i_rxn_cid + j);
      double lnQK = lt.percell_logK ? tsm[c.vlk + (lt.ncplx + lt.nkin + icplx) * CPB] : TD(lt, nlk0 + icplx);
      const double sh2o = TD(lt, lt.d_sh2o + icplx), site_st = TD(lt, lt.d_site_st + icplx);
      if (sh2o != 0.0) lnQK = lnQK + sh2o * c.ln_act_h2o;
      lnQK = lnQK + site_st * ln_free_site;
      const int p1 = TI(lt, lt.i_sptr + icplx + 1);
#pragma unroll 1
      for (int p = TI(lt, lt.i_sptr + icplx); p < p1; ++p) lnQK = lnQK + TD(lt, lt.d_sst + p) * tsm[c.vlna + TI(lt, lt.i_sid + p) * CPB];
      const double sc = c_exp(lnQK);
      if (!done) tsm[c.vsc + (j - c0) * CPB] = sc;
      part += site_st * sc;
      part2 += site_st * sc / free_site_conc;
    }
    double total = free_site_conc + grp_sum<CPB, G>(c, part);   // REASSOC: tree sum
    if (one_more) done = true;                                 // the reference leaves the loop here
    if (stoich_flag) {
      const double res = site_density - total;
      const double dres_dfree_site = 1.0 + grp_sum<CPB, G>(c, part2);
      if (!done) {
        const double dfree_site_conc = res / dres_dfree_site;
        if (num_iterations > 1000) damping_factor = 0.5;
        free_site_conc = free_site_conc + damping_factor * dfree_site_conc;
        const double rel_change = fabs(dfree_site_conc / free_site_conc);
        if (rel_change < tol) one_more = true;
        if (num_iterations > 100000) { if (on) c.flags |= RXN_FLAG_CAPPED; one_more = true; }   // reference would spin
      }
    } else if (!done) {
      total = total / free_site_conc;
      free_site_conc = site_density / total;
      one_more = true;
    }
    if (!warp_any(!done)) break;
  }
  grp_sync<G>(c);
  const bool keep = on && live;
  if (c.l == 0 && keep) tsm[c.vfree + irxn * CPB] = free_site_conc;
  if (store_conc && keep) {
#pragma unroll 1
    for (int j = c0 + c.l; j < c1; j += G) GSL(S, RXN_F_EQSRFCPLX_CONC, TI(lt, lt.i_rxn_cid + j), c.cell) += tsm[c.vsc + (j - c0) * CPB];
  }
  warp_converge();
  const double inv_free = 1.0 / free_site_conc;
  const double lv = live ? 1.0 : 0.0;                          // a skipped reaction adds exact zeros
  if (addJ) {                                                  // :838-866 (tempreal redundantly per member: few complexes)
#pragma unroll 1
    for (int row = c.l; row < n; row += G) tsm[c.vscr + row * CPB] = 0.0;
    double tempreal = 0.0;
#pragma unroll 1
    for (int j = c0; j < c1; ++j) {
      const int icplx = TI(lt, lt.i_rxn_cid + j);
      const double sc = tsm[c.vsc + (j - c0) * CPB], site_st = TD(lt, lt.d_site_st + icplx);
      const int p1 = TI(lt, lt.i_sptr + icplx + 1);
#pragma unroll 1
      for (int p = TI(lt, lt.i_sptr + icplx); p < p1; ++p) {
        const int row = TI(lt, lt.i_sid + p);
        if (G > 1 && (row % G) != c.l) continue;
        const int o = c.vscr + row * CPB;
        tsm[o] = tsm[o] + TD(lt, lt.d_sst + p) * site_st * sc;
      }
      tempreal = tempreal + site_st * site_st * sc;
    }
    tempreal = tempreal / free_site_conc;
    tempreal = tempreal + 1.0;
    {
      const double itemp = 1.0 / tempreal;                       // the quotients below are the correctly rounded ones (tm_div)
#pragma unroll 1
      for (int row = c.l; row < n; row += G) tsm[c.vscr + row * CPB] = tm_div(-tsm[c.vscr + row * CPB], tempreal, itemp);   // dSx/d ln m_row
    }
    grp_sync<G>(c);
  }
#pragma unroll 1
  for (int j = c0; j < c1; ++j) {                              // :872-931
    const int icplx = TI(lt, lt.i_rxn_cid + j);
    const double sc = tsm[c.vsc + (j - c0) * CPB];
    const int p0 = TI(lt, lt.i_sptr + icplx), p1 = TI(lt, lt.i_sptr + icplx + 1);
    const double nui_Si_over_Sx = TD(lt, lt.d_site_st + icplx) * sc * inv_free;
#pragma unroll 1
    for (int p = p0; p < p1; ++p) {
      const int row = TI(lt, lt.i_sid + p);
      if (G > 1 && (row % G) != c.l) continue;
      const double stp = TD(lt, lt.d_sst + p);
      const int o = tb + row * CPB;
      tsm[o] = tsm[o] + (stp * sc) * lv;
      if (addJ) {
#pragma unroll 1
        for (int q = p0; q < p1; ++q) {
          const int jc = TI(lt, lt.i_sid + q);
          const double tr = TD(lt, lt.d_sst + q) * sc + nui_Si_over_Sx * tsm[c.vscr + jc * CPB];
          const double cur = tm_ld_el(c.tb, row, jc);
          tm_st_el(c.tb, row, jc, cur + ((stp * tr) * fac) * lv);
          tm_wait_st();                                        // the next entry may be this one again (same pair in another complex)
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RKineticMineral — reaction_mineral.F90:564-1000 (as lane_kinetic_mineral).  Every member evaluates the (few) rate
// laws redundantly; the member that owns a primary's row adds its residual entry and its Jacobian row.  The reference's
// early exits (:723, :730) are the predicate `act`: an inactive mineral adds exact zeros.  keep: the rate is stored.
template <int N, int CPB, int G>
TM_DEV void tm_kinetic_mineral(const LaneTab &lt, Ctx<N, G> &c, bool keep) {
#pragma unroll 1
  for (int imnrl = 0; imnrl < lt.nkin; ++imnrl) {
    double lnQK = lt.percell_logK ? tsm[c.vlk + (lt.ncplx + imnrl) * CPB] : TD(lt, lt.d_nlk + lt.ncplx + imnrl);
    const double h2ost = TD(lt, lt.d_kh2o + imnrl);
    if (h2ost != 0.0) lnQK = lnQK + h2ost * c.ln_act_h2o;
    const int p0 = TI(lt, lt.i_kptr + imnrl), p1 = TI(lt, lt.i_kptr + imnrl + 1);
#pragma unroll 1
    for (int p = p0; p < p1; ++p) lnQK = lnQK + TD(lt, lt.d_kst + p) * tsm[c.vlna + TI(lt, lt.i_kid + p) * CPB];
    double QK;
    if (lnQK <= 6.90776) QK = c_exp(lnQK); else QK = 1.0e3;
    const double k_scale = lt.has_scale ? TD(lt, lt.d_k_scale + imnrl) : 1.0;
    const double k_Temkin = lt.has_Temkin ? TD(lt, lt.d_k_Temkin + imnrl) : 1.0;
    const double k_power = lt.has_power ? TD(lt, lt.d_k_power + imnrl) : 1.0;
    const double k_lim = TD(lt, lt.d_k_lim + imnrl);
    const double k_aff = TD(lt, lt.d_k_aff + imnrl);
    double affinity_factor;
    if (lt.has_Temkin) {
      if (lt.has_scale) affinity_factor = 1.0 - c_pow(QK, 1.0 / (k_scale * k_Temkin));
      else affinity_factor = 1.0 - c_pow(QK, 1.0 / k_Temkin);
    } else if (lt.has_scale) {
      affinity_factor = 1.0 - c_pow(QK, 1.0 / k_scale);
    } else {
      affinity_factor = 1.0 - QK;
    }
    const double sign_ = copysign(1.0, affinity_factor);
    const double volfrac = tsm[c.vmnrl + imnrl * CPB];
    bool act = (volfrac > 0 || sign_ < 0.0);                    // :723
    if (k_aff > 0.0 && sign_ < 0.0 && QK < k_aff) act = false;  // :730
    if (k_lim > 0.0) affinity_factor = affinity_factor / (1.0 + (1.0 - affinity_factor) / k_lim);
    double arrhenius_factor = 1.0;
    const double Ea = TD(lt, lt.d_k_Ea + imnrl);
    if (Ea > 0.0) arrhenius_factor = c_exp(Ea / RXN_IDEAL_GAS_CONSTANT * (1.0 / (25.0 + 273.15) - 1.0 / (c.temp + 273.15)));
    const double sum_prefactor_rate = TD(lt, lt.d_k_rate + imnrl) * arrhenius_factor;
    double Im_const = -tsm[c.vmnrl + (lt.nkin + imnrl) * CPB], Im;
    if (lt.has_scale) Im_const = Im_const / k_scale;
    if (lt.has_power) Im = Im_const * sign_ * c_pow(fabs(affinity_factor), k_power) * sum_prefactor_rate;
    else Im = Im_const * sign_ * fabs(affinity_factor) * sum_prefactor_rate;
    const double rate_out = act ? Im : 0.0;                    // :575 (zeroed) / :816

    Im_const = Im_const * c.volume;
    Im = Im * c.volume;
    double dIm_dQK;
    if (lt.has_power) dIm_dQK = -Im * k_power / fabs(affinity_factor);
    else dIm_dQK = -Im_const * sum_prefactor_rate;
    if (lt.has_Temkin) {
      if (lt.has_scale) dIm_dQK = dIm_dQK * (1.0 / (k_scale * k_Temkin)) / QK * (1.0 - affinity_factor);
      else dIm_dQK = dIm_dQK * (1.0 / k_Temkin) / QK * (1.0 - affinity_factor);
    } else if (lt.has_scale) {
      dIm_dQK = dIm_dQK * (1.0 / k_scale) / QK * (1.0 - affinity_factor);
    }
    const double den = (k_lim <= 0.0) ? 1.0 : 1.0 + (1.0 - affinity_factor) / k_lim;
    if (!act) { Im = 0.0; dIm_dQK = 0.0; }
    warp_converge();
#pragma unroll 1
    for (int p = p0; p < p1; ++p) {
      const int ip = TI(lt, lt.i_kid + p);
      if (G > 1 && (ip % G) != c.l) continue;            // owner member of primary ip
      const double stp = TD(lt, lt.d_kst + p);
      tsm[c.vres + ip * CPB] = tsm[c.vres + ip * CPB] + stp * Im;
#pragma unroll 1
      for (int q = p0; q < p1; ++q) {
        const int jcomp = TI(lt, lt.i_kid + q);
        const double dQK_dCj = TD(lt, lt.d_kst + q) * QK;     // d/d ln m_j: the reference's exp(-ln m_j) factor is not applied
        const double dQK_dmj = dQK_dCj * c.den_kg * 1.0e-3;
        double add;
        if (k_lim <= 0.0) add = stp * dIm_dQK * dQK_dmj;
        else add = stp * dIm_dQK * (1.0 + QK / k_lim / den) * dQK_dmj / den;
        const double cur = tm_ld_el(c.tb, ip, jcomp);
        tm_st_el(c.tb, ip, jcomp, cur + (act ? add : 0.0));
        tm_wait_st();
      }
    }
    if (c.l == 0 && keep) tsm[c.vmnrl + (2 * lt.nkin + imnrl) * CPB] = rate_out;
  }
}

// ---------------------------------------------------------------------------------------------
// RSolve (reaction.F90:4835-4880) + ludcmp/lubksb (utility.F90:393-523) on the TMEM rows [Jln_i | b_i].
// Row i is scaled and later updated by member i mod G.  Right-looking elimination, k unrolled at compile time: the pivot
// row and the row being updated are register arrays indexed by constants; per element the same a(i,j) -= a(i,k) a(k,j),
// k ascending, as Crout; pivot = last maximum of vv(i) |a(i,k)|, i >= k, found while step k-1 updates the rows; b (column N)
// goes through the elimination (= forward substitution of lubksb); row-oriented back substitution in the reference's order
// by member 0.  A row swap cannot be an address swap (TMEM addresses are warp-uniform): each member exchanges its
// column slice of rows k and imax through a select, visiting only the rows some lane of the warp pivots to
// (300A: row 4 at step 0 in 87 % of the solves, otherwise no swap at all).
// column slice [e0, e0 + W) of rows K and imax exchanged through a select; only the rows some lane of the warp pivots to are visited
template <int N, int G, int K, int W>
TM_DEV void tm_swap_slice(Ctx<N, G> &c, int e0, int imax) {
  if (e0 + W <= K) return;                                      // the whole slice is dead
  double pk[W], ri[W];
  tm_ld<W>(c.tb, K, e0, pk);
#pragma unroll 1
  for (int i = K + 1; i < N; ++i) {
    if (!warp_any(imax == i)) continue;
    tm_ld<W>(c.tb, i, e0, ri);
    const bool sel = imax == i;
#pragma unroll
    for (int e = 0; e < W; ++e) { const double a = ri[e], b = pk[e]; ri[e] = sel ? b : a; pk[e] = sel ? a : b; }
    tm_st<W>(c.tb, i, e0, ri);
  }
  tm_st<W>(c.tb, K, e0, pk);
}

template <int N, int CPB, int G, int K>
TM_DEV void tm_lu_step(Ctx<N, G> &c, double &best, int &imax) {
  constexpr int HALF = (K >= 8) ? 1 : 0;
  constexpr int E0 = HALF ? 8 : 0, ND = HALF ? 8 : 16;
  const double tiny = 1.0e-20;
  grp_argmax_last<CPB, G>(c, best, imax);                       // also orders step K-1's row updates before the loads below
  if (imax < 0) imax = K;
  if (warp_any(imax != K)) {
    // swap rows K and imax from column K on (columns left of it are never read again): member l takes the column slice
    // [e0, e0 + W) of both rows; slices of 16/G doubles, for G = 3: 8, 4, 4 (tcgen05 shapes are powers of two)
    if (G == 3) {
      if (c.l == 0) tm_swap_slice<N, G, K, 8>(c, 0, imax);
      else tm_swap_slice<N, G, K, 4>(c, 4 + 4 * c.l, imax);
    } else {
      tm_swap_slice<N, G, K, TM_LD / (G == 3 ? 4 : G)>(c, c.l * (TM_LD / (G == 3 ? 4 : G)), imax);
    }
    tm_wait_st();
    if (c.l == 0 && imax != K) tsm[c.vscr + imax * CPB] = tsm[c.vscr + K * CPB];
    warp_converge();
    grp_sync<G>(c);
  }
  double p[ND];
  tm_ld<ND>(c.tb, K, E0, p);
  double piv = p[K - E0];
  if (piv == 0.0) piv = tiny;
  const double dum = 1.0 / piv;
  if (c.l == 0) tsm[c.vscr + K * CPB] = dum;                   // vv(K) is dead: keep 1/a(K,K) for the back substitution
  best = -1.0; imax = -1;
  {
    int i1 = c.l;                                               // first owned row > K
    if (i1 <= K) i1 += ((K - i1) / G + 1) * G;
#pragma unroll 1
    for (int i = i1; i < N; i += G) {
      double r[ND];
      tm_ld<ND>(c.tb, i, E0, r);
      const double lik = r[K - E0] * dum;
#pragma unroll
      for (int j = K + 1; j <= N; ++j) r[j - E0] = r[j - E0] - lik * p[j - E0];
      tm_st<ND>(c.tb, i, E0, r);
      // pivot search of the next step (ludcmp :440-449) on the fly: rows > K are exactly the candidates of step K+1
      if (K + 1 < N) {
        const double cand = tsm[c.vscr + i * CPB] * fabs(r[K + 1 - E0]);
        if (cand >= best) { best = cand; imax = i; }
      }
    }
    tm_wait_st();
  }
}
template <int N, int CPB, int G, int K>
TM_DEV void tm_lu_steps(Ctx<N, G> &c, double &best, int &imax) {
  if constexpr (K < N) {
    tm_lu_step<N, CPB, G, K>(c, best, imax);
    tm_lu_steps<N, CPB, G, K + 1>(c, best, imax);
  }
}
template <int N, int CPB, int G, int I>
TM_DEV void tm_backsub(Ctx<N, G> &c, double (&x)[N]) {          // lubksb :511-520 (REASSOC: times 1/a(i,i))
  if constexpr (I >= 0) {
    constexpr int HALF = (I >= 8) ? 1 : 0;
    constexpr int E0 = HALF ? 8 : 0, ND = HALF ? 8 : 16;
    double r[ND];
    tm_ld<ND>(c.tb, I, E0, r);
    double sum = r[N - E0];
#pragma unroll
    for (int j = I + 1; j < N; ++j) sum = sum - r[j - E0] * x[j];
    x[I] = sum * tsm[c.vscr + I * CPB];
    tm_backsub<N, CPB, G, I - 1>(c, x);
  }
}

// returns true for a lane whose matrix has an all-zero row (reference: MPI_Abort); the solution replaces tsm[vres + i]
template <int N, int CPB, int G>
TM_DEV bool tm_rsolve(const LaneTab &lt, Ctx<N, G> &c) {
  const bool use_log = lt.use_log != 0;
  bool zero = false;
  double best = -1.0;                                          // running pivot search: value / row of the next step
  warp_converge();
  int imax = -1;
  // rows scaled by 1/max(1, max_j |J_ij|), J_ij = Jln_ij/m_j (:4851-4858), log form: times m_j (:4866-4870);
  // implicit-scaling factors vv(i) = 1/max_j |a(i,j)| (ludcmp :413-425) -> scratch
  {
    double invm[N];
#pragma unroll
    for (int j = 0; j < N; ++j) invm[j] = 1.0 / tsm[c.vm + j * CPB];
#pragma unroll 1
    for (int i = c.l; i < N; i += G) {
      double r[TM_LD];
      tm_ld<TM_LD>(c.tb, i, 0, r);
      // maxima of non-negative values: their IEEE bit patterns order like integers (a 64-bit integer max is 4 instructions, fmax
      // on doubles 7); = if (v > mx) mx = v for the finite values the solve sees
      long long mxi = 0, mri = 0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double av = fabs(r[j]), v = av * invm[j];
        mxi = tm_max_ll(mxi, tm_bits(v));
        mri = tm_max_ll(mri, tm_bits(av));
      }
      const double mx = tm_from_bits(mxi), mraw = tm_from_bits(mri);
      const double norm = 1.0 / ((mx > 1.0) ? mx : 1.0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double v = r[j];
        if (!use_log) v = v * invm[j];
        r[j] = v * norm;
      }
      r[N] = tsm[c.vres + i * CPB] * norm;
#pragma unroll
      for (int j = N + 1; j < TM_LD; ++j) r[j] = 0.0;
      tm_st<TM_LD>(c.tb, i, 0, r);
      // max_j |a(i,j)|: rounding is monotone, so in the log form it is |.|max of the unscaled row times norm
      const double aamax = use_log ? mraw * norm : mx * norm;
      if (aamax <= 0.0) zero = true;
      const double vvi = 1.0 / aamax;
      tsm[c.vscr + i * CPB] = vvi;
      const double cand = vvi * fabs(r[0]);                     // pivot search of step 0 (ludcmp :440-449)
      if (cand >= best) { best = cand; imax = i; }
    }
    tm_wait_st();
  }
  zero = grp_any<CPB, G>(c, zero);
  tm_lu_steps<N, CPB, G, 0>(c, best, imax);
  grp_sync<G>(c);
  if (c.l == 0) {
    double x[N];
    tm_backsub<N, CPB, G, N - 1>(c, x);
#pragma unroll
    for (int i = 0; i < N; ++i) tsm[c.vres + i * CPB] = x[i];
  }
  grp_sync<G>(c);
  return zero;
}

// ---------------------------------------------------------------------------------------------
// lane life cycle: load a cell -> trips (one Newton iteration each) -> finish (closing RTAuxVarCompute + write back)

// x / d with r = RN(1/d): RN(q + fma(-d, q, x) r), q = RN(x r), is the correctly rounded quotient (Markstein)
TM_DEV double hpt_div(double x, double d, double r) {
#ifndef RXN_TM_HOST
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
#else
  (void)r;
  return x / d;
#endif
}
// RUpdateTempDependentCoefs reaction.F90:5433-5524 -> -logK*LOG_TO_LN of this cell's T (and P)
template <int CPB, int G>
TM_COLD void tm_percell_logK(int l, int ncoef, int logK_mode, int vlk, double temp, double pres, const double *blob_d, DSpec s0,
                             DSpec s1, DSpec s2) {
  const double tk = temp + 273.15;
  // the cell's T, P terms of the hpt fit, once per cell (the same values the reference forms inside every evaluation)
  const double tr = tk / 273.15, pr = pres / 1.0e7;
  double logtr = 0.0, sqtr = 0.0, itr = 0.0, ipr = 0.0;
  if (logK_mode == RXN_LOGK_HPT) { logtr = log(tr) / log(10.0); sqtr = sqrt(tr); itr = 1.0 / tr; ipr = 1.0 / pr; }
  int o0 = 0;
#pragma unroll 1
  for (int q = 0; q < 3; ++q) {
    const DSpec sp = q == 0 ? s0 : q == 1 ? s1 : s2;
#pragma unroll 1
    for (int r = l; r < sp.n; r += G) {
      double lk;
      const bool fixed = sp.o_coef < 0 || (q == 2 && logK_mode == RXN_LOGK_HPT);   // :5517-5521: hpt not applied to srfcplx
      if (fixed) lk = blob_d[sp.o_logK + r];
      else {
        const double *cf = blob_d + sp.o_coef + r * ncoef;
        if (logK_mode == RXN_LOGK_HPT) {                      // reaction_aux.F90:1529-1571
          // the reference's expression term by term; divisions by tr / pr as Markstein-corrected products with the reciprocal
          // (hpt_div: the correctly rounded quotient, i.e. the division's own bits - the fit's terms cancel, so none may change)
          lk = cf[0] + cf[1] * tr + hpt_div(cf[2], tr, itr) + cf[3] * logtr + cf[4] * tr * tr + hpt_div(hpt_div(cf[5], tr, itr), tr, itr) +
               cf[6] * sqtr + cf[7] * pr + cf[8] * pr * tr + hpt_div(cf[9] * pr, tr, itr) + cf[10] * pr * logtr + hpt_div(cf[11], pr, ipr) +
               hpt_div(cf[12], pr, ipr) * tr + hpt_div(hpt_div(cf[13], pr, ipr), tr, itr) + cf[14] * pr * pr + cf[15] * pr * pr * tr +
               hpt_div(cf[16] * pr * pr, tr, itr);
        } else {                                              // reaction_aux.F90:1461-1488
          lk = cf[0] * log(tk) + cf[1] + cf[2] * tk + cf[3] / tk + cf[4] / (tk * tk);
        }
      }
      tsm[vlk + (o0 + r) * CPB] = -lk * RXN_LOG_TO_LN;
    }
    o0 += sp.n;
  }
}

// benign content for a column no cell has used yet: the lane runs the trips on it until it gets a cell
template <int N, int CPB, int G>
TM_DEV void tm_init_column(const LaneTab &lt, Ctx<N, G> &c) {
  const int nel = lt.s_x + 4 * G;                              // all vector slots
#pragma unroll 1
  for (int e = c.l; e < nel; e += G) tsm[lt.o_vec + e * CPB + c.s] = 0.0;
  grp_sync<G>(c);
#pragma unroll 1
  for (int i = c.l; i < N; i += G) tsm[c.vm + i * CPB] = 1.0;
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) tsm[c.vfree + q * CPB] = 1.0e-9;
#pragma unroll
  for (int r = 0; r < Ctx<N, G>::R; ++r) c.fix[r] = 1.0;
  c.den_kg = 1000.0; c.den_kg_per_L = 1.0; c.psv = 1.0; c.psvd = 1.0; c.v_t = 1.0; c.volume = 1.0; c.porosity = 0.5;
  c.soil_density = 1.0; c.temp = 25.0; c.ln_act_h2o = 0.0; c.item = -1; c.cell = 0; c.iter = 0; c.flags = 0;
  grp_sync<G>(c);
}

// everything of a new cell that lives on chip (lanes with `fresh`; no barrier and no TMEM access inside)
template <int N, int CPB, int G>
TM_DEV void tm_load(const LaneTab &lt, Ctx<N, G> &c, const DevState &S, const double *blob_d, const int *blob_i, const DevTab &h,
                    long long item, long long cell, const double *tran_xx, double tran_dt) {
  constexpr int R = Ctx<N, G>::R;
  const int n = lt.n;
  c.item = item; c.cell = cell;
  c.flags = 0; c.iter = 0;
  c.ln_act_h2o = GSL(S, RXN_F_LN_ACT_H2O, 0, cell);
  c.den_kg = GSL(S, RXN_F_DEN_KG, 0, cell);
  c.temp = GSL(S, RXN_F_TEMP, 0, cell);
  c.volume = GSL(S, RXN_F_VOLUME, 0, cell);
  c.porosity = GSL(S, RXN_F_POROSITY, 0, cell);
  c.soil_density = GSL(S, RXN_F_SOIL_PARTICLE_DENSITY, 0, cell);
  const double sat = GSL(S, RXN_F_SAT, 0, cell);
  c.psv = c.porosity * sat * 1000.0 * c.volume;
  c.psvd = c.porosity * sat * 1000.0 * c.volume / tran_dt;                     // :5189
  c.v_t = c.volume / tran_dt;                                                  // :4590
  c.den_kg_per_L = c.den_kg * 1.0 * 1.0e-3;
  {
    double pm[R], xx[R], ts[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = c.l + r * G, ic = i < n ? i : n - 1;
      pm[r] = GSL(S, RXN_F_PRI_MOLAL, ic, cell);
      xx[r] = tran_xx[item * n + ic];
      ts[r] = lt.neqsorb > 0 ? GSL(S, RXN_F_TOTAL_SORB_EQ, ic, cell) : 0.0;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = c.l + r * G;
      if (i < n) {
        tsm[c.vm + i * CPB] = pm[r];
        double fx = c.psv * xx[r];                               // :3370, RTAccumulation :5072-5148
        if (lt.neqsorb > 0) fx = fx + ts[r] * c.volume;          // RAccumulationSorb :4539-4568
        c.fix[r] = fx;
      } else {
        // padding row of the shape: m = 1, no complexes -> total = den, residual = psv*den - fix = 0 exactly,
        // Jln_ii = den*psvd: the row stays decoupled and its Newton update is 0
        if (i < N) tsm[c.vm + i * CPB] = 1.0;
        c.fix[r] = c.psv * ((1.0 + 0.0) * c.den_kg_per_L);
      }
    }
  }
  if (c.l == 0) {
    tsm[c.vlna + n * CPB] = 0.0;
    tsm[c.vlna + (n + 1) * CPB] = c.ln_act_h2o;
    tsm[c.vsm + lt.ncplx * CPB] = 0.0;
  }
  if (lt.act_off) {                                            // ln gamma from the state, one class per species
#pragma unroll 1
    for (int i = c.l; i < n; i += G) tsm[c.vlng + i * CPB] = c_log(GSL(S, RXN_F_PRI_ACT_COEF, i, cell));
#pragma unroll 1
    for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vlng + (n + k) * CPB] = c_log(GSL(S, RXN_F_SEC_ACT_COEF, k, cell));
  } else {
    // lagged sec_molal (for the ionic strength): straight from HBM into the cell's column, all copies in flight at once
#ifndef RXN_TM_HOST
#pragma unroll 4
    for (int k = c.l; k < lt.ncplx; k += G)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(&tsm[c.vsm + k * CPB])),
                   "l"(&GSL(S, RXN_F_SEC_MOLAL, k, cell)) : "memory");
#else
    for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vsm + k * CPB] = GSL(S, RXN_F_SEC_MOLAL, k, cell);
#endif
  }
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) tsm[c.vfree + q * CPB] = GSL(S, RXN_F_FREE_SITE_CONC, q, cell);
#pragma unroll 1
  for (int q = c.l; q < lt.nkin; q += G) {                     // read-only inside RReact
    tsm[c.vmnrl + q * CPB] = GSL(S, RXN_F_MNRL_VOLFRAC, q, cell);
    tsm[c.vmnrl + (lt.nkin + q) * CPB] = GSL(S, RXN_F_MNRL_AREA, q, cell);
  }
  if (lt.percell_logK)
    tm_percell_logK<CPB, G>(c.l, lt.ncoef, lt.logK_mode, c.vlk, c.temp, GSL(S, RXN_F_PRES, 0, cell), blob_d, h.cplx, h.kin, h.srf);
#ifndef RXN_TM_HOST
  asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// k_r / (1 + k_r dt) of rate r of multirate reaction ikr: the same for every cell of a launch, so the CTA forms the table once
// (kk, TM_MR_KK entries in shared memory; NULL: more rates than the table holds, evaluated in place - same rounding either way)
#ifndef TM_MR_KK
#define TM_MR_KK 100
#endif
TM_DEV double tm_mr_kk(const double *kk, const double *blob_d, const DevTab &h, int ikr, int irate, double tran_dt) {
  if (kk) return kk[ikr * h.mr_ld + irate];
  const double rate = blob_d[h.o_mr_rate + ikr * h.mr_ld + irate];
  return rate / (1.0 + rate * tran_dt);
}
TM_DEV void tm_mr_kk_fill(double *kk, int first, int stride, const double *blob_d, const DevTab &h, double tran_dt) {
  for (int w = first; w < h.nmr * h.mr_ld && w < TM_MR_KK; w += stride) {
    const double rate = blob_d[h.o_mr_rate + w];
    kk[w] = rate / (1.0 + rate * tran_dt);
  }
}

// multirate sorption of a cell just taken: R0_i = sum_r k_r/(1+k_r dt) S_r,i (multirate_prepare, rxn_device.cuh; REASSOC: even
// and odd rates summed separately, then added, as lane_coop_in_mr).  750 doubles per 300A cell: every lane of the warp loads
// for ONE cell at a time - lane (ii, half) sums the even (half 0) or odd (half 1) rates of row l + G ii, all its loads in
// flight before the first add - and the sum lands in that cell's column.  W = lanes of the warp (host: 1).
template <int N, int CPB, int G>
TM_DEV void tm_coop_in_mr(const LaneTab &lt, const DevState &S, const DevTab &h, const double *blob_d, const int *blob_i, int l, int slot,
                          long long cell, double tran_dt, int w, int W, const double *kk) {
  const int vr0 = lt.o_vec + lt.s_r0 * CPB + slot;
  const int n = lt.n;
  const int H = W >= 32 ? 2 : 1, per = W / H;
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr) {
    const int nrate = blob_i[h.o_mr_nrate + ikr];
    const long long row0 = ((long long)ikr * (h.mr_ld + 1) + 1) * n;
#pragma unroll 1
    for (int i0 = l; i0 < n; i0 += G * per) {
      const int i = i0 + G * (w % per), half = w / per;
      double acc0 = 0.0, acc1 = 0.0;
      if (i < n) {
        if (H == 2) {
#pragma unroll 1
          for (int r0 = half; r0 < nrate; r0 += 16) {           // 8 loads of this lane in flight
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int irate = r0 + 2 * u;
              v[u] = irate < nrate ? GSL(S, RXN_F_KINMR_TOTAL_SORB, row0 + (long long)irate * n + i, cell) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int irate = r0 + 2 * u;
              if (irate < nrate) {
                acc0 = acc0 + tm_mr_kk(kk, blob_d, h, ikr, irate, tran_dt) * v[u];
              }
            }
          }
        } else {
#pragma unroll 1
          for (int irate = 0; irate < nrate; irate += 2) {
            acc0 = acc0 + tm_mr_kk(kk, blob_d, h, ikr, irate, tran_dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, row0 + (long long)irate * n + i, cell);
            if (irate + 1 < nrate)
              acc1 = acc1 + tm_mr_kk(kk, blob_d, h, ikr, irate + 1, tran_dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, row0 + (long long)(irate + 1) * n + i, cell);
          }
        }
      }
#ifndef RXN_TM_HOST
      if (H == 2) acc1 = __shfl_xor_sync(0xffffffffu, acc0, per);   // the odd-rate sum of lane (ii, 1)
#endif
      if (i < n && half == 0) tsm[vr0 + (ikr * N + i) * CPB] = acc0 + acc1;
    }
  }
}

// One trip of the warp through the Newton loop of RReact (reaction.F90:3411-3500).  Per lane: `run` - a normal
// iteration; `closing` - the shortened last pass that redoes RTotal for the closing RTAuxVarCompute (:3507) after an
// abnormal exit changed pri_molal; neither - the lane has no cell and iterates on stale data.  exit_code: 0 to continue,
// else the exit reason / flag of a `run` lane; `recompute` is set when the closing RTAuxVarCompute needs a closing pass.
template <int N, int CPB, int G>
TM_DEV void tm_trip(const LaneTab &lt, Ctx<N, G> &c, const DevState &S, double tran_dt, double inv_dt, int dt_mode, bool run,
                    bool closing, int &exit_code, bool &recompute) {
  constexpr int R = Ctx<N, G>::R;
  const int n = lt.n;
  exit_code = 0; recompute = false;
  if (run) c.iter = c.iter + 1;
  warp_converge();
  // next Newton matrix starts from zero (ordered before plan B by the barriers of the speciation)
  {
    double z[TM_LD];
#pragma unroll
    for (int j = 0; j < TM_LD; ++j) z[j] = 0.0;
#pragma unroll 1
    for (int i = c.l; i < N; i += G) tm_st<TM_LD>(c.tb, i, 0, z);
    tm_wait_st();
  }
  // :3407-3409 (once, before the loop) and :3413-3418 (every iteration): the call before the loop and
  // the call of iteration 1 see identical inputs, so one evaluation serves both
  if (!lt.act_off) {
    const bool need = run && (c.iter == 1 || lt.act_newton_iter);
    if (warp_any(need)) tm_act_coefs<N, CPB, G>(lt, c, need);
  }
  // RTAuxVarCompute :3419 -> RTotal + RTotalSorb
  tm_speciate<N, CPB, G>(lt, c);
  tm_planA<N, CPB, G>(lt, c);
  const double dp = c.den_kg_per_L * c.psvd;                   // dtotal * psvd_t  (:3429-3437; RTAccumulationDerivative :5189-5204)
  tm_planB<N, CPB, G>(lt, c, dp);
#pragma unroll 1
  for (int i = c.l; i < N; i += G) tsm[c.vres + i * CPB] = 0.0;
  grp_sync<G>(c);
  // sorption: equilibrium reactions (RTotalSorb :4182-4216, sorbed totals -> res) and the equilibrium part of the
  // multirate reactions (RMultiRateSorption reaction_surf_complex.F90:566-654, S_eq -> its own vector), one call site.
  // REASSOC: the multirate derivative block enters J before the mineral block.
#pragma unroll 1
  for (int task = 0; task < lt.neq + lt.nmr; ++task) {
    const bool eq = task < lt.neq;
    const int ikr = task - lt.neq;
    int tb = c.vres;
    double fac = c.v_t;
    if (!eq) {
      tb = closing ? c.vres : c.vseq + ikr * N * CPB;           // a closing pass leaves S_eq as the last iteration formed it
      fac = c.volume * lt.mrK1[ikr];
#pragma unroll 1
      for (int i = c.l; i < n; i += G) tsm[tb + i * CPB] = 0.0;
    }
    tm_srf_rxn<N, CPB, G>(lt, c, S, TI(lt, (eq ? lt.i_eq_rxn : lt.i_mr_rxn - lt.neq) + task), fac, true, false, !closing, tb);
  }
  const bool consistent = dt_mode == RXN_DT_CONSISTENT;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = c.l + r * G;
    if (i < N) {
      const double tot = (tsm[c.vm + i * CPB] + tsm[c.vtot + i * CPB]) * c.den_kg_per_L;   // :4095,4124,4148
      tsm[c.vtot + i * CPB] = tot;
      double res = c.psv * tot;
      res = res - c.fix[r];                                    // :3424-3426
      if (lt.neqsorb > 0) res = res + tsm[c.vres + i * CPB] * c.volume;
      if (consistent) res = tm_div(res, tran_dt, inv_dt);
      tsm[c.vres + i * CPB] = res;
    }
  }
  // RReaction :3440 (minerals, then multirate)
  if (lt.nkin > 0) tm_kinetic_mineral<N, CPB, G>(lt, c, !closing);
#pragma unroll 1
  for (int ikr = 0; ikr < lt.nmr; ++ikr) {
#pragma unroll 1
    for (int i = c.l; i < n; i += G)
      tsm[c.vres + i * CPB] += c.volume * (lt.mrK1[ikr] * tsm[c.vseq + (ikr * N + i) * CPB] - tsm[c.vr0 + (ikr * N + i) * CPB]);
  }
  double mx = 0.0;
  bool bad = false;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (c.l + r * G < N) {
      const double v = tsm[c.vres + (c.l + r * G) * CPB];
      mx = fmax(mx, fabs(v));
      if (!isfinite(v)) bad = true;
    }
  }
  {
    double ob[G], om[G];
    grp_gather2<CPB, G>(c, bad ? 1.0 : 0.0, mx, ob, om);
    if (G > 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) { bad = bad || ob[g] != 0.0; mx = fmax(mx, om[g]); }
    }
  }
  if (bad) { exit_code = RXN_FLAG_NONFINITE; recompute = true; }
  else if (mx < lt.res_tol) exit_code = RXN_EXIT_RESIDUAL;     // :3443
  tm_wait_st();
  grp_sync<G>(c);
  if (tm_rsolve<N, CPB, G>(lt, c) && exit_code == 0) { exit_code = RXN_FLAG_LU_ZERO_ROW; recompute = true; }
  double maxrel = 0.0, min_ratio = 1.0e20;
  if (!lt.use_log) {                                           // :3459-3471
#pragma unroll 1
    for (int i = c.l; i < n; i += G) {
      const double prev = tsm[c.vm + i * CPB], u = tsm[c.vres + i * CPB];
      if (prev <= u) {
        const double ratio = fabs(prev / u);
        if (ratio < min_ratio) min_ratio = ratio;
      }
    }
    min_ratio = grp_min<CPB, G>(c, min_ratio);
  }
  // the new solution is staged in res: it is discarded when the relative change has converged (:3476)
  bad = false;
#pragma unroll 2
  for (int i = c.l; i < n; i += G) {
    double u = tsm[c.vres + i * CPB];
    const double prev = tsm[c.vm + i * CPB];
    double nw;
    if (lt.use_log) {                                          // :3454-3458
      u = copysign(1.0, u) * fmin(fabs(u), lt.max_dlnC);
      nw = prev * exp(-u);
    } else {
      if (min_ratio < 1.0) u = u * min_ratio * 0.99;
      nw = prev - u;
    }
    const double rc = fabs((nw - prev) / prev);
    if (!isfinite(rc)) bad = true;
    maxrel = fmax(maxrel, rc);
    if (c.iter > 50) nw = 0.1 * (nw - prev) + prev;            // :3478-3496
    tsm[c.vres + i * CPB] = nw;
  }
  {
    double ob[G], om[G];
    grp_gather2<CPB, G>(c, bad ? 1.0 : 0.0, maxrel, ob, om);
    if (G > 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) { bad = bad || ob[g] != 0.0; maxrel = fmax(maxrel, om[g]); }
    }
  }
  if (exit_code == 0) {
    if (bad) { exit_code = RXN_FLAG_NONFINITE; recompute = true; }
    else if (maxrel < lt.rel_tol) exit_code = RXN_EXIT_REL_CHANGE;   // :3476 (update discarded)
  }
  const bool advance = exit_code == 0 && !closing;
  if (advance) {
#pragma unroll 4
    for (int i = c.l; i < n; i += G) tsm[c.vm + i * CPB] = tsm[c.vres + i * CPB];   // :3498
  }
  warp_converge();
  if (exit_code == 0 && c.iter >= lt.maxit) { exit_code = RXN_FLAG_CAPPED; recompute = true; }   // GPU-only guard (reference spins)
  grp_sync<G>(c);
}

// closing RTAuxVarCompute (:3507) + write back.  After a normal exit pri_molal and the activity coefficients are those of
// the last RTotal, so sec_molal and total are already final; only RTotalSorb sees a different input (the warm-start
// free-site concentration).  The sorption pass is walked by every lane of the warp (`fin` lanes keep its results), the
// stores to global memory are done by the `fin` lanes alone.
template <int N, int CPB, int G>
TM_DEV void tm_finish(const LaneTab &lt, Ctx<N, G> &c, const DevState &S, const DevTab &h, double *tran_xx, int32_t *iters, int32_t *flags,
                      bool fin, int status) {
  const int n = lt.n;
  const long long cell = c.cell;
  grp_sync<G>(c);
  if (lt.neqsorb > 0) {
#pragma unroll 1
    for (int i = c.l; i < n; i += G) tsm[c.vres + i * CPB] = 0.0;
    if (lt.neq > 0 && fin) {                                   // RZeroSorb :4162-4178
#pragma unroll 1
      for (int k = c.l; k < lt.nsrf; k += G) GSL(S, RXN_F_EQSRFCPLX_CONC, k, cell) = 0.0;
    }
    warp_converge();
#pragma unroll 1
    for (int ieq = 0; ieq < lt.neq; ++ieq)
      tm_srf_rxn<N, CPB, G>(lt, c, S, TI(lt, lt.i_eq_rxn + ieq), 0.0, false, true, fin, c.vres);
    grp_sync<G>(c);
  }
  if (!lt.act_off) {                                           // gamma per class (the ln gamma of a finished cell are dead)
#pragma unroll 1
    for (int q = c.l; q < lt.ncls; q += G) {
      const double e = c_exp(tsm[c.vlng + q * CPB]);
      if (fin) tsm[c.vlng + q * CPB] = e;
    }
    warp_converge();
    grp_sync<G>(c);
  }
  if (fin) {
#pragma unroll 2
    for (int i = c.l; i < n; i += G) {
      const double mm = tsm[c.vm + i * CPB];
      tran_xx[c.item * n + i] = mm;
      GSL(S, RXN_F_PRI_MOLAL, i, cell) = mm;
      GSL(S, RXN_F_TOTAL, i, cell) = tsm[c.vtot + i * CPB];
      if (lt.neqsorb > 0) GSL(S, RXN_F_TOTAL_SORB_EQ, i, cell) = tsm[c.vres + i * CPB];
    }
#pragma unroll 1
    for (int ikr = 0; ikr < lt.nmr; ++ikr)
#pragma unroll 1
      for (int i = c.l; i < n; i += G)
        GSL(S, RXN_F_KINMR_TOTAL_SORB, (long long)ikr * (h.mr_ld + 1) * n + i, cell) = tsm[c.vseq + (ikr * N + i) * CPB];
    if (!lt.act_off) {
#pragma unroll 1
      for (int i = c.l; i < n; i += G) GSL(S, RXN_F_PRI_ACT_COEF, i, cell) = tsm[c.vlng + TI(lt, lt.i_pcls + i) * CPB];
    }
#pragma unroll 2
    for (int k = c.l; k < lt.ncplx; k += G) {
      GSL(S, RXN_F_SEC_MOLAL, k, cell) = tsm[c.vsm + k * CPB];
      if (!lt.act_off) GSL(S, RXN_F_SEC_ACT_COEF, k, cell) = tsm[c.vlng + TI(lt, lt.i_ccls + k) * CPB];
    }
#pragma unroll 1
    for (int q = c.l; q < lt.nrxn; q += G) GSL(S, RXN_F_FREE_SITE_CONC, q, cell) = tsm[c.vfree + q * CPB];
#pragma unroll 1
    for (int q = c.l; q < lt.nkin; q += G) GSL(S, RXN_F_MNRL_RATE, q, cell) = tsm[c.vmnrl + (2 * lt.nkin + q) * CPB];
    if (c.l == 0) {
      GSL(S, RXN_F_LN_ACT_H2O, 0, cell) = c.ln_act_h2o;
      if (iters) iters[c.item] = c.iter;
      if (flags) flags[c.item] = status | c.flags;
    }
  }
  warp_converge();
  grp_sync<G>(c);
}

// ---------------------------------------------------------------------------------------------
// Global-implicit cell loops on the same layout (cell = TMEM lane, G member warps, uniform warps): ONE pass per cell, no
// Newton loop.  Three entry points share it (GiArgs::mode):
//   GI_AUX  RTUpdateAuxVars cells part (reactive_transport.F90:3790-3846) [+ RActivityCoefficients, :3620-3700] and
//           RTUpdateFixedAccumulation (:786-843): RTAuxVarCompute = RTotal + RTotalSorb from a new free-ion iterate; optional
//           dtotal / dtotal_sorb_eq blocks for the flux side (DTOTAL materialised), optional accumulation output
//   GI_RJ   accumulation + reaction parts of RTResidualNonFlux (:2545-2586, 2735-2758) and RTJacobianNonFlux (:3342-3389,
//           3445-3465), as lane_gi_cell (rxn_lane_dev.cuh) / cell_residual_jacobian (rxn_device.cuh)
// Activity coefficients: `update_act` -> LAG Debye-Hueckel per class (tm_act_coefs) from the lagged sec_molal, written back
// per species; else the state's per-species values are used: ln a_i = ln m_i + ln gamma_i in the cell's column, a complex
// reads its own gamma_k from HBM at the one place it is needed (sec_molal_k = exp(lnQK_k) / gamma_k, reaction.F90:4112) -
// consecutive lanes are consecutive cells, so that read is coalesced and no per-species slot is needed on chip.
// Every lane of a warp walks the same code; `on` = this lane holds a cell of the batch (stores and flags are predicated).
// (GiArgs, GI_AUX, GI_RJ: rxn_lane.h)

// Jln rows of this member -> a cell-fastest SoA block (DTOTAL / DTOTAL_SORB_EQ: element (i, j) at row j*n + i): divided by m_j
template <int N, int CPB, int G>
TM_DEV void tm_gi_store_block(const LaneTab &lt, Ctx<N, G> &c, const DevState &S, int field, bool on) {
  const int n = lt.n;
  double invm[N];
#pragma unroll
  for (int j = 0; j < N; ++j) invm[j] = 1.0 / tsm[c.vm + j * CPB];
#pragma unroll 1
  for (int i = c.l; i < n; i += G) {
    double r[TM_LD];
    tm_ld<TM_LD>(c.tb, i, 0, r);
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < n && on) GSL(S, field, j * n + i, c.cell) = r[j] * invm[j];
  }
}
template <int N, int CPB, int G>
TM_DEV void tm_gi_zero_J(Ctx<N, G> &c) {
  double z[TM_LD];
#pragma unroll
  for (int j = 0; j < TM_LD; ++j) z[j] = 0.0;
#pragma unroll 1
  for (int i = c.l; i < N; i += G) tm_st<TM_LD>(c.tb, i, 0, z);
  tm_wait_st();
}

template <int N, int CPB, int G>
TM_DEV void tm_gi_cell(const LaneTab &lt, Ctx<N, G> &c, const DevState &S, const DevTab &h, const double *blob_d, const int *blob_i,
                       const GiArgs &a, long long item, long long cell, bool on, const double *kk) {
  const int n = lt.n, s = c.s;
  const bool rj = a.mode == GI_RJ;
  const bool from_state = !lt.act_off && !a.update_act;        // per-species gamma of the state with the class-based plan
  const double dt = a.dt;
  c.item = item; c.cell = cell; c.flags = 0; c.iter = 0;
  c.ln_act_h2o = GSL(S, RXN_F_LN_ACT_H2O, 0, cell);
  c.den_kg = GSL(S, RXN_F_DEN_KG, 0, cell);
  c.temp = GSL(S, RXN_F_TEMP, 0, cell);
  c.volume = GSL(S, RXN_F_VOLUME, 0, cell);
  c.porosity = GSL(S, RXN_F_POROSITY, 0, cell);
  c.soil_density = GSL(S, RXN_F_SOIL_PARTICLE_DENSITY, 0, cell);
  const double sat = GSL(S, RXN_F_SAT, 0, cell);
  c.psv = c.porosity * sat * 1000.0 * c.volume;
  c.psvd = rj ? c.porosity * sat * 1000.0 * c.volume / dt : 1.0;
  c.v_t = rj ? c.volume / dt : 1.0;
  c.den_kg_per_L = c.den_kg * 1.0 * 1.0e-3;
  grp_sync<G>(c);                                              // the previous cell of this column is done everywhere
  const long long xrow = a.xx_by_item ? item : cell;
#pragma unroll 1
  for (int i = c.l; i < N; i += G) {
    double mm = 1.0;                                           // padding rows of the shape: m = 1 (as tm_load)
    if (i < n) mm = a.xx ? a.xx[xrow * n + i] : GSL(S, RXN_F_PRI_MOLAL, i, cell);
    tsm[c.vm + i * CPB] = mm;
  }
  if (c.l == 0) {
    tsm[c.vlna + n * CPB] = 0.0;
    tsm[c.vlna + (n + 1) * CPB] = c.ln_act_h2o;
    tsm[c.vsm + lt.ncplx * CPB] = 0.0;
  }
  if (lt.act_off) {                                            // one class per species: ln gamma from the state
#pragma unroll 1
    for (int i = c.l; i < n; i += G) tsm[c.vlng + i * CPB] = c_log(GSL(S, RXN_F_PRI_ACT_COEF, i, cell));
#pragma unroll 1
    for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vlng + (n + k) * CPB] = c_log(GSL(S, RXN_F_SEC_ACT_COEF, k, cell));
  } else if (a.update_act) {                                   // lagged sec_molal for the ionic strength
#pragma unroll 4
    for (int k = c.l; k < lt.ncplx; k += G) tsm[c.vsm + k * CPB] = GSL(S, RXN_F_SEC_MOLAL, k, cell);
  }
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G) tsm[c.vfree + q * CPB] = GSL(S, RXN_F_FREE_SITE_CONC, q, cell);
#pragma unroll 1
  for (int q = c.l; q < lt.nkin; q += G) {
    tsm[c.vmnrl + q * CPB] = GSL(S, RXN_F_MNRL_VOLFRAC, q, cell);
    tsm[c.vmnrl + (lt.nkin + q) * CPB] = GSL(S, RXN_F_MNRL_AREA, q, cell);
  }
  if (lt.percell_logK)
    tm_percell_logK<CPB, G>(c.l, lt.ncoef, lt.logK_mode, c.vlk, c.temp, GSL(S, RXN_F_PRES, 0, cell), blob_d, h.cplx, h.kin, h.srf);
  if (rj && lt.nmr > 0) {                                      // multirate_prepare: R0_i = sum_r k_r/(1+k_r dt) S_r,i (even / odd rates)
#pragma unroll 1
    for (int ikr = 0; ikr < lt.nmr; ++ikr) {
      const int nrate = blob_i[h.o_mr_nrate + ikr];
#pragma unroll 1
      for (int i = c.l; i < n; i += G) {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
        for (int irate = 0; irate < nrate; irate += 2) {
          acc0 = acc0 + tm_mr_kk(kk, blob_d, h, ikr, irate, dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, ((long long)ikr * (h.mr_ld + 1) + irate + 1) * n + i, cell);
          if (irate + 1 < nrate)
            acc1 = acc1 + tm_mr_kk(kk, blob_d, h, ikr, irate + 1, dt) * GSL(S, RXN_F_KINMR_TOTAL_SORB, ((long long)ikr * (h.mr_ld + 1) + irate + 2) * n + i, cell);
        }
        tsm[c.vr0 + (ikr * N + i) * CPB] = acc0 + acc1;
      }
    }
  }
  const bool want_dtot = !rj && S.f[RXN_F_DTOTAL] != nullptr;
  const bool want_dsorb = !rj && lt.neqsorb > 0 && S.f[RXN_F_DTOTAL_SORB_EQ] != nullptr;
  const bool want_J = rj && a.jac_out != nullptr;
  // plan B writes (i, j) and (j, i), i.e. into rows of the other members: the rows are zeroed here, ordered before plan B by
  // the barriers of the speciation (as in tm_trip)
  if (want_dtot || want_J) tm_gi_zero_J<N, CPB, G>(c);
  grp_sync<G>(c);
  if (!lt.act_off && a.update_act) tm_act_coefs<N, CPB, G>(lt, c, true);     // RActivityCoefficients, LAG (:3994-4050)
  // RTotal (:4057-4158): ln a_i, then sec_molal_k
  if (from_state) {
#pragma unroll 2
    for (int i = c.l; i < n; i += G)
      tsm[c.vlna + i * CPB] = log(tsm[c.vm + i * CPB]) + log(GSL(S, RXN_F_PRI_ACT_COEF, i, cell));    // :4090
    grp_sync<G>(c);
#pragma unroll 1
    for (int g = c.l; g < lt.spec.ng; g += G) {
      const int4 hd = TI4(lt, (lt.spec.g0 >> 2) + 2 * g), h2 = TI4(lt, (lt.spec.g0 >> 2) + 2 * g + 1);
      const int c0 = hd.x, o0 = hd.y, nsteps = hd.z, cb = h2.x >> 2;
      const int4 m0 = TI4(lt, cb), m1 = TI4(lt, cb + 1), m2 = TI4(lt, cb + 2), m3 = TI4(lt, cb + 3);
      // the state's gamma_k of the 4 complexes of the group (padding members: 1), issued before the sums
      const double g0 = m0.z >= 0 ? GSL(S, RXN_F_SEC_ACT_COEF, m0.z, cell) : 1.0, g1 = m1.z >= 0 ? GSL(S, RXN_F_SEC_ACT_COEF, m1.z, cell) : 1.0,
                   g2 = m2.z >= 0 ? GSL(S, RXN_F_SEC_ACT_COEF, m2.z, cell) : 1.0, g3 = m3.z >= 0 ? GSL(S, RXN_F_SEC_ACT_COEF, m3.z, cell) : 1.0;
      double a0, a1, a2, a3;
      if (lt.percell_logK) {
        a0 = tsm[c.vlk + (m0.z < 0 ? 0 : m0.z) * CPB]; a1 = tsm[c.vlk + (m1.z < 0 ? 0 : m1.z) * CPB];
        a2 = tsm[c.vlk + (m2.z < 0 ? 0 : m2.z) * CPB]; a3 = tsm[c.vlk + (m3.z < 0 ? 0 : m3.z) * CPB];
      } else {
        const double2 ia = TD2(lt, c0 * 2), ib = TD2(lt, c0 * 2 + 1);
        a0 = ia.x; a1 = ia.y; a2 = ib.x; a3 = ib.y;
      }
      tm_run_group(lt, s, c0, o0, nsteps, a0, a1, a2, a3);
      tsm[m0.x + s] = exp(a0) / g0; tsm[m1.x + s] = exp(a1) / g1; tsm[m2.x + s] = exp(a2) / g2; tsm[m3.x + s] = exp(a3) / g3;   // :4112
    }
    grp_sync<G>(c);
  } else {
    tm_speciate<N, CPB, G>(lt, c);
  }
  tm_planA<N, CPB, G>(lt, c);
  if (want_dtot || want_J)
    tm_planB<N, CPB, G>(lt, c, rj ? c.den_kg_per_L * c.psvd : c.den_kg_per_L);      // dtotal [* psvd_t, RTAccumulationDerivative :5189-5204]
  if (want_dtot) {
    grp_sync<G>(c);
    tm_gi_store_block<N, CPB, G>(lt, c, S, RXN_F_DTOTAL, on);
  }
  if (want_dsorb) { grp_sync<G>(c); tm_gi_zero_J<N, CPB, G>(c); }
#pragma unroll 1
  for (int i = c.l; i < N; i += G) tsm[c.vres + i * CPB] = 0.0;
  if (lt.neq > 0 && on) {                                      // RZeroSorb :4162-4178
#pragma unroll 1
    for (int k = c.l; k < lt.nsrf; k += G) GSL(S, RXN_F_EQSRFCPLX_CONC, k, cell) = 0.0;
  }
  grp_sync<G>(c);
  // RTotalSorb (:4182-4216); GI_RJ: + the equilibrium part of the multirate reactions (RMultiRateSorption)
  const int ntask = lt.neq + (rj ? lt.nmr : 0);
#pragma unroll 1
  for (int task = 0; task < ntask; ++task) {
    const bool eq = task < lt.neq;
    const int ikr = task - lt.neq;
    int tb = c.vres;
    double fac = rj ? c.v_t : 1.0;
    if (!eq) {
      tb = c.vseq + ikr * N * CPB;
      fac = c.volume * lt.mrK1[ikr];
#pragma unroll 1
      for (int i = c.l; i < n; i += G) tsm[tb + i * CPB] = 0.0;
    }
    tm_srf_rxn<N, CPB, G>(lt, c, S, TI(lt, (eq ? lt.i_eq_rxn : lt.i_mr_rxn - lt.neq) + task), fac, eq ? (want_J || want_dsorb) : want_J, eq, on, tb);
  }
  tm_wait_st();
  grp_sync<G>(c);
  if (want_dsorb) { tm_gi_store_block<N, CPB, G>(lt, c, S, RXN_F_DTOTAL_SORB_EQ, on); grp_sync<G>(c); }
  bool bad = false;
#pragma unroll 1
  for (int i = c.l; i < n; i += G) {
    const double mm = tsm[c.vm + i * CPB];
    const double tot = (mm + tsm[c.vtot + i * CPB]) * c.den_kg_per_L;               // :4095, 4124, 4148
    const double tsorb = lt.neqsorb > 0 ? tsm[c.vres + i * CPB] : 0.0;
    double acc = c.psv * tot;                                                       // RTAccumulation :5072-5148
    if (lt.neqsorb > 0) acc = acc + tsorb * c.volume;                               // RAccumulationSorb :4539-4568
    if (!isfinite(tot) || !isfinite(acc)) bad = true;
    if (on) {
      GSL(S, RXN_F_TOTAL, i, cell) = tot;
      if (lt.neqsorb > 0) GSL(S, RXN_F_TOTAL_SORB_EQ, i, cell) = tsorb;
      if (!rj) {
        if (a.xx) GSL(S, RXN_F_PRI_MOLAL, i, cell) = mm;
        if (a.accum_out) a.accum_out[item * n + i] = acc;
        if (!lt.act_off && a.update_act) GSL(S, RXN_F_PRI_ACT_COEF, i, cell) = c_exp(tsm[c.vlng + TI(lt, lt.i_pcls + i) * CPB]);
      }
    }
    if (rj) tsm[c.vres + i * CPB] = acc / dt;
  }
  if (rj) {
    // RReaction :3515-3584 (minerals, then multirate)
    if (lt.nkin > 0) { grp_sync<G>(c); tm_kinetic_mineral<N, CPB, G>(lt, c, true); }
#pragma unroll 1
    for (int ikr = 0; ikr < lt.nmr; ++ikr) {
#pragma unroll 1
      for (int i = c.l; i < n; i += G) {
        tsm[c.vres + i * CPB] += c.volume * (lt.mrK1[ikr] * tsm[c.vseq + (ikr * N + i) * CPB] - tsm[c.vr0 + (ikr * N + i) * CPB]);
        if (on) GSL(S, RXN_F_KINMR_TOTAL_SORB, (long long)ikr * (h.mr_ld + 1) * n + i, cell) = tsm[c.vseq + (ikr * N + i) * CPB];
      }
    }
    tm_wait_st();
    grp_sync<G>(c);
#pragma unroll 1
    for (int i = c.l; i < n; i += G) {
      const double r = tsm[c.vres + i * CPB];
      if (!isfinite(r)) bad = true;
      if (a.res_out && on) a.res_out[item * n + i] = r;
    }
    if (want_J) {
      // block column-major, with respect to m_j: Jln_ij / m_j.  A member writes whole COLUMNS of the block (n contiguous
      // doubles of the caller's array) instead of the rows it assembled: column j of the cell's TMEM lane is read element by
      // element (a row write would scatter n 8-byte stores over n sectors: measured 35 % of this kernel's stall samples)
#pragma unroll 1
      for (int j = c.l; j < n; j += G) {
        const double invm = 1.0 / tsm[c.vm + j * CPB];
        double col[N];
#pragma unroll
        for (int i = 0; i < N; ++i) col[i] = tm_ld_el(c.tb, i, j) * invm;
        double *dst = a.jac_out + item * (long long)(n * n) + j * n;
        if (on) {
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (i < n) dst[i] = col[i];
        }
      }
    }
#pragma unroll 1
    for (int q = c.l; q < lt.nkin; q += G)
      if (on) GSL(S, RXN_F_MNRL_RATE, q, cell) = tsm[c.vmnrl + (2 * lt.nkin + q) * CPB];
  }
#pragma unroll 2
  for (int k = c.l; k < lt.ncplx; k += G) {
    if (on) {
      GSL(S, RXN_F_SEC_MOLAL, k, cell) = tsm[c.vsm + k * CPB];
      if (!rj && !lt.act_off && a.update_act) GSL(S, RXN_F_SEC_ACT_COEF, k, cell) = c_exp(tsm[c.vlng + TI(lt, lt.i_ccls + k) * CPB]);
    }
  }
#pragma unroll 1
  for (int q = c.l; q < lt.nrxn; q += G)
    if (on) GSL(S, RXN_F_FREE_SITE_CONC, q, cell) = tsm[c.vfree + q * CPB];
  if (c.l == 0 && on && !rj && a.update_act) GSL(S, RXN_F_LN_ACT_H2O, 0, cell) = c.ln_act_h2o;
  {
    const int fl = c.flags | (bad ? RXN_FLAG_NONFINITE : 0);
    if (fl != 0 && on && S.fail) {
#ifndef RXN_TM_HOST
      atomicOr(S.fail, (unsigned int)fl);
#else
      __atomic_fetch_or(S.fail, (unsigned int)fl, __ATOMIC_RELAXED);
#endif
    }
  }
  warp_converge();
}

}  // namespace tmk
}  // namespace rxn
