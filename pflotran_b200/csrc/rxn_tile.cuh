// rxn_tile.cuh — cooperative RReact kernel: a group of G lanes (G = 4, 8, 16 or 32, a slice of
// one warp) solves one cell; lane l owns rows l, l+G, ... (R rows per lane) of the Newton system.
//
// Why: the Newton matrix of one cell is naq x naq doubles (300A: 15 x 15 = 1.8 KB).  One thread
// per cell (rxn_device.cuh) has to keep it in local memory, which spills to L1/L2/HBM and runs at
// ~2 % of the FP64 roofline.  Here every per-cell array lives in shared memory
// (J + 4 vectors + sec_molal ~ 3.7 KB/cell for 300A, ~48 cells per SM), the chemistry tables and
// a host-built accumulation plan are staged once per (persistent) CTA, and the work of one cell
// is spread over the lanes of its group:
//   * complexes k = l, l+G, ...      -> lnQK sums + one exp each            (RTotal, reaction.F90:4104-4122)
//   * total / dtotal entries          -> balanced sparse plans A and B      (RTotal, :4124-4146)
//   * rows of the residual / Jacobian -> owner lane                         (RReact, :3424-3437)
//   * LU with partial pivoting        -> right-looking, one row per owner   (ludcmp, utility.F90:393-476)
// Group-wide reductions are xor-butterflies over the group's lanes; every lane ends up with
// bit-identical results, so the control flow of a group is uniform.
//
// Arithmetic differences from the reference order (all deterministic, all far below the 1e-10
// parity bar; each marked REASSOC at its site):
//   - sec_molal = exp(lnQK - ln gamma) instead of exp(lnQK)/gamma; ln gamma of a species is the
//     Debye-Hueckel exponent itself instead of log(exp(exponent));
//   - d(total_i)/d(m_j) = (sum_k nu_ik nu_jk sec_molal_k) / m_j instead of one
//     exp(lnQK - ln m_j)/gamma per (complex, species) pair: removes S = sum nspec exps per
//     Newton iteration (202 of ~430 transcendentals for 300A);
//   - ionic strength and free-site sums are tree reductions over the group;
//   - sorption / multirate derivative blocks are added into J term by term instead of through
//     a dense naq x naq temporary;
//   - back-substitution subtracts columns in descending order.
// The LU itself applies, per matrix element, the same operations in the same order as Crout's
// method with the reference's implicit-scaling pivot rule (including the `>=` tie-break).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "rxn_tab.h"

namespace rxn {

// CTA size bound per lane-group width (registers per thread = 64K / bound)
constexpr int tile_max_threads(int G) { return G <= 2 ? 128 : G <= 8 ? 384 : G == 16 ? 768 : 1024; }

struct TileTab {
  int G, NP, LDJ, threads, cpb, gpw;   // gpw: lane groups (cells) per warp
  int ncls, maxsrf, need_gam, percell_logK;
  // plan blob: doubles.  *_rec: 16-byte records {double coef; int code; int key}, lane l at [t*G + l]
  int o_A_rec, o_B_rec, o_cls_z2, o_cls_a0, o_pz2, o_cz2, o_nlk;
  // plan blob: ints
  int o_pri_cls, o_cplx_cls;
  int TA, TB;
  int ndbl, nint;
  // per-cell shared memory (offsets in doubles; J = [NP rows][LDJ] at 0, b = column NP)
  int c_m, c_invm, c_lna, c_tot, c_lgp, c_fix, c_tsorb, c_gam, c_sm, c_lng, c_sc, c_dsx, c_free, c_mnrl, c_r0, c_seq, c_lk, pc_dbl;
};

struct TilePlan {
  bool usable = false;
  std::string err = "not built";
  TileTab tt;
  double *d_blob = nullptr;
  size_t blob_bytes = 0;
  size_t smem_bytes = 0;
  int grid = 0;
  // plan statistics (DESIGN.md / bench)
  int termsA = 0, termsB = 0;
};

int tile_plan_build(const RxnTablesDesc *d, const DevTab &h, const std::vector<double> &bd, const std::vector<int32_t> &bi,
                    size_t main_blob_bytes, int device, TilePlan *p);
void tile_plan_free(TilePlan *p);
int tile_launch_react(const TilePlan &p, const DevTab &h, const double *blob, const DevState &S, double *tran_xx, const int32_t *l2g,
                      long long nlocal, double dt, int dt_mode, int32_t *iters, int32_t *flags, cudaStream_t stream);

}  // namespace rxn
