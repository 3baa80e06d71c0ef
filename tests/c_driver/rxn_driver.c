/* rxn_driver.c - plain C99 host of the B200 reaction library: the call sequence a Fortran / C caller walks, with nothing but
 * include/rxn_b200.h.  TEST INFRASTRUCTURE.  Build (tests/test_c_driver.py does it):
 *   python tests/c_driver/gen_case.py calcite 512 case.h
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude -DCASE_HEADER='"case.h"' tests/c_driver/rxn_driver.c -L pflotran_b200 -lrxn_b200 -lm
 * Sequence (reference call sites in reactive_transport.F90): tables once (after BasisInit) -> state (RTAuxVarInit) -> flow
 * coupling scalars -> RTUpdateAuxVars -> RTUpdateFixedAccumulation -> residual / Jacobian blocks -> RTUpdateKineticState,
 * then an operator-split RTReact on the transported totals.  Results go to a binary file the test compares with the same
 * calls made through ctypes: [n][ncomp] accumulation, residual, [n][ncomp^2] Jacobian, [n][ncomp] free-ion result of RTReact,
 * [n] iteration counts and flags as doubles. */
#include <stdio.h>
#include <stdlib.h>
#include CASE_HEADER

static int check(int rc, const char *what) {
  if (rc != RXN_OK) {
    char buf[1024];
    rxn_last_error(buf, (int32_t)sizeof buf);
    fprintf(stderr, "%s failed: status %d: %s\n", what, rc, buf);
  }
  return rc;
}
#define CK(call) do { if (check((call), #call) != RXN_OK) return 2; } while (0)

int main(int argc, char **argv) {
  const int n = CASE_NCELLS, nc = CASE_NCOMP;
  const double dt = 1800.0;
  RxnTablesDesc desc;
  RxnTables *tables = NULL;
  RxnState *state = NULL;
  double *xx, *accum, *res, *jac, *out_iters;
  int32_t *iters, *flags;
  FILE *f;
  int i;
  if (argc < 2) { fprintf(stderr, "usage: rxn_driver out.bin [compile-only]\n"); return 1; }
  printf("%s\n", rxn_version());
  case_fill_desc(&desc);
  if (desc.struct_size != (int32_t)sizeof(RxnTablesDesc)) { fprintf(stderr, "descriptor size mismatch\n"); return 1; }
  if (argc > 2) return 0;                         /* link check only (no GPU) */
  CK(rxn_tables_create(&desc, 0, &tables));
  CK(rxn_state_create(tables, (int64_t)n, &state));
  for (i = 0; i < CASE_NBASE; ++i) CK(rxn_state_broadcast(state, case_base[i].field, case_base[i].values));
  CK(rxn_set_cell_scalars(state, NULL, NULL, case_temp, case_pres, NULL, case_porosity, NULL, NULL));
  if (CASE_NKIN > 0) CK(rxn_state_upload(state, RXN_F_MNRL_VOLFRAC, case_volfrac, (int64_t)n, 1));
  xx = (double *)malloc(sizeof(double) * (size_t)n * nc);
  accum = (double *)malloc(sizeof(double) * (size_t)n * nc);
  res = (double *)malloc(sizeof(double) * (size_t)n * nc);
  jac = (double *)malloc(sizeof(double) * (size_t)n * nc * nc);
  out_iters = (double *)malloc(sizeof(double) * (size_t)n * 2);
  iters = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  flags = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  if (!xx || !accum || !res || !jac || !out_iters || !iters || !flags) return 1;
  /* a global-implicit Newton iterate: the base free-ion molalities, 2 % up */
  for (i = 0; i < n * nc; ++i) xx[i] = case_base[0].values[i % nc] * 1.02;
  if (case_base[0].field != RXN_F_PRI_MOLAL) { fprintf(stderr, "case: first base row is not PRI_MOLAL\n"); return 1; }
  CK(rxn_update_auxvars_batch(state, xx, 1));
  CK(rxn_fixed_accum_batch(state, xx, NULL, (int64_t)n, accum));
  CK(rxn_residual_jacobian_blocks_batch(state, NULL, (int64_t)n, dt, res, jac));
  CK(rxn_update_kinetic_state_batch(state, dt));
  /* operator split: RTReact on the transported totals */
  for (i = 0; i < n * nc; ++i) xx[i] = case_tran_xx[i];
  CK(rxn_react_batch(state, xx, NULL, (int64_t)n, 3600.0, RXN_DT_CONSISTENT, iters, flags));
  for (i = 0; i < n; ++i) { out_iters[i] = (double)iters[i]; out_iters[n + i] = (double)flags[i]; }
  f = fopen(argv[1], "wb");
  if (!f) return 1;
  fwrite(accum, sizeof(double), (size_t)n * nc, f);
  fwrite(res, sizeof(double), (size_t)n * nc, f);
  fwrite(jac, sizeof(double), (size_t)n * nc * nc, f);
  fwrite(xx, sizeof(double), (size_t)n * nc, f);
  fwrite(out_iters, sizeof(double), (size_t)n * 2, f);
  fclose(f);
  printf("kernel of the last call: %.3f ms\n", (double)rxn_last_kernel_ms(state));
  CK(rxn_state_destroy(state));
  CK(rxn_tables_destroy(tables));
  free(xx); free(accum); free(res); free(jac); free(out_iters); free(iters); free(flags);
  return 0;
}
