// emul.cpp — HOST COMPILATION OF THE DEVICE ROUTINES, for tests only.
//
// TEST INFRASTRUCTURE.  This file compiles pflotran_b200/csrc/rxn_device.cuh (the per-cell
// device code of the CUDA kernels) with g++ by defining the CUDA qualifiers away, and runs the
// per-cell bodies in a plain loop over a host SoA image.  It exists so that the `-m "not gpu"`
// suite can check the device code's logic and the table packer (rxn_pack.h) against the oracle
// in the build container, which has no GPU.  It is NOT part of the product, is never loaded by
// pflotran_b200, and is not a fallback: the library (librxn_b200.so) fails with
// RXN_ERR_NO_DEVICE when no CUDA device is present.
#define __device__
#define __host__
#define __forceinline__ inline
#include <cmath>
using std::isfinite;
#include "../../pflotran_b200/csrc/rxn_pack.h"
#include "../../pflotran_b200/csrc/rxn_device.cuh"

using namespace rxn;

struct HostView { int64_t ncells, ld; double *f[RXN_F_COUNT]; };

struct Emu {
  PackResult R;
  std::vector<unsigned char> blob;
  Tab T;
  int nv;
};

static DevState mk_state(const HostView *v, const uint8_t *active) {
  DevState S;
  for (int f = 0; f < RXN_F_COUNT; ++f) S.f[f] = v->f[f];
  S.ld = v->ld; S.ncells = v->ncells; S.active = active;
  return S;
}

#define EMU_DISPATCH(nv, CALL)          \
  switch (nv) {                         \
    case 4: { constexpr int N = 4; CALL; } break;   \
    case 8: { constexpr int N = 8; CALL; } break;   \
    case 16: { constexpr int N = 16; CALL; } break; \
    default: { constexpr int N = 24; CALL; } break; \
  }

extern "C" {

void *emu_create(const RxnTablesDesc *d, char *err, int errlen) {
  Emu *e = new Emu();
  int rc = pack_tables(d, e->R);
  if (rc != RXN_OK) {
    if (err && errlen > 0) { strncpy(err, e->R.err.c_str(), errlen - 1); err[errlen - 1] = 0; }
    delete e;
    return nullptr;
  }
  e->blob = blob_bytes(e->R);
  e->T.d = reinterpret_cast<const double *>(e->blob.data());
  e->T.i = reinterpret_cast<const int *>(e->blob.data() + (size_t)e->R.h.ndbl * 8);
  e->T.h = &e->R.h;
  e->nv = variant_for(e->R.h.naq);
  return e;
}
void emu_destroy(void *h) { delete (Emu *)h; }
int emu_pack_status(const RxnTablesDesc *d, char *err, int errlen) {
  PackResult R;
  int rc = pack_tables(d, R);
  if (rc != RXN_OK && err && errlen > 0) { strncpy(err, R.err.c_str(), errlen - 1); err[errlen - 1] = 0; }
  return rc;
}
int emu_field_rows(void *h, int f) { return ((Emu *)h)->R.rows[f]; }
void emu_set_maxit(void *h, int maxit) { ((Emu *)h)->R.h.maxit = maxit; }

int emu_react_batch(void *h, const HostView *v, double *tran_xx, const uint8_t *active, const int32_t *l2g, int64_t nlocal,
                    double dt, int dt_mode, int32_t *iters, int32_t *flags) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long i = 0; i < nlocal; ++i) EMU_DISPATCH(e->nv, cell_react<N>(e->T, S, i, tran_xx, l2g, dt, dt_mode, iters, flags));
  return 0;
}
int emu_update_auxvars_batch(void *h, const HostView *v, const double *xx_loc, const uint8_t *active, int update_act_coefs) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long c = 0; c < v->ncells; ++c) EMU_DISPATCH(e->nv, cell_update_auxvars<N>(e->T, S, c, xx_loc, update_act_coefs));
  return 0;
}
int emu_fixed_accum_batch(void *h, const HostView *v, const double *xx, const uint8_t *active, const int32_t *l2g, int64_t nlocal,
                          double *accum_out) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long i = 0; i < nlocal; ++i) EMU_DISPATCH(e->nv, cell_fixed_accum<N>(e->T, S, i, xx, l2g, accum_out));
  return 0;
}
int emu_residual_jacobian_batch(void *h, const HostView *v, const uint8_t *active, const int32_t *l2g, int64_t nlocal, double dt,
                                double *res_out, double *jac_out) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long i = 0; i < nlocal; ++i) EMU_DISPATCH(e->nv, cell_residual_jacobian<N>(e->T, S, i, l2g, dt, res_out, jac_out));
  return 0;
}
int emu_update_kinetic_state_batch(void *h, const HostView *v, const uint8_t *active, double dt) {
  Emu *e = (Emu *)h;
  DevState S = mk_state(v, active);
  for (long long c = 0; c < v->ncells; ++c) EMU_DISPATCH(e->nv, cell_update_kinetic_state<N>(e->T, S, c, dt));
  return 0;
}

}  // extern "C"
