// rxn_small.cuh — host interface of the register RReact kernel for small chemistries (rxn_small.h, rxn_small_dev.cuh)
#pragma once
#include <cuda_runtime.h>

#include "rxn_small.h"

namespace rxn {

int small_launch_react(const SmallPlan &p, const DevState &S, double *tran_xx, const int32_t *l2g, long long nlocal, double dt, int dt_mode,
                       int32_t *iters, int32_t *flags, cudaStream_t stream, long long cell0 = 0);

}  // namespace rxn
