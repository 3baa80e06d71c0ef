// rxn_tile.cuh — cooperative (lane-group per cell) RReact kernel.  Placeholder plan: not built yet.
#pragma once
#include <string>
#include <vector>
#include "rxn_tab.h"

namespace rxn {
struct TilePlan {
  bool usable = false;
  std::string err = "not built";
};
inline int tile_plan_build(const RxnTablesDesc *, const DevTab &, const std::vector<double> &, const std::vector<int32_t> &, TilePlan *p) {
  p->usable = false;
  return RXN_OK;
}
inline void tile_plan_free(TilePlan *) {}
inline void tile_launch_react(const TilePlan &, const DevTab &, const double *, const DevState &, double *, const int32_t *,
                              long long, double, int, int32_t *, int32_t *, cudaStream_t) {}
}  // namespace rxn
