// rxn_tile_variant.cu — one instantiation of the cooperative RReact kernel: TILE_G lanes per cell
// (compiled once per group width, see Makefile).
#if !defined(TILE_G)
#error "compile with -DTILE_G=<lanes per cell>"
#endif
#include "rxn_tile_dev.cuh"

namespace rxn {
template void tile_launch_variant<TILE_G>(const TilePlan &, const DevTab &, const double *, const DevState &, double *,
                                          const int32_t *, long long, double, int, int32_t *, int32_t *, cudaStream_t);
}
