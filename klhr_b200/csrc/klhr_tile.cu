// klhr_b200 -- instantiations of the tile kernel (klhr_tile.cuh) for the diagonal-Gaussian targets.
#include "klhr_tile.cuh"

namespace klhr {

int launch_tile(const StepArgs& a, int dtype, bool replay, cudaStream_t st, LaunchInfo* info) {
    const bool scaled = a.mp.id == KLHR_MODEL_ILL_NORMAL;
    if (dtype == KLHR_F64)
        return scaled ? launch_tile_typed<double, true>(a, replay, st, info)
                      : launch_tile_typed<double, false>(a, replay, st, info);
    return scaled ? launch_tile_typed<float, true>(a, replay, st, info)
                  : launch_tile_typed<float, false>(a, replay, st, info);
}

}  // namespace klhr
