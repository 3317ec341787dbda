// klhr_b200 -- instantiations of the lane kernel (klhr_lane.cuh) for the diagonal-Gaussian targets.
#include <cstdlib>
#include "klhr_lane.cuh"

namespace klhr {

// lanes per chain: 2 by default (klhr_lane.cuh); KLHR_LANE_G=1 in the environment selects the one-thread-per-chain
// shape for measurements (tools/perf_probe.py)
static int lane_g() {
    static const int g = [] {
        const char* e = std::getenv("KLHR_LANE_G");
        return (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : KLHR_LANE_G);
    }();
    return g;
}

// KLHR_LANE_TAIL=0 in the environment keeps the last, single-quad trip inside the trip loop (measurements)
static bool lane_tail_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("KLHR_LANE_TAIL");
        return !(e && e[0] == '0');
    }();
    return on;
}

int launch_lane(const StepArgs& a, cudaStream_t st, LaunchInfo* info) {
    const bool scaled = a.mp.id == KLHR_MODEL_ILL_NORMAL;
    if (lane_g() == 1)
        return scaled ? launch_lane_typed<true, 1, false>(a, st, info) : launch_lane_typed<false, 1, false>(a, st, info);
    if (lane_split_tail(a.mp.D) && lane_tail_enabled())
        return scaled ? launch_lane_typed<true, 2, true>(a, st, info) : launch_lane_typed<false, 2, true>(a, st, info);
    return scaled ? launch_lane_typed<true, 2, false>(a, st, info) : launch_lane_typed<false, 2, false>(a, st, info);
}

}  // namespace klhr
