// klhr_b200 -- instantiations of the dense kernel (klhr_densek.cuh): stan/corr-normal.stan, D = 128 or 256.
#include "klhr_densek.cuh"

namespace klhr {

int launch_densek(const StepArgs& a, bool replay, cudaStream_t st, LaunchInfo* info) {
    return a.mp.D == 256 ? launch_densek_nt<4>(a, replay, st, info) : launch_densek_nt<2>(a, replay, st, info);
}

}  // namespace klhr
