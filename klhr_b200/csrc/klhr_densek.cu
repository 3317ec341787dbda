// klhr_b200 -- instantiations of the dense kernels (klhr_densek.cuh, klhr_densews.cuh): stan/corr-normal.stan, D = 128 or 256.
#include <cstdlib>
#include "klhr_densews.cuh"

namespace klhr {

// KLHR_DENSE_PLAIN=1 in the environment keeps free-running launches on dense_kernel (measurements, tests)
static bool force_plain() {
    static const bool v = [] { const char* e = std::getenv("KLHR_DENSE_PLAIN"); return e && e[0] == '1'; }();
    return v;
}

int launch_densek(const StepArgs& a, bool replay, cudaStream_t st, LaunchInfo* info) {
    // free-running launches without thinned output: the warp-specialised kernel; replay (fp64 directions) and
    // sample() rows: dense_kernel
    const bool ws = !replay && a.acc.draws == nullptr && !force_plain() && densews_smem_bytes(a) <= (size_t)227 * 1024;
    if (ws) return a.mp.D == 256 ? launch_densews_nt<4>(a, st, info) : launch_densews_nt<2>(a, st, info);
    return a.mp.D == 256 ? launch_densek_nt<4>(a, replay, st, info) : launch_densek_nt<2>(a, replay, st, info);
}

}  // namespace klhr
