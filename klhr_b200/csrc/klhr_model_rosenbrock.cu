// klhr_b200 -- step/eval kernel instantiations for one Stan target (one translation unit
// per model so the build parallelises).  See klhr_models.cuh for the density itself.
#include "klhr_chain.cuh"
#include "klhr_mh.cuh"
#include "klhr_slice.cuh"

namespace klhr {
using M64_rosenbrock = Rosenbrock<double>;
using M32_rosenbrock = Rosenbrock<float>;
KLHR_DEFINE_MODEL(rosenbrock, M64_rosenbrock, M32_rosenbrock)
KLHR_DEFINE_MODEL_CHAIN(rosenbrock, M64_rosenbrock, M32_rosenbrock)
KLHR_DEFINE_MODEL_MH(rosenbrock, M64_rosenbrock, M32_rosenbrock)
KLHR_DEFINE_MODEL_SLICE(rosenbrock, M64_rosenbrock, M32_rosenbrock)
}  // namespace klhr
