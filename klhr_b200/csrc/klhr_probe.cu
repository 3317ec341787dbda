// klhr_b200 -- peak probes: micro-kernels that measure, on the device the benchmark runs on, the ceilings the
// roofline of the KLHR step is stated against (BASELINE.md section 3 asks for an FMA micro-kernel: FP64 / FP32
// vector peaks are not in MEASURED_PEAKS.json).  bench.py times them with CUDA events in its own process and
// prints them as `peaks`.
//   kind 0  fp64 FMA      8 independent DFMA chains per thread                     ops = DFMA
//   kind 1  normals       Philox4x32-10 -> fp32 Box-Muller, the direction stream    ops = normals
//                         of the step kernels (klhr_common.cuh), summed
//   kind 2  fp32 FMA      8 independent FFMA chains per thread                     ops = FFMA
//   kind 3  cvt.f64.f32   8 conversions + 8 DADD per trip                           ops = conversions
//   kind 4  MUFU          lg2 / sqrt / sin / cos round-robin                        ops = MUFU
//   kind 5  mul.wide.u32  the Philox multiply                                       ops = IMAD.WIDE
//   kind 6  fp32 -> fp64 promotion on the integer pipe + DADD                        ops = conversions
//   kind 7  DMMA          mma.sync.m8n8k4.f64, 8 independent accumulator fragments   ops = FMA (256 per warp-level DMMA)
#include <cmath>
#include "klhr_lane.cuh"
#include "klhr_dense.cuh"

namespace klhr {

template <int kKind>
__global__ void __launch_bounds__(256) probe_kernel(long long iters, double* out, double sentinel) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0;
    if constexpr (kKind == 0) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 1.0 + 1e-9 * (tid + k);
        const double a = 1.0 - 1e-12 * (tid & 7), b = 1e-13;
        for (long long it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    } else if constexpr (kKind == 1) {
        float s = 0;
        for (long long it = 0; it < iters; ++it) {
            uint32_t w[4][4];
            Philox::blockN<4>(tid, 0u, (uint32_t)it, kSlotDir, 1u, 0x1234u, 0x5678u, w);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float z0, z1, z2, z3;
                box_muller_f32(w[jj][0], w[jj][1], z0, z1);
                box_muller_f32(w[jj][2], w[jj][3], z2, z3);
                s += (z0 + z1) + (z2 + z3);
            }
        }
        acc = s;
    } else if constexpr (kKind == 2) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 1.0f + 1e-6f * (tid + k);
        const float a = 1.0f - 1e-7f * (tid & 7), b = 1e-8f;
        for (long long it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], a, b);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    } else if constexpr (kKind == 3 || kKind == 6) {
        float f[8];
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { f[k] = 1.0f + 0.125f * ((tid + k) & 7); v[k] = 0; }
        for (long long it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double d;
                if constexpr (kKind == 3) {
                    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f[k]));
                } else {
                    const uint32_t b = __float_as_uint(f[k]);
                    const uint32_t hi = (((b >> 3) & 0x0fffffffu) + 0x38000000u) | (b & 0x80000000u);
                    d = __hiloint2double((int)hi, (int)(b << 29));
                }
                v[k] += d;
                f[k] = __uint_as_float(__float_as_uint(f[k]) ^ ((uint32_t)it & 1u));   // keeps the conversion in the loop
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    } else if constexpr (kKind == 4) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.5f + 0.01f * ((tid + k) & 15);
        for (long long it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; k += 4) {
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
                asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k + 1]));
                asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(v[k + 2]));
                asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(v[k + 3]));
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    } else if constexpr (kKind == 5) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = tid * 2654435761u + k;
        for (long long it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t p = (uint64_t)v[k] * 0xD2511F53u;
                v[k] = (uint32_t)(p >> 32) ^ (uint32_t)p;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    }
    else if constexpr (kKind == 7) {
        double d[8][2];
#pragma unroll
        for (int k = 0; k < 8; ++k) { d[k][0] = 1e-3 * (tid & 3); d[k][1] = 0; }
        const double fa = 1.0 - 1e-9 * (tid & 31), fb = 0.25;
        for (long long it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma_m8n8k4(d[k][0], d[k][1], fa, fb);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += d[k][0] + d[k][1];
    }
    if (acc == sentinel) out[0] = acc;       // sentinel = NaN: never true, and the compiler cannot know
}

// Philox4x32-10 blocks for given (counter[4], key[2]) rows: the known-answer test of the in-kernel generator
__global__ void philox_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[4];
    Philox::block(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], in[6 * i + 4], in[6 * i + 5], w);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = w[k];
}

}  // namespace klhr

using namespace klhr;

extern "C" int klhr_philox_eval(const uint32_t* ctr_key_dev, uint32_t* out_dev, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!ctr_key_dev || !out_dev))) return -1;
    if (n == 0) return 0;
    philox_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctr_key_dev, out_dev, n);
    return (int)cudaGetLastError();
}

extern "C" int64_t klhr_peak_probe(int kind, int64_t iters, int ctas, double* out_dev, void* stream) {
    if (iters <= 0 || ctas <= 0 || !out_dev) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    long long per_iter = 8;
    const double nan = std::nan("");
    switch (kind) {
        case 0: probe_kernel<0><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;
        case 1: probe_kernel<1><<<ctas, 256, 0, st>>>(iters, out_dev, nan); per_iter = 16; break;
        case 2: probe_kernel<2><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;
        case 3: probe_kernel<3><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;
        case 4: probe_kernel<4><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;
        case 5: probe_kernel<5><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;
        case 6: probe_kernel<6><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;
        case 7: probe_kernel<7><<<ctas, 256, 0, st>>>(iters, out_dev, nan); break;      // 8 FMA per thread per DMMA
        default: return -2;
    }
    if (cudaGetLastError() != cudaSuccess) return -3;
    return (int64_t)(per_iter * iters * 256LL * ctas);     // operations issued by the launch (see the table above)
}
