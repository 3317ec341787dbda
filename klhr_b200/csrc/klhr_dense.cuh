// klhr_b200 -- CTA-cooperative dense quadratic forms on the FP64 tensor cores.
//
// stan/corr-normal.stan has a dense precision P (D x D, shared by all chains); the line
// restriction needs A = rho^T P rho and Bq = -theta^T P rho for every chain and draw, i.e. the
// row-wise products of the CTA's rho tile with P: V = R P (chains x D), a small GEMM.  Doing it
// per chain (CorrNormal::setup) re-reads all of P for every chain; here the CTA's warps split
// the columns of P, every element of P is loaded once per CTA and draw, and the products run on
// mma.sync.m8n8k4.f64 (DMMA).  V never leaves registers: each lane folds its 8x8 accumulator
// fragments straight into partial sums of rho.V and theta.V.
#pragma once
#include "klhr_common.cuh"

namespace klhr {

// Chains per CTA on the dense path = kDenseMT m8 row tiles; every B fragment of P loaded from L2 feeds kDenseMT
// DMMAs, and a warp keeps kDenseMT x kDenseQT accumulator fragments (64 registers either way).  Measured on
// B200, corr-normal fp64: 32 chains (one 256-thread CTA per SM, ~208 registers) beat 16 chains (3 CTAs per SM,
// 128 registers) by 34 % at D = 256 and 21 % at D = 128 -- the kernel waits on the L2 -> SM stream of P.
#ifndef KLHR_DENSE_CHAINS
#define KLHR_DENSE_CHAINS 32
#endif
constexpr int kDenseMaxChains = KLHR_DENSE_CHAINS;
constexpr int kDenseMT = kDenseMaxChains / 8;          // m8 row tiles per warp
constexpr int kDenseQT = 16 / kDenseMT;                // column tiles per warp and chunk
static_assert(kDenseMaxChains == 16 || kDenseMaxChains == 32, "dense path: 16 or 32 chains per CTA");

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// th_tile / rh_tile: [cpb][Dpad] rows in shared memory (rows of absent chains must be zero).
// red: [n_warps][kDenseMaxChains][2] scratch.  Every thread of the CTA must call this (it synchronises).
// Returns A and Bq of chain `o` (octet-uniform).
__device__ inline void dense_cta_quadratic(const double* th_tile, const double* rh_tile, const double* __restrict__ P,
                                           int D, int Dpad, int cpb, double (*red)[kDenseMaxChains][2], int o,
                                           double& A_out, double& Bq_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int r8 = lane >> 2, k4 = lane & 3;          // fragment coordinates
    __syncthreads();                                   // rho / theta rows of every chain are in place
    constexpr int MT = kDenseMT, QT = kDenseQT;
    double pa[MT], pb[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) pa[m] = pb[m] = 0.0;
    const int n_tiles = (D + 7) >> 3;
    for (int chunk = 0; chunk * QT * n_warps < n_tiles; ++chunk) {
        // this warp's up to QT column tiles of the chunk: nt = (chunk * QT + q) * n_warps + warp
        double acc[MT][QT][2];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[m][q][0] = acc[m][q][1] = 0.0;
        // Fast path (every column tile of the chunk is full and D % 4 == 0): no bounds predicates, one
        // base pointer per tile, and the 8 B fragments of step k0 + 4 are loaded while the 16 DMMAs of
        // step k0 run (register double buffer) so the L2 latency of P hides behind the tensor pipe.
        const int nt_last = (chunk * QT + QT - 1) * n_warps + warp;
        if ((D & 3) == 0 && (nt_last + 1) * 8 <= D) {
            const double* pb0[QT];
#pragma unroll
            for (int q = 0; q < QT; ++q)
                pb0[q] = P + (size_t)k4 * D + ((chunk * QT + q) * n_warps + warp) * 8 + r8;
            const double* ra = rh_tile + (size_t)r8 * Dpad + k4;     // row tile m starts 8 m rows further down
            double bn[QT], an[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) an[m] = 8 * m + r8 < cpb ? ra[(size_t)8 * m * Dpad] : 0.0;
#pragma unroll
            for (int q = 0; q < QT; ++q) bn[q] = __ldg(pb0[q]);
            for (int k0 = 0; k0 < D; k0 += 4) {
                double bc[QT], ac[MT];
#pragma unroll
                for (int m = 0; m < MT; ++m) ac[m] = an[m];
#pragma unroll
                for (int q = 0; q < QT; ++q) bc[q] = bn[q];
                if (k0 + 4 < D) {
                    const size_t off = (size_t)(k0 + 4) * D;
#pragma unroll
                    for (int q = 0; q < QT; ++q) bn[q] = __ldg(pb0[q] + off);
#pragma unroll
                    for (int m = 0; m < MT; ++m) an[m] = 8 * m + r8 < cpb ? ra[(size_t)8 * m * Dpad + k0 + 4] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < QT; ++q)
#pragma unroll
                    for (int m = 0; m < MT; ++m) dmma_m8n8k4(acc[m][q][0], acc[m][q][1], ac[m], bc[q]);
            }
        } else {
            for (int k0 = 0; k0 < D; k0 += 4) {
                const int k = k0 + k4;
                const bool kin = k < D;
                double ac[MT];
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    ac[m] = (kin && 8 * m + r8 < cpb) ? rh_tile[(size_t)(8 * m + r8) * Dpad + k] : 0.0;
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    const int nt = (chunk * QT + q) * n_warps + warp;
                    const int n = nt * 8 + r8;             // B fragment: row k4, column r8 of the tile
                    const double b = (kin && n < D) ? __ldg(P + (size_t)k * D + n) : 0.0;
#pragma unroll
                    for (int m = 0; m < MT; ++m) dmma_m8n8k4(acc[m][q][0], acc[m][q][1], ac[m], b);
                }
            }
        }
        // C fragment: row r8, columns 2*k4 + {0,1} of tile nt.  Fold V into rho.V and theta.V.
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            const int nt = (chunk * QT + q) * n_warps + warp;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = nt * 8 + 2 * k4 + e;
                if (n < D) {
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const int row = 8 * m + r8;
                        if (row < cpb) {
                            const double v = acc[m][q][e];
                            pa[m] += rh_tile[(size_t)row * Dpad + n] * v;
                            pb[m] -= th_tile[(size_t)row * Dpad + n] * v;
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {                      // sum over the 4 lanes sharing a row
        pa[m] += __shfl_xor_sync(0xffffffffu, pa[m], 1);
        pa[m] += __shfl_xor_sync(0xffffffffu, pa[m], 2);
        pb[m] += __shfl_xor_sync(0xffffffffu, pb[m], 1);
        pb[m] += __shfl_xor_sync(0xffffffffu, pb[m], 2);
        if (k4 == 0) {
            red[warp][8 * m + r8][0] = pa[m];
            red[warp][8 * m + r8][1] = pb[m];
        }
    }
    __syncthreads();
    double A = 0, B = 0;
    for (int w = 0; w < n_warps; ++w) {                 // fixed order: deterministic
        A += red[w][o][0];
        B += red[w][o][1];
    }
    A_out = A;
    Bq_out = B;
    __syncthreads();                                    // red may be rewritten by the next draw
}

}  // namespace klhr
