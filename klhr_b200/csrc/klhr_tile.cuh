// klhr_b200 -- the TILE kernel: the fast path for diagonal-Gaussian targets
// (stan/normal.stan, stan/ill-normal.stan) with the Gaussian line family.
//
// One warp owns a tile of 32 chains and alternates between two shapes per draw:
//   D-phase   8 passes; in pass p the 4 octets of the warp each process ONE chain (slot 4p+o)
//             cooperatively: theta row streamed from global memory (L2-resident between
//             draws), the pending update of the previous draw applied on the fly
//             (theta += c * x_prev, written back only if the draw was accepted), Philox ->
//             Box-Muller -> x = mean + sd z kept as fp32 in shared memory, and the three sums
//             ||x+tol||^2, sum x^2 w, sum x theta w reduced with 3-level octet shuffles;
//   fit-phase thread-per-chain: lane 8o+j owns chain slot 4j+o, turns the sums into the line
//             coefficients (A, Bq) of klhr_models.cuh:QuadCoef and runs stage 1 / stage 2 /
//             proposal / MH of klhr_fit.cuh with G = 1 -- no shuffles, no redundant lanes.
// Compared with the octet kernel (klhr_step.cuh) the per-chain scalar work is done once
// instead of eight times and theta is read once and written at most once per draw.
//
// Variate streams are the same function of (seed, chain, draw, element) as in the octet
// kernel, so both kernels produce the same chains up to fp32 rounding of the direction.
#pragma once
#include <type_traits>
#include "klhr_step.cuh"

namespace klhr {

// Tuning knobs (measured on B200, ill-normal D = 100, 65 536 chains):
//   KLHR_TILE_PASSES  chains per octet per draw; a warp owns 4 * passes chains.  8 (32-chain tiles, 13.8
//                     one-warp CTAs per SM) beat 6 and 4: smaller tiles need <= 96 / 72 registers and spill.
//   KLHR_TILE_MINCTAS 16 -> at most 128 registers, so that 14 tiles per SM are resident in ONE wave.
//   KLHR_TILE_W_SMEM  per-coordinate weights staged in shared memory (1) or read through L1 (0).
#ifndef KLHR_TILE_PASSES
#define KLHR_TILE_PASSES 8
#endif
#ifndef KLHR_TILE_MINCTAS
#define KLHR_TILE_MINCTAS 16
#endif
#ifndef KLHR_TILE_W_SMEM
#define KLHR_TILE_W_SMEM 1
#endif
constexpr int kTilePasses = KLHR_TILE_PASSES;
constexpr int kTileChains = 4 * kTilePasses;
constexpr int kWarp = 32;

// Octet shuffles with the FULL-warp mask: the whole warp reaches them together in the tile kernel, and a constant
// full mask spares the MATCH / REDUX / BRA.DIV preamble the compiler emits for a per-octet mask (5 % of the
// kernel's stall samples).
template <typename R>
__device__ __forceinline__ R warp_oct_sum(R v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
template <typename R>
__device__ __forceinline__ R warp_oct_bcast(R v, int src) { return __shfl_sync(0xffffffffu, v, src, kOct); }

// per-draw scalar variates of one chain (slots 0..2), computed by the owning thread
template <typename R>
__device__ __forceinline__ void chain_scalars(uint32_t c0, uint32_t c1, uint32_t d0, uint32_t k0, uint32_t k1d,
                                              R& u_col, R& z_init, R& z_prop, R& u) {
    uint32_t w[4];
    Philox::block(c0, c1, d0, kSlotScalarA, k0, k1d, w);
    float f0, f1;
    box_muller_f32(w[2], w[3], f0, f1);
    u_col = (R)u01_32(w[0]);
    z_init = (R)f0;
    Philox::block(c0, c1, d0, kSlotProposal, k0, k1d, w);
    if (sizeof(R) == 8) {
        const double z0 = box_muller_f64(u01_53(w[0], w[1]), u01_53(w[2], w[3]));
        z_prop = (R)z0;
    } else {
        box_muller_f32(w[0], w[2], f0, f1);
        z_prop = (R)f0;
    }
    Philox::block(c0, c1, d0, kSlotAccept, k0, k1d, w);
    u = sizeof(R) == 8 ? (R)u01_53(w[0], w[1]) : (R)u01_32(w[0]);
}

// kDraws: thinned-draw output (MCMCBase.sample rows) -- a separate instantiation so that the plain run()
// kernel carries none of it (the extra variants cost the hot path 6 % through register allocation otherwise)
template <typename R, bool kScaled, typename XT, bool kReplay, bool kDraws>
__global__ void __launch_bounds__(kWarp, KLHR_TILE_MINCTAS) tile_kernel(const __grid_constant__ StepArgs a) {
    using Model = DiagNormal<R, kScaled>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.mp.D;
    const int Dx = a.Dpad;                               // row stride of xs
    const int L = threadIdx.x;
    const int o = L >> 3, j = L & 7;
    const unsigned om = oct_mask();
    // shared memory: w[D] (R) | xs[32][Dx] (XT) | sd[D] (float) | cdf[n_cols] (float) | mean[n_stored][D] (float).
    // Every byte counts: 14 one-warp CTAs must fit per SM (<= 15.6 KB each) for the 2048 tiles of
    // B = 65 536 to be resident in one wave; the dispatcher (klhr_api.cu:tile_applies) therefore sends
    // direction laws with more than two stored mean columns to the octet kernel.
    R* s_w = reinterpret_cast<R*>(smem_raw);
    XT* xs = reinterpret_cast<XT*>(s_w + (KLHR_TILE_W_SMEM ? ((D + 1) & ~1) : 0));
    float* s_sd = reinterpret_cast<float*>(xs + (size_t)kTileChains * Dx);
    const int n_cols = (!kReplay && a.dir.mean_cols) ? a.dir.n_cols : 0;
    const int n_stored = n_cols ? n_cols - a.dir.n_zero_cols : 0;
    float* s_cdf = s_sd + D;
    float* s_mean = s_cdf + n_cols;
    const R* g_mean = reinterpret_cast<const R*>(a.dir.mean_cols);

    const long long tile0 = (long long)blockIdx.x * kTileChains;
    const long long c_own = tile0 + 4 * j + o;           // the chain this thread owns in the fit phase
    const bool own_valid = j < kTilePasses && c_own < a.B;
    R* g_theta = reinterpret_cast<R*>(a.theta);
    R* g_draws = kDraws ? reinterpret_cast<R*>(a.acc.draws) : nullptr;
    const R* g_w = reinterpret_cast<const R*>(a.mp.p0);

    if (KLHR_TILE_W_SMEM)
        for (int i = L; i < D; i += kWarp) s_w[i] = kScaled ? g_w[i] : R(1);
    if constexpr (!kReplay) {
        const R* g_sd = reinterpret_cast<const R*>(a.dir.sd);
        for (int i = L; i < D; i += kWarp) s_sd[i] = g_sd ? (float)g_sd[i] : 1.0f;
        if (n_cols > 1)
            for (int i = L; i < n_cols; i += kWarp) s_cdf[i] = (float)reinterpret_cast<const R*>(a.dir.cdf)[i];
        for (int i = L; i < n_stored * D; i += kWarp) s_mean[i] = (float)g_mean[i];
    }
    __syncwarp();

    const R tol = (R)a.fp.tol;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    R c_pend = 0;                                        // zp / ||x+tol|| of the last accepted draw, else 0
    long long n_acc = 0;
    unsigned long long n_evals = 0;

    for (int step = 0; step < a.n_steps; ++step) {
        const unsigned long long draw = (unsigned long long)(a.draw_offset + step);
        const uint32_t d0 = (uint32_t)draw, k1d = k1 ^ (uint32_t)(draw >> 32);
        R u_col = 0, z_init = 0, z_prop = 0, u = 0;
        int jcol = 0;
        if constexpr (kReplay) {
            if (own_valid) {
                z_init = reinterpret_cast<const R*>(a.z_init)[c_own];
                z_prop = reinterpret_cast<const R*>(a.z_prop)[c_own];
                u = reinterpret_cast<const R*>(a.u)[c_own];
            }
        } else {
            const unsigned long long cid = (unsigned long long)(a.chain_offset + c_own);
            chain_scalars<R>((uint32_t)cid, (uint32_t)(cid >> 32), d0, k0, k1d, u_col, z_init, z_prop, u);
            if (n_cols > 1)                               // searchsorted(cdf, u, 'right'), klhr.py:147
                while (jcol < n_cols - 1 && (float)u_col >= s_cdf[jcol]) ++jcol;
        }
        R my_ss = 1, my_A = 0, my_B = 0;
        // thinned output (MCMCBase.sample rows): the state after draw g = thin_offset + step (1-based) is
        // formed on the fly in the D-phase of the NEXT draw, so that is where it is written
        const long long g_prev = a.acc.thin_offset + step;
        const bool emit_rt = kDraws && step > 0 && g_prev % a.acc.thin == 0;
        const long long emit_row = emit_rt ? g_prev / a.acc.thin - 1 : 0;
        // -------------------------------------------------------------------- D-phase
        R cp_next = warp_oct_bcast(c_pend, 0);
        int col_next = warp_oct_bcast(jcol, 0);
#pragma unroll 1
        for (int p = 0; p < kTilePasses; ++p) {
            const int cs = 4 * p + o;
            const long long c = tile0 + cs;
            const R cp = cp_next;                          // pending move of this pass's chain
            const int col = col_next;
            cp_next = warp_oct_bcast(c_pend, (p + 1) % kTilePasses);  // one pass ahead: hides the shuffle latency
            col_next = warp_oct_bcast(jcol, (p + 1) % kTilePasses);
            R ss = 0, sA = 0, sB = 0;
            if (c < a.B) {                                // octet-uniform; absent chains contribute zeros
                R* row = g_theta + c * D;
                XT* xr = xs + (size_t)cs * Dx;
                const bool has_mean = col < n_stored;         // a zero column (klhr.py:64-66) is not stored
                const float* mcol_s = s_mean + (size_t)(has_mean ? col : 0) * D;
                const unsigned long long cid = (unsigned long long)(a.chain_offset + c);
                const uint32_t c0 = (uint32_t)cid, c1 = (uint32_t)(cid >> 32);
                const bool pend_rt = cp != R(0);
                R* drow = emit_rt ? g_draws + ((emit_row * a.B + c) * D) : nullptr;
                // Element i = g0 + j + 8 s, slot s = 4 t + r of this lane  <->  Philox counter slot
                // kSlotDir + j + 8 t + g0 / 4, word r.  A half handles slots 8h..8h+7 (blocks 2h, 2h+1).  When
                // the half is FULL (every slot live for every lane) the body carries no predicates at all;
                // only the last, partial half pays for bounds checks and skips its dead slots.
                auto do_half = [&](auto full_tag, auto pend_tag, auto emit_tag, const int g0, const int h) {
                    constexpr bool kFull = decltype(full_tag)::value;
                    constexpr bool pend = decltype(pend_tag)::value;   // a move of the previous draw is pending
                    constexpr bool emit = decltype(emit_tag)::value;   // the previous draw goes to the thinned output
                    // 1. issue the long-latency loads first (theta slice from L2, previous x); they hide
                    //    behind the RNG arithmetic of step 2.  Weights, scales and means are read from
                    //    shared memory at the point of use (preloading them needs > 128 registers).
                    R th[8];
                    XT xo[8];
    #pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int i = g0 + j + 8 * (8 * h + s);
                        const bool live = kFull || i < D;
                        th[s] = live ? row[i] : R(0);
                        xo[s] = (pend && live) ? xr[i] : XT(0);
                    }
                    // 2. 8 normals: two Philox blocks advanced in lockstep, four Box-Muller pairs
                    float z[8];
                    if constexpr (!kReplay) {
                        uint32_t w[2][4];
                        Philox::blockN<2>(c0, c1, d0, kSlotDir + (uint32_t)(j + 16 * h) + (uint32_t)(g0 / 4), 8u,
                                          k0, k1d, w);
    #pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            box_muller_f32(w[t][0], w[t][1], z[4 * t + 0], z[4 * t + 1]);
                            box_muller_f32(w[t][2], w[t][3], z[4 * t + 2], z[4 * t + 3]);
                        }
                    }
                    // 3. apply the pending move, form the new x and the three sums
    #pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int ib = g0 + 8 * (8 * h + s);
                        if (!kFull && ib >= D) continue;      // warp-uniform: dead slot of the partial half
                        const int i = ib + j;
                        if (kFull || i < D) {
                            R t0 = th[s];
                            if (pend) {
                                t0 = t0 + cp * (R)xo[s];
                                row[i] = t0;
                            }
                            if (emit) drow[i] = t0;           // the state after the previous draw (moved or not)
                            R x;
                            if constexpr (kReplay) {
                                x = reinterpret_cast<const R*>(a.rho)[c * D + i];
                                xr[i] = (XT)x;
                            } else {
                                const float xf = fmaf(s_sd[i], z[s], has_mean ? mcol_s[i] : 0.0f);
                                xr[i] = (XT)xf;
                                x = (R)xf;
                            }
                            const R xt = x + tol;
                            ss += xt * xt;
                            const R xw = x * (KLHR_TILE_W_SMEM ? s_w[i] : Model::wgt(i, a.mp));
                            sA += x * xw;
                            sB += t0 * xw;
                        }
                    }
                };
                for (int g0 = 0; g0 < D; g0 += 128) {
    #pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        if (g0 + 64 * h >= D) break;          // warp-uniform: nothing left
                        const bool full = g0 + 64 * h + 63 < D;
                        if constexpr (kDraws) {
                            if (emit_rt) {                    // warp-uniform (thinned sample() output)
                                if (pend_rt) {
                                    if (full) do_half(std::true_type{}, std::true_type{}, std::true_type{}, g0, h);
                                    else do_half(std::false_type{}, std::true_type{}, std::true_type{}, g0, h);
                                } else {
                                    if (full) do_half(std::true_type{}, std::false_type{}, std::true_type{}, g0, h);
                                    else do_half(std::false_type{}, std::false_type{}, std::true_type{}, g0, h);
                                }
                                continue;
                            }
                        }
                        if (pend_rt) {
                            if (full) do_half(std::true_type{}, std::true_type{}, std::false_type{}, g0, h);
                            else do_half(std::false_type{}, std::true_type{}, std::false_type{}, g0, h);
                        } else {
                            if (full) do_half(std::true_type{}, std::false_type{}, std::false_type{}, g0, h);
                            else do_half(std::false_type{}, std::false_type{}, std::false_type{}, g0, h);
                        }
                    }
                }
            }
            ss = warp_oct_sum(ss);
            sA = warp_oct_sum(sA);
            sB = warp_oct_sum(sB);
            if (j == p && c < a.B) { my_ss = ss; my_A = sA; my_B = sB; }
        }
        // -------------------------------------------------------------------- fit phase (thread per chain)
        R inv = 1;
        if (own_valid) {
            typename Model::Coef cf;
            if constexpr (kReplay) {                      // injected rho is used as given (already normalised)
                cf.A = my_A;
                cf.Bq = -my_B;
            } else {
                inv = r_rsqrt(my_ss);                     // rho = x / ||x + tol||  (klhr.py:153)
                cf.A = my_A * inv * inv;
                cf.Bq = -my_B * inv;
            }
            StepOut<R> so;
            OrCtx<R> oc;
            oc.K = a.fp.or_K;
            oc.inject = kReplay;
            oc.r = 0;
            oc.v = 1;
            if (oc.K > 0) {
                if constexpr (kReplay) {
                    oc.r = a.tr.or_r[c_own];
                    oc.v = reinterpret_cast<const R*>(a.tr.or_v)[c_own];
                } else {
                    const unsigned long long ocid = (unsigned long long)(a.chain_offset + c_own);
                    const unsigned long long odraw = (unsigned long long)(a.draw_offset + step);
                    oc.c0 = (uint32_t)ocid; oc.c1 = (uint32_t)(ocid >> 32); oc.d0 = (uint32_t)odraw;
                    oc.k0 = (uint32_t)a.seed; oc.k1d = (uint32_t)(a.seed >> 32) ^ (uint32_t)(odraw >> 32);
                }
            }
            fit_and_propose<1, R, Model, 2>(cf, a.fp, 0, 0u, z_init, R(0), R(0), z_prop, u, so, oc);
            if (!kReplay && oc.K > 0 && a.tr.or_r) {
                a.tr.or_r[(long long)step * a.B + c_own] = oc.r;
                reinterpret_cast<R*>(a.tr.or_v)[(long long)step * a.B + c_own] = oc.v;
            }

            c_pend = so.accept ? so.zp * inv : R(0);
            n_acc += so.accept ? 1 : 0;
            n_evals += (unsigned long long)so.evals;
            const long long trow = (long long)step * a.B + c_own;
            if (a.tr.eta) {
                R* e = reinterpret_cast<R*>(a.tr.eta) + trow * 2;
                e[0] = so.eta[0];
                e[1] = so.eta[1];
            }
            if (a.tr.zp) reinterpret_cast<R*>(a.tr.zp)[trow] = so.zp;
            if (a.tr.r) reinterpret_cast<R*>(a.tr.r)[trow] = so.r;
            if (a.tr.accept) a.tr.accept[trow] = so.accept ? 1 : 0;
            if (a.tr.evals) a.tr.evals[trow] = so.evals;
            if (!kReplay && a.tr.z_init) {
                reinterpret_cast<R*>(a.tr.z_init)[trow] = z_init;
                reinterpret_cast<R*>(a.tr.z_prop)[trow] = z_prop;
                reinterpret_cast<R*>(a.tr.u)[trow] = u;
            }
        } else {
            c_pend = 0;
        }
        if (a.tr.rho) {                                   // emit rho = x * inv (test / debugging path)
            for (int p = 0; p < kTilePasses; ++p) {
                const int cs = 4 * p + o;
                const long long c = tile0 + cs;
                const R ip = oct_bcast(inv, p, om);
                if (c >= a.B) continue;
                R* g = reinterpret_cast<R*>(a.tr.rho) + ((long long)step * a.B + c) * D;
                for (int i = j; i < D; i += kOct) g[i] = (R)xs[(size_t)cs * Dx + i] * ip;
            }
        }
        __syncwarp();
    }
    // ------------------------------------------------------------------------ flush pending moves
    for (int p = 0; p < kTilePasses; ++p) {
        const int cs = 4 * p + o;
        const long long c = tile0 + cs;
        const R cp = oct_bcast(c_pend, p, om);
        if (c >= a.B || cp == R(0)) continue;
        R* row = g_theta + c * D;
        const XT* xr = xs + (size_t)cs * Dx;
        for (int i = j; i < D; i += kOct) row[i] = row[i] + cp * (R)xr[i];
    }
    if constexpr (kDraws) {                               // the last draw of the launch, if it is a kept one
        const long long g_last = a.acc.thin_offset + a.n_steps;
        if (a.n_steps > 0 && g_last % a.acc.thin == 0) {
            __syncwarp();
            for (int p = 0; p < kTilePasses; ++p) {
                const long long c = tile0 + 4 * p + o;
                if (c >= a.B) continue;
                const R* row = g_theta + c * D;
                R* drow = g_draws + ((g_last / a.acc.thin - 1) * a.B + c) * D;
                for (int i = j; i < D; i += kOct) drow[i] = row[i];
            }
        }
    }
    if (own_valid) {
        if (a.acc.accept_count) a.acc.accept_count[c_own] += n_acc;
    }
    if (a.acc.evals_total) {
        unsigned long long tot = own_valid ? n_evals : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
        if (L == 0 && tot) atomicAdd(a.acc.evals_total, tot);
    }
}

template <typename R, bool kScaled>
int launch_tile_typed(const StepArgs& args_in, bool replay, cudaStream_t st, LaunchInfo* info) {
    StepArgs a = args_in;
    const int xbytes = replay ? (int)sizeof(R) : 4;
    a.Dpad = pad_dim(a.mp.D, xbytes);
    const int n_cols = (!replay && a.dir.mean_cols) ? a.dir.n_cols : 0;
    const int n_stored = n_cols ? n_cols - a.dir.n_zero_cols : 0;
    size_t smem = (KLHR_TILE_W_SMEM ? (size_t)((a.mp.D + 1) & ~1) * sizeof(R) : 0) + (size_t)kTileChains * a.Dpad * xbytes +
                  (size_t)(a.mp.D + n_cols) * sizeof(float);
    smem += (size_t)n_stored * a.mp.D * sizeof(float);
    if (smem > 227 * 1024) return -20;
    const void* fn = replay ? (const void*)tile_kernel<R, kScaled, R, true, false>
                            : (a.acc.draws ? (const void*)tile_kernel<R, kScaled, float, false, true>
                                           : (const void*)tile_kernel<R, kScaled, float, false, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    // one warp per CTA: ask for the largest shared-memory carve-out so that 14+ tiles fit per SM
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    if (info) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, fn);
        if (e != cudaSuccess) return (int)e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kWarp, smem);
        if (e != cudaSuccess) return (int)e;
        info->threads = kWarp;
        info->smem = (int)smem;
        info->regs = fa.numRegs;
        info->ctas_per_sm = nb;
        return 0;
    }
    const long long grid = (a.B + kTileChains - 1) / kTileChains;
    if (grid <= 0) return 0;
    void* kargs[] = {(void*)&a};
    e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kWarp), kargs, smem, st);
    return (int)e;
}

// defined in klhr_tile.cu
int launch_tile(const StepArgs& a, int dtype, bool replay, cudaStream_t st, LaunchInfo* info);

}  // namespace klhr
