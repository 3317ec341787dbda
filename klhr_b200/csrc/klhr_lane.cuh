// klhr_b200 -- the LANE kernel: diagonal-Gaussian targets (stan/normal.stan, stan/ill-normal.stan) with the
// Gaussian line family; chain state resident in shared memory for the whole launch, G lanes per chain.
//
// Why a third shape.  The tile kernel (klhr_tile.cuh) keeps theta in global memory (L2-resident) and pays an
// L2 round trip of 2 D reals per chain-draw plus 3-level shuffles, per-lane table loads and a generic Newton
// fit: 243 warp-instructions per chain-draw at 61 % issue utilisation (profiles/r01_tile_kernel.txt).  Here
//   * a CTA owns 32 chains for all n_steps draws of the launch; theta (fp64) and the current direction x (fp32)
//     live in shared memory, coordinate-pair major / chain minor, so that a lane reads and writes its own chain
//     with conflict-free 128-bit accesses -- HBM and L2 see one read and one write of theta per LAUNCH;
//   * G = 1: one thread per chain (a one-warp CTA); G = 2: two lanes (l, l + 16) per chain, each walking every
//     other 16-coordinate trip, a two-warp CTA -- twice the warps per SM for the same shared memory, which is
//     what hides the latencies of the three pipes the sweep alternates between (IMAD.WIDE of Philox, MUFU of
//     Box-Muller, FP64 of the sums);
//   * no shuffles in the sweep, and the per-coordinate tables (weights, scales) are warp-uniform broadcast
//     loads, two / four coordinates per instruction;
//   * the pending move of the previous draw (theta += c x_prev) is folded into the same sweep that draws the
//     new direction and forms the three sums ||x+tol||^2, x'Wx, theta'Wx (reference klhr.py:143-153 and the
//     line restriction of klhr.py:122-124);
//   * the sweep is software-pipelined by hand: the Philox blocks of trip t+2, the Box-Muller transforms of trip
//     t+1 and the FP64 work of trip t sit in ONE basic block, so the scheduler interleaves three independent
//     instruction streams bound to three different pipes (the pipeline runs across draws: counter-based streams
//     need nothing from the draw before);
//   * the line fit of a quadratic target is evaluated in closed form (quad_fit_closed below): on a concave
//     quadratic the fixed-budget optimiser of klhr_fit.cuh takes exactly one full Newton step in stage 1 and
//     zero steps in stage 2, so its result is m = xi0 + l'(xi0)/A, log s = -log(A)/2 (reference fit,
//     klhr.py:126-141).  Chains for which that premise fails (non-finite sums, a capped step, a start inside
//     the stage-1 tolerance) take the generic iteration instead, so both paths return the same iterates.
// Variate streams are the same function of (seed, chain, draw, element) as in the tile and octet kernels.
#pragma once
#include "klhr_tile.cuh"

namespace klhr {

#ifndef KLHR_LANE_G
#define KLHR_LANE_G 2
#endif
#ifndef KLHR_LANE_MINCTAS
#define KLHR_LANE_MINCTAS 5
#endif
constexpr int kLaneCta = 32;       // chains per CTA

// Closed-form line fit + proposal + MH ratio for a quadratic line restriction l(y) - l(0) = Bq y - A y^2 / 2.
// Returns false when the chain must take the generic path (see header).  `logu` = log of the accept uniform.
// Every operation is spelled as an explicit fma / mul / add so that all instantiations round identically.
__device__ __forceinline__ bool quad_fit_closed(double A, double Bq, double z_init, double z_prop, double logu,
                                                const FitParams& fp, bool want_eta, StepOut<double>& o) {
    const double xi0 = __dmul_rn(z_init, fp.initscale);
    const double nA = -A, nhA = __dmul_rn(-0.5, A);
    const double l10 = __fma_rn(nA, xi0, Bq);                    // l'(xi0)
    const double l0 = __dmul_rn(xi0, __fma_rn(nhA, xi0, Bq));
    const double gt1A = __dmul_rn(__dmul_rn(fp.gtol1, fp.gtol1), A);
    const double s = rsqrt(A);                                   // A^-1/2: the Laplace scale, and the KL optimum
    const double s2 = __dmul_rn(s, s);
    const double newton = __dmul_rn(l10, s2);                    // -l'/l''
    const double xi1 = __dadd_rn(xi0, newton);
    const double l11 = __fma_rn(nA, xi1, Bq);                    // l'(xi1): rounding noise
    const double l1v = __dmul_rn(xi1, __fma_rn(nhA, xi1, Bq));
    // premises of the one-step / zero-step iteration (stage1_mode, stage2_newton): concave and finite, start
    // not already converged, full Newton step inside the cap and improving, stage-1 converged at xi1, KL
    // gradient (-l'(m) s, A s^2 - 1) below gtol2 at (xi1, -log(A)/2)
    bool ok = (A > 0.0) && isfinite(A) && isfinite(l0) && isfinite(l1v);
    ok = ok && !(__dmul_rn(l10, l10) <= gt1A);
    ok = ok && (fabs(newton) <= kNewtonCapD) && (l1v > l0);
    ok = ok && (__dmul_rn(l11, l11) <= gt1A);
    ok = ok && (__dmul_rn(fabs(l11), s) <= fp.gtol2) && (fabs(__fma_rn(A, s2, -1.0)) <= fp.gtol2);
    if (!ok) return false;
    const double zp = __fma_rn(s, z_prop, xi1);                  // klhr.py:180
    const double is = __dmul_rn(A, s);                           // 1 / s
    const double z0 = __dmul_rn(-xi1, is), z1 = __dmul_rn(__dadd_rn(zp, -xi1), is);
    const double lq0 = __dmul_rn(__dmul_rn(-0.5, z0), z0), lq1 = __dmul_rn(__dmul_rn(-0.5, z1), z1);   // _logq up to the common -log s (klhr.py:155-158)
    const double lz = __dmul_rn(zp, __fma_rn(nhA, zp, Bq));
    const double r = __dadd_rn(__dadd_rn(isfinite(lz) ? lz : -Num<double>::inf(), lq0), -lq1);   // klhr.py:183-186 with lp(theta) = l(0)
    const double rm = r < 0.0 ? r : 0.0;
    o.accept = (r == r) && (logu < rm);
    o.eta[0] = xi1;
    o.eta[1] = want_eta ? __dmul_rn(-0.5, r_log(A)) : 0.0;
    o.zp = zp;
    o.r = r;
    o.converged = true;
    o.evals = 1 + kOct + fp.N + 2;                               // stage 1: start + 8 candidates; stage 2: one KL evaluation; MH: 2
    return true;
}

// trips of the sweep: trip t covers coordinates 32 k + j0 + jj + 8 r (k = t / 2, j0 = 4 (t % 2), jj, r in 0..3)
__host__ __device__ __forceinline__ int lane_n_trips(int D) {
    const int rem = D & 31;
    return 2 * (D >> 5) + (rem > 4 ? 2 : (rem > 0 ? 1 : 0));
}

template <bool kScaled, int G, bool kDraws, bool kTail>
__global__ void __launch_bounds__(kWarp * G, KLHR_LANE_MINCTAS) lane_kernel(const __grid_constant__ StepArgs a) {
    static_assert(!kTail || G == 2, "the split tail is the two-lanes-per-chain form");
    using R = double;
    using Model = DiagNormal<R, kScaled>;
    constexpr int NC = kWarp / G;                             // chains per warp
    constexpr int kPitch = NC + 1;                            // 16-byte slots per coordinate pair of theta (+1: transposed copies stay conflict-free)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.mp.D;
    const int Dp = (D + 3) & ~3;
    const int nq = Dp >> 2;                                   // coordinate quads
    const int wid = threadIdx.x >> 5, L = threadIdx.x & 31;
    const int cl = L & (NC - 1), p = L / NC;                  // chain slot in the warp, lane's part of the chain
    const int n_cols = a.dir.mean_cols ? a.dir.n_cols : 0;
    const int n_stored = n_cols ? n_cols - a.dir.n_zero_cols : 0;
    // shared memory: theta pairs [G][Dp/2][NC+1] double2 | x quads [G][Dp/4][NC] float4 | w pairs [Dp/2] double2 |
    //                sd quads [Dp/4] float4 | mean quads [n_stored + 1][Dp/4] float4 (last column = zeros) | cdf
    double2* th2 = reinterpret_cast<double2*>(smem_raw) + (size_t)wid * (Dp >> 1) * kPitch;
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<double2*>(smem_raw) + (size_t)G * (Dp >> 1) * kPitch) +
                 (size_t)wid * nq * NC;
    double2* s_w2 = reinterpret_cast<double2*>(reinterpret_cast<double2*>(smem_raw) + (size_t)G * (Dp >> 1) * kPitch +
                                               (size_t)G * nq * NC);
    float4* s_sd4 = reinterpret_cast<float4*>(s_w2 + (Dp >> 1));
    float4* s_mean4 = s_sd4 + nq;
    float* s_cdf = reinterpret_cast<float*>(s_mean4 + (size_t)(n_stored + 1) * nq);

    const long long tile0 = (long long)blockIdx.x * kLaneCta + wid * NC;     // first chain of this warp
    const long long c_own = tile0 + cl;
    const bool own_valid = c_own < a.B;
    R* g_theta = reinterpret_cast<R*>(a.theta);
    R* g_draws = kDraws ? reinterpret_cast<R*>(a.acc.draws) : nullptr;

    {   // tables (padded coordinates: weight 0, scale 0, mean 0)
        const R* g_w = reinterpret_cast<const R*>(a.mp.p0);
        const R* g_sd = reinterpret_cast<const R*>(a.dir.sd);
        const R* g_mean = reinterpret_cast<const R*>(a.dir.mean_cols);
        double* s_w = reinterpret_cast<double*>(s_w2);
        float* s_sd = reinterpret_cast<float*>(s_sd4);
        float* s_mean = reinterpret_cast<float*>(s_mean4);
        for (int i = threadIdx.x; i < Dp; i += kWarp * G) {
            s_w[i] = i < D ? (kScaled ? (double)g_w[i] : 1.0) : 0.0;
            s_sd[i] = i < D ? (g_sd ? (float)g_sd[i] : 1.0f) : 0.0f;
            for (int q = 0; q <= n_stored; ++q)
                s_mean[(size_t)q * Dp + i] = (q < n_stored && i < D) ? (float)g_mean[(size_t)q * D + i] : 0.0f;
        }
        for (int i = threadIdx.x; i < n_cols; i += kWarp * G)
            s_cdf[i] = n_cols > 1 ? (float)reinterpret_cast<const R*>(a.dir.cdf)[i] : 1.0f;
    }
    // theta tile of this warp: coalesced rows from global memory, transposed into [pair][chain]
    {
        double* th = reinterpret_cast<double*>(th2);
        for (int cc = 0; cc < NC; ++cc) {
            const long long c = tile0 + cc;
            for (int i = L; i < Dp; i += kWarp)
                th[((size_t)(i >> 1) * kPitch + cc) * 2 + (i & 1)] = (c < a.B && i < D) ? g_theta[c * D + i] : 0.0;
        }
        for (int q = p; q < nq; q += G) x4[(size_t)q * NC + cl] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if constexpr (G > 1) __syncthreads(); else __syncwarp();

    const R tol = (R)a.fp.tol;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const unsigned long long cid = (unsigned long long)(a.chain_offset + c_own);
    const uint32_t c0 = (uint32_t)cid, c1 = (uint32_t)(cid >> 32);
    R c_pend = 0;                                             // zp / ||x+tol|| of the last accepted draw, else 0
    long long n_acc = 0;
    unsigned long long n_evals = 0;
    double2* my_th = th2 + cl;
    float4* my_x = x4 + cl;
    // kTail (D mod 32 in 1..4, see lane_split_tail): the last trip would hold ONE coordinate quad -- 16 normals drawn
    // for <= 4 coordinates, on one lane of the chain while the other idles (D = 100: a quarter of the sweep's
    // instructions for 4 % of its coordinates).  That quad is taken out of the trip loop and split between the two
    // lanes of the chain (tail_* below); the loop then runs full trips only, the same number on both lanes.
    const int n_trips = lane_n_trips(D) - (kTail ? 1 : 0);
    const bool traced = a.tr.eta || a.tr.zp || a.tr.r || a.tr.accept || a.tr.evals || a.tr.rho || a.tr.z_init;

    // cooperative copy of the warp's states to global rows dst[c][0..D) (thinned draws, final write-back)
    auto store_tile = [&](R* dst_base) {
        const double* th = reinterpret_cast<const double*>(th2);
        for (int cc = 0; cc < NC; ++cc) {
            const long long c = tile0 + cc;
            if (c >= a.B) break;
            for (int i = L; i < D; i += kWarp)
                dst_base[c * D + i] = th[((size_t)(i >> 1) * kPitch + cc) * 2 + (i & 1)];
        }
    };

    // ---- the pipeline: Philox of trip n+2 | Box-Muller of trip n+1 | FP64 sweep of trip n (trips of THIS lane:
    //      t = p, p + G, ..., wrapping into the next draw)
    auto philox_trip = [&](int step, int t, uint32_t (&w)[4][4]) {
        const unsigned long long draw = (unsigned long long)(a.draw_offset + step);
        Philox::blockN<4>(c0, c1, (uint32_t)draw, kSlotDir + (uint32_t)(4 * (t & 1) + 8 * (t >> 1)), 1u, k0,
                          k1 ^ (uint32_t)(draw >> 32), w);
    };
    auto box_muller_trip = [&](const uint32_t (&w)[4][4], float (&z)[4][4]) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            box_muller_f32(w[jj][0], w[jj][1], z[jj][0], z[jj][1]);
            box_muller_f32(w[jj][2], w[jj][3], z[jj][2], z[jj][3]);
        }
    };
    uint32_t w_mid[4][4];
    float z_cur[4][4];
    int s2 = 0, t2 = p;                                       // cursor of the Philox stage
    auto advance = [&]() { t2 += G; if (t2 >= n_trips) { t2 = p; s2 += 1; } };
    const bool has_trips = p < n_trips;
    if (has_trips) {
        philox_trip(s2, t2, w_mid);
        box_muller_trip(w_mid, z_cur);
        advance();
        philox_trip(s2, t2, w_mid);
        advance();
    }
    // split tail: lane part p owns coordinates e_tail + 2 p, e_tail + 2 p + 1 = word 0 of the Philox blocks at slots
    // kSlotDir + 8 (D / 32) + 2 p (+ 1) -- the same (slot, word) -> element map as everywhere else (kSlotDir)
    const int q_tail = (D & ~31) >> 2;
    float zt0 = 0.f, zt1 = 0.f;                               // the two tail normals of the NEXT draw to sweep
    auto tail_normals = [&](int step) {
        const unsigned long long draw = (unsigned long long)(a.draw_offset + step);
        uint32_t wt[2][4];
        Philox::blockN<2>(c0, c1, (uint32_t)draw, kSlotDir + (uint32_t)(8 * (D >> 5) + 2 * p), 1u, k0,
                          k1 ^ (uint32_t)(draw >> 32), wt);
        zt0 = box_muller_f32_cos(wt[0][0], wt[0][1]);
        zt1 = box_muller_f32_cos(wt[1][0], wt[1][1]);
    };
    if constexpr (kTail) tail_normals(0);

    // per-draw scalar variates: lane part p draws them for draw step + p once every G draws (they depend on
    // nothing but the counters), the G lanes of a chain then read them from one another -- each chain-draw
    // pays for one set, not G
    R my_zi = 0, my_zp = 0, my_lu = 0, my_u = 0;
    int my_col = 0;
    for (int step = 0; step < a.n_steps; ++step) {
        const int jg = step & (G - 1);
        if (jg == 0) {
            const unsigned long long draw = (unsigned long long)(a.draw_offset + step + p);
            R u_col;
            chain_scalars<R>(c0, c1, (uint32_t)draw, k0, k1 ^ (uint32_t)(draw >> 32), u_col, my_zi, my_zp, my_u);
            my_lu = r_log(my_u);
            int jc = 0;
            if (n_cols > 1)                                   // searchsorted(cdf, u, 'right'), klhr.py:147
                while (jc < n_cols - 1 && (float)u_col >= s_cdf[jc]) ++jc;
            my_col = jc;
        }
        R z_init = my_zi, z_prop = my_zp, logu = my_lu, u = my_u;
        int jcol = my_col;
        if constexpr (G > 1) {
            const int src = cl + jg * NC;
            z_init = __shfl_sync(0xffffffffu, my_zi, src);
            z_prop = __shfl_sync(0xffffffffu, my_zp, src);
            logu = __shfl_sync(0xffffffffu, my_lu, src);
            jcol = __shfl_sync(0xffffffffu, my_col, src);
            u = __shfl_sync(0xffffffffu, my_u, src);
        }
        // a zero column (klhr.py:64-66) or no mean at all: the all-zero column stored last
        const float4* mcol = s_mean4 + (size_t)((n_cols && jcol < n_stored) ? jcol : n_stored) * nq;
        const R cp = c_pend;
        // ||x + tol||^2 = sum x^2 + tol (2 sum x + D tol): sum x^2 in fp64, sum x in fp32 (it is scaled by tol = 1e-12)
        R ss0 = 0, ss1 = 0, sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0;
        float sx0 = 0, sx1 = 0;

        auto trip = [&](auto full_tag, const int t) {
            constexpr bool kFull = decltype(full_tag)::value;
            const int e0 = 32 * (t >> 1) + 4 * (t & 1);
            // stage 2 of trip n+1 and stage 1 of trip n+2: independent of everything below
            float z_next[4][4];
            box_muller_trip(w_mid, z_next);
            philox_trip(s2, t2, w_mid);
            advance();
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int e = e0 + 8 * r;                     // first coordinate of the run (multiple of 4)
                if (!kFull && e >= D) break;                  // warp-uniform per part
                const int q = e >> 2;
                double2 ta = my_th[(size_t)(2 * q) * kPitch], tb = my_th[(size_t)(2 * q + 1) * kPitch];
                const float4 xo = my_x[(size_t)q * NC];
                const double2 wa = s_w2[2 * q], wb = s_w2[2 * q + 1];
                const float4 sd = s_sd4[q], mn = mcol[q];
                float4 xn;
                xn.x = fmaf(sd.x, z_cur[0][r], mn.x);
                xn.y = fmaf(sd.y, z_cur[1][r], mn.y);
                xn.z = fmaf(sd.z, z_cur[2][r], mn.z);
                xn.w = fmaf(sd.w, z_cur[3][r], mn.w);
                ta.x = fma(cp, (double)xo.x, ta.x);           // pending move of the previous draw
                ta.y = fma(cp, (double)xo.y, ta.y);
                tb.x = fma(cp, (double)xo.z, tb.x);
                tb.y = fma(cp, (double)xo.w, tb.y);
                const bool l1 = kFull || e + 1 < D, l2 = kFull || e + 2 < D, l3 = kFull || e + 3 < D;
                {
                    const R x = (double)xn.x, xw = kScaled ? x * wa.x : x;
                    sx0 += xn.x;
                    ss0 = fma(x, x, ss0); if (kScaled) sA0 = fma(x, xw, sA0); sB0 = fma(ta.x, xw, sB0);
                }
                if (l1) {
                    const R x = (double)xn.y, xw = kScaled ? x * wa.y : x;
                    sx1 += xn.y;
                    ss1 = fma(x, x, ss1); if (kScaled) sA1 = fma(x, xw, sA1); sB1 = fma(ta.y, xw, sB1);
                }
                if (l2) {
                    const R x = (double)xn.z, xw = kScaled ? x * wb.x : x;
                    sx0 += xn.z;
                    ss0 = fma(x, x, ss0); if (kScaled) sA0 = fma(x, xw, sA0); sB0 = fma(tb.x, xw, sB0);
                }
                if (l3) {
                    const R x = (double)xn.w, xw = kScaled ? x * wb.y : x;
                    sx1 += xn.w;
                    ss1 = fma(x, x, ss1); if (kScaled) sA1 = fma(x, xw, sA1); sB1 = fma(tb.y, xw, sB1);
                }
                my_th[(size_t)(2 * q) * kPitch] = ta;
                my_th[(size_t)(2 * q + 1) * kPitch] = tb;
                my_x[(size_t)q * NC] = xn;
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int r = 0; r < 4; ++r) z_cur[jj][r] = z_next[jj][r];
        };
#pragma unroll 1
        for (int t = p; t < n_trips; t += G) {
            if (kTail || 32 * (t >> 1) + 4 * (t & 1) + 27 < D) trip(std::true_type{}, t);
            else trip(std::false_type{}, t);
        }
        if constexpr (kTail) {
            // padded coordinates (>= D) carry weight, scale, mean, theta and x = 0: they add exact zeros
            double2 tt = my_th[(size_t)(2 * q_tail + p) * kPitch];
            float2* xp = reinterpret_cast<float2*>(&my_x[(size_t)q_tail * NC]) + p;
            const float2 xo = *xp;
            const double2 wt2 = s_w2[2 * q_tail + p];
            const float2 sd = reinterpret_cast<const float2*>(&s_sd4[q_tail])[p];
            const float2 mn = reinterpret_cast<const float2*>(&mcol[q_tail])[p];
            float2 xn;
            xn.x = fmaf(sd.x, zt0, mn.x);
            xn.y = fmaf(sd.y, zt1, mn.y);
            tt.x = fma(cp, (double)xo.x, tt.x);               // pending move of the previous draw
            tt.y = fma(cp, (double)xo.y, tt.y);
            {
                const R x = (double)xn.x, xw = kScaled ? x * wt2.x : x;
                sx0 += xn.x;
                ss0 = fma(x, x, ss0); if (kScaled) sA0 = fma(x, xw, sA0); sB0 = fma(tt.x, xw, sB0);
            }
            {
                const R x = (double)xn.y, xw = kScaled ? x * wt2.y : x;
                sx1 += xn.y;
                ss1 = fma(x, x, ss1); if (kScaled) sA1 = fma(x, xw, sA1); sB1 = fma(tt.y, xw, sB1);
            }
            my_th[(size_t)(2 * q_tail + p) * kPitch] = tt;
            *xp = xn;
        }
        __syncwarp();
        if constexpr (kDraws) {
            // the sweep has just applied the move of draw g = thin_offset + step: the state after it is complete
            const long long g_prev = a.acc.thin_offset + step;
            if (step > 0 && g_prev % a.acc.thin == 0) {
                store_tile(g_draws + (g_prev / a.acc.thin - 1) * a.B * D);
                __syncwarp();
            }
        }
        // ------------------------------------------------------------------ fit, proposal, MH (every lane of the chain)
        if constexpr (kTail) tail_normals(step + 1);          // integer work that depends on nothing: it fills the
                                                              // issue slots under the fit's dependent FP64 chain
        R ss = ss0 + ss1, sA = sA0 + sA1, sB = sB0 + sB1;
        float sx = sx0 + sx1;
#pragma unroll
        for (int off = NC; off < kWarp; off <<= 1) {          // a + b == b + a: all lanes of a chain end with identical sums
            ss += __shfl_xor_sync(0xffffffffu, ss, off);
            if (kScaled) sA += __shfl_xor_sync(0xffffffffu, sA, off);
            sB += __shfl_xor_sync(0xffffffffu, sB, off);
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
        }
        if (!kScaled) sA = ss;                                // unit weights: x'Wx = sum x^2
        ss = fma(tol, fma(2.0, (double)sx, (double)D * tol), ss);
        const R inv = rsqrt(ss);                              // rho = x / ||x + tol||  (klhr.py:153)
        typename Model::Coef cf;
        cf.A = __dmul_rn(__dmul_rn(sA, inv), inv);
        cf.Bq = __dmul_rn(-sB, inv);
        StepOut<R> so;
        const bool want_eta = a.tr.eta != nullptr;
        if (!quad_fit_closed(cf.A, cf.Bq, z_init, z_prop, logu, a.fp, want_eta, so)) {
            OrCtx<R> oc;
            oc.K = 0; oc.inject = false; oc.r = 0; oc.v = 1;
            fit_and_propose<1, R, Model, 2>(cf, a.fp, 0, 0u, z_init, R(0), R(0), z_prop, u, so, oc);
        }
        c_pend = so.accept ? __dmul_rn(so.zp, inv) : R(0);
        n_acc += so.accept ? 1 : 0;
        n_evals += (unsigned long long)so.evals;
        // traces (tests): runtime-checked so that traced and untraced launches run the very same arithmetic
        if (traced) {
            if (own_valid && p == 0) {
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
                if (a.tr.z_init) {
                    reinterpret_cast<R*>(a.tr.z_init)[trow] = z_init;
                    reinterpret_cast<R*>(a.tr.z_prop)[trow] = z_prop;
                    reinterpret_cast<R*>(a.tr.u)[trow] = u;
                }
                if (a.tr.rho) {                               // rho = x * inv (test / debugging path)
                    R* g = reinterpret_cast<R*>(a.tr.rho) + trow * D;
                    const float* xf = reinterpret_cast<const float*>(x4);
                    for (int i = 0; i < D; ++i) g[i] = __dmul_rn((double)xf[((size_t)(i >> 2) * NC + cl) * 4 + (i & 3)], inv);
                }
            }
        }
    }
    // ------------------------------------------------------------------------ flush the pending move, write back
    {
        const float* xf = reinterpret_cast<const float*>(x4);
        double* th = reinterpret_cast<double*>(th2);
        for (int i = p; i < D; i += G) {
            double* t = th + ((size_t)(i >> 1) * kPitch + cl) * 2 + (i & 1);
            *t = fma(c_pend, (double)xf[((size_t)(i >> 2) * NC + cl) * 4 + (i & 3)], *t);
        }
    }
    __syncwarp();
    store_tile(g_theta);
    if constexpr (kDraws) {                                   // the last draw of the launch, if it is a kept one
        const long long g_last = a.acc.thin_offset + a.n_steps;
        if (a.n_steps > 0 && g_last % a.acc.thin == 0) store_tile(g_draws + (g_last / a.acc.thin - 1) * a.B * D);
    }
    if (own_valid && p == 0 && a.acc.accept_count) a.acc.accept_count[c_own] += n_acc;
    if (a.acc.evals_total) {
        unsigned long long tot = (own_valid && p == 0) ? n_evals : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
        if (L == 0 && tot) atomicAdd(a.acc.evals_total, tot);
    }
}

__host__ inline size_t lane_smem_bytes(const StepArgs& a, int G = KLHR_LANE_G) {
    const int D = a.mp.D, Dp = (D + 3) & ~3;
    const int n_cols = a.dir.mean_cols ? a.dir.n_cols : 0;
    const int n_stored = n_cols ? n_cols - a.dir.n_zero_cols : 0;
    size_t b = (size_t)(Dp / 2) * (kLaneCta + G) * 16;        // theta
    b += (size_t)(Dp / 4) * kLaneCta * 16;                    // x
    b += (size_t)Dp * 8 + (size_t)Dp * 4;                     // w, sd
    b += (size_t)(n_stored + 1) * Dp * 4;                     // mean columns + the zero column
    b += (size_t)((n_cols + 3) & ~3) * 4;                     // cdf
    return b;
}

// dimensions whose last trip is a single coordinate quad (see kTail in lane_kernel)
__host__ __device__ __forceinline__ bool lane_split_tail(int D) { return (D & 31) >= 1 && (D & 31) <= 4; }

template <bool kScaled, int G, bool kTail>
int launch_lane_typed(const StepArgs& a, cudaStream_t st, LaunchInfo* info) {
    const size_t smem = lane_smem_bytes(a, G);
    if (smem > 227 * 1024) return -20;
    const void* fn = a.acc.draws ? (const void*)lane_kernel<kScaled, G, true, kTail>
                                 : (const void*)lane_kernel<kScaled, G, false, kTail>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    if (info) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, fn);
        if (e != cudaSuccess) return (int)e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kWarp * G, smem);
        if (e != cudaSuccess) return (int)e;
        info->threads = kWarp * G;
        info->smem = (int)smem;
        info->regs = fa.numRegs;
        info->ctas_per_sm = nb;
        return 0;
    }
    const long long grid = (a.B + kLaneCta - 1) / kLaneCta;
    if (grid <= 0) return 0;
    void* kargs[] = {(void*)&a};
    e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kWarp * G), kargs, smem, st);
    return (int)e;
}

// defined in klhr_lane.cu
int launch_lane(const StepArgs& a, cudaStream_t st, LaunchInfo* info);

}  // namespace klhr
