// klhr_b200 -- the WARP-SPECIALISED dense kernel: free-running launches of corr-normal D = 128 / 256 (fp64, Gaussian
// family).  Same algorithm, same variate streams and the same tensor-core product as dense_kernel (klhr_densek.cuh),
// re-timed: there, one set of 8 warps alternates between the tensor-core phase (46 % of a draw) and the scalar phases
// (direction 24 %, move 18 %, fit 7 %), and since a warp issues in order its own scalar work cannot hide under its own
// DMMAs.  Here the roles are split over warps, which the schedulers interleave for free:
//   * 8 TENSOR warps per CTA: V(t) = X(t) L, fold (A, Bq partial sums), w += c(t) V(t).  Their w = L' theta lives in shared
//     memory in the fragment layout's home positions (each thread only ever touches its own 32 entries, and only
//     between the product and the update), so that the product loop has the registers to keep loads, converts and
//     DMMAs of neighbouring k-steps in flight;
//   * 8 PRODUCER warps (8 lanes per chain): under the product on X(t) they apply the move of draw t-1 (theta += c x, theta
//     in global memory / L2: nothing else on the device reads it), draw X(t+1) and the scalar variates of draw t+1
//     (Philox + Box-Muller) into the buffer the move has just released; the first of them also runs the closed-form fit
//     of draw t (one thread per chain) as soon as the partial sums are in.
// X is fp32 (the directions ARE fp32 values: fmaf(sd, z, mean)) in two buffers.  One CTA-wide barrier per draw hands the
// buffers over; two partial barriers (bar.arrive / bar.sync on 288 threads) pass the sums to the fit and c back.
#pragma once
#include "klhr_densek.cuh"

namespace klhr {

#ifndef KLHR_WS_PRODUCER_LANES
#define KLHR_WS_PRODUCER_LANES 8
#endif
constexpr int kWsLpc = KLHR_WS_PRODUCER_LANES;     // producer lanes per chain: 4 -> 4 producer warps (168 registers per thread), 8 -> 8 (128)
constexpr int kWsTensor = 256, kWsProducer = 32 * kWsLpc, kWsThreads = kWsTensor + kWsProducer;
static_assert(kWsLpc == 4 || kWsLpc == 8, "producer lanes per chain: 4 or 8");
constexpr int kWsDepth = 4;                        // k-PAIRS (8 rows of L) in flight per tensor warp
constexpr int kWsBuf = 2;

__device__ __forceinline__ void ws_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void ws_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NT>
__global__ void __launch_bounds__(kWsThreads, 1) dense_ws_kernel(const __grid_constant__ StepArgs a) {
    using R = double;
    using Model = CorrNormal<R>;
    constexpr int CH = 32, MT = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.mp.D;
    const int S = D + 4;                               // pitch of X (floats) and of the theta staging rows (doubles): conflict-free A fragments
    const int SW = D + 8;                              // pitch of w (doubles): conflict-free 128-bit fragment accesses
    const int tid = threadIdx.x;
    const bool tensor = tid < kWsTensor;
    const int n_cols = a.dir.mean_cols ? a.dir.n_cols : 0;
    const int n_stored = n_cols ? n_cols - a.dir.n_zero_cols : 0;
    // shared memory: w[CH][SW] (double) | ring[8][depth][NT][32] | red[8][CH][2] | s_c[2][CH] | var[2][5][CH]
    //   | (floats) xf[2][CH][S] (= one [CH][S] double tile while w is initialised) | sd[D] | mean[n_stored][D] | cdf[n_cols]
    R* w_all = reinterpret_cast<R*>(smem_raw);
    R* ring_all = w_all + (size_t)CH * SW;
    R* red = ring_all + (size_t)8 * kWsDepth * NT * 64;
    R* s_c = red + 8 * CH * 2;
    R* s_var = s_c + 2 * CH;                           // per buffer: inv, z_init, z_prop, log u, u
    float* xf_all = reinterpret_cast<float*>(s_var + kWsBuf * 5 * CH);
    float* s_sd = xf_all + (size_t)kWsBuf * CH * S;
    float* s_mean = s_sd + D;
    float* s_cdf = s_mean + (size_t)n_stored * D;
    const long long chain0 = (long long)blockIdx.x * CH;
    R* g_theta = reinterpret_cast<R*>(a.theta);
    const R* Lm = reinterpret_cast<const R*>(a.mp.p1);
    {
        const R* g_sd = reinterpret_cast<const R*>(a.dir.sd);
        const R* g_mean = reinterpret_cast<const R*>(a.dir.mean_cols);
        for (int i = tid; i < D; i += kWsThreads) s_sd[i] = g_sd ? (float)g_sd[i] : 1.0f;
        for (int i = tid; i < n_stored * D; i += kWsThreads) s_mean[i] = (float)g_mean[i];
        for (int i = tid; i < n_cols; i += kWsThreads) s_cdf[i] = n_cols > 1 ? (float)reinterpret_cast<const R*>(a.dir.cdf)[i] : 1.0f;
        R* th_stage = reinterpret_cast<R*>(xf_all);    // theta rows, only to form w = L' theta
        for (int i = tid; i < CH * S; i += kWsThreads) {
            const int rr = i / S, cc = i - rr * S;
            th_stage[i] = (chain0 + rr < a.B && cc < D) ? g_theta[(chain0 + rr) * D + cc] : R(0);
        }
    }
    __syncthreads();
    const R tol = (R)a.fp.tol;
    const uint32_t k0s = (uint32_t)a.seed, k1s = (uint32_t)(a.seed >> 32);
    // w = L' theta, written out in the fragment layout's home positions (row 8 m + r8, columns 8 tile_q + 2 k4 + {0, 1})
    if (tensor) {
        const int warp = tid >> 5, lane = tid & 31;
        const int r8 = lane >> 2, k4 = lane & 3;
        double wf[MT][NT][2];
        dk_tri_product<MT, NT, double, kWsDepth>(reinterpret_cast<const R*>(xf_all), S, Lm, D,
                                                 ring_all + (size_t)warp * kWsDepth * NT * 64, warp, lane, wf);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int q = 0; q < NT; ++q)
                *reinterpret_cast<double2*>(w_all + (size_t)(8 * m + r8) * SW + 8 * dk_tile(q, warp) + 2 * k4) = make_double2(wf[m][q][0], wf[m][q][1]);
    }
    __syncthreads();
    for (int i = tid; i < kWsBuf * CH * S; i += kWsThreads) xf_all[i] = 0.0f;
    __syncthreads();

    if (tensor) {
        // ============================================================ tensor warps
        const int warp = tid >> 5, lane = tid & 31;
        const int o = tid >> 3, j = tid & 7;
        const int r8 = lane >> 2, k4 = lane & 3;
        const bool valid = chain0 + o < a.B;
        R* ring = ring_all + (size_t)warp * kWsDepth * NT * 64;
        for (int step = 0; step < a.n_steps; ++step) {
#ifdef KLHR_DENSE_TIMING
            const long long tq0 = clock64();
#endif
            __syncthreads();                           // X(step) drawn; c(step - 1) applied to w
#ifdef KLHR_DENSE_TIMING
            const long long tq1 = clock64();
#endif
            const int cur = step & 1;
            double vf[MT][NT][2];
            dk_tri_product<MT, NT, float, kWsDepth>(xf_all + (size_t)cur * CH * S, S, Lm, D, ring, warp, lane, vf);
#ifdef KLHR_DENSE_TIMING
            const long long tq2 = clock64();
#endif
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                double pa = 0, pb = 0;
#pragma unroll
                for (int q = 0; q < NT; ++q) {         // this thread's fragments of w: read here, read-modify-written below
                    const double2 wq = *reinterpret_cast<const double2*>(w_all + (size_t)(8 * m + r8) * SW + 8 * dk_tile(q, warp) + 2 * k4);
                    pa = fma(vf[m][q][0], vf[m][q][0], pa);
                    pb = fma(wq.x, vf[m][q][0], pb);
                    pa = fma(vf[m][q][1], vf[m][q][1], pa);
                    pb = fma(wq.y, vf[m][q][1], pb);
                }
                pa += __shfl_xor_sync(0xffffffffu, pa, 1);
                pa += __shfl_xor_sync(0xffffffffu, pa, 2);
                pb += __shfl_xor_sync(0xffffffffu, pb, 1);
                pb += __shfl_xor_sync(0xffffffffu, pb, 2);
                if (k4 == 0) {
                    red[(warp * CH + 8 * m + r8) * 2 + 0] = pa;
                    red[(warp * CH + 8 * m + r8) * 2 + 1] = pb;
                }
            }
            ws_bar_arrive(1, kWsTensor + 32);          // the sums of draw `step` are in: the fit may start
            if (valid && a.tr.rho) {                   // rho = x / ||x + tol|| (tests)
                R* g = reinterpret_cast<R*>(a.tr.rho) + ((long long)step * a.B + chain0 + o) * D;
                const float* xc = xf_all + ((size_t)cur * CH + o) * S;
                const R inv = s_var[(size_t)cur * 5 * CH + o];
                for (int i = j; i < D; i += kOct) g[i] = (R)xc[i] * inv;
            }
#ifdef KLHR_DENSE_TIMING
            const long long tq3 = clock64();
#endif
            ws_bar_sync(2, kWsTensor + 32);            // c(step) is known: w += c V
#ifdef KLHR_DENSE_TIMING
            if (blockIdx.x == 0 && lane == 0 && step == 40)
                printf("ws tensor warp %d: wait top %lld | product %lld | fold %lld | wait c %lld\n", warp, tq1 - tq0, tq2 - tq1, tq3 - tq2, clock64() - tq3);
#endif
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const R cm = s_c[(step & 1) * CH + 8 * m + r8];
#pragma unroll
                for (int q = 0; q < NT; ++q) {
                    double2* wp = reinterpret_cast<double2*>(w_all + (size_t)(8 * m + r8) * SW + 8 * dk_tile(q, warp) + 2 * k4);
                    double2 t = *wp;
                    t.x = fma(cm, vf[m][q][0], t.x);
                    t.y = fma(cm, vf[m][q][1], t.y);
                    *wp = t;
                }
            }
        }
        __syncthreads();
    } else {
        // ============================================================ producer warps
        const int ptid = tid - kWsTensor;
        constexpr int LPC = kWsLpc;
        const int pc = ptid / LPC, q4 = ptid % LPC;    // chain slot, lane of its group of LPC
        const bool pvalid = chain0 + pc < a.B;
        const unsigned gmask = (LPC == 8 ? 0xFFu : 0xFu) << (LPC * ((ptid & 31) / LPC));   // the lanes of this chain (validity is uniform over them)
        const unsigned long long cid = (unsigned long long)(a.chain_offset + chain0 + pc);
        const uint32_t c0 = (uint32_t)cid, c1 = (uint32_t)(cid >> 32);
        const bool fitter = ptid < 32;                 // the first producer warp also fits: thread ptid <-> chain ptid
        const long long c_fit = chain0 + ptid;
        long long n_acc = 0;
        unsigned long long n_evals = 0;
        double2* th2 = reinterpret_cast<double2*>(g_theta + (chain0 + pc) * D);     // D even: rows are 16-byte aligned

        // theta += c(step) x(step): theta lives in global memory (L2) -- only this move and the caller ever touch it
        auto move = [&](int step) {
            if (!pvalid) return;
            const R cm = s_c[(step & 1) * CH + pc];
            if (cm == R(0)) return;
            const float2* x2 = reinterpret_cast<const float2*>(xf_all + ((size_t)(step & 1) * CH + pc) * S);
#pragma unroll 8
            for (int i = q4; i < D / 2; i += LPC) {
                double2 t = th2[i];
                const float2 x = x2[i];
                t.x = fma(cm, (double)x.x, t.x);
                t.y = fma(cm, (double)x.y, t.y);
                th2[i] = t;
            }
        };
        auto draw_direction = [&](int step) {          // X(step) and the scalar variates of draw `step` into buffer step & 1
            if (!pvalid) return;
            const int nb = step & 1;
            float* xn = xf_all + ((size_t)nb * CH + pc) * S;
            R* var = s_var + (size_t)nb * 5 * CH;
            const unsigned long long draw = (unsigned long long)(a.draw_offset + step);
            const uint32_t d0 = (uint32_t)draw, k1d = k1s ^ (uint32_t)(draw >> 32);
            R sv0 = 0;
            if (q4 < 3) {                              // slots 0..2, same mapping and arithmetic as chain_scalars; the logarithm
                uint32_t wv[4];                        // lanes 1 and 2 both need is ONE convergent call (see dense_kernel)
                Philox::block(c0, c1, d0, (uint32_t)q4, k0s, k1d, wv);
                const R u53 = u01_53(wv[0], wv[1]);
                const R lg = r_log(u53);
                float z0, z1;
                box_muller_f32(wv[2], wv[3], z0, z1);
                if (q4 == 0) {
                    sv0 = (R)u01_32(wv[0]);
                    var[1 * CH + pc] = (R)z0;
                } else if (q4 == 1) {
                    var[2 * CH + pc] = sqrt(-2.0 * lg) * cospi(2.0 * u01_53(wv[2], wv[3]));     // = box_muller_f64
                } else {
                    var[4 * CH + pc] = u53;
                    var[3 * CH + pc] = lg;
                }
            }
            const R u_col = __shfl_sync(gmask, sv0, 0, LPC);
            int jcol = 0;
            if (n_cols > 1)                            // searchsorted(cdf, u, 'right'), klhr.py:147
                while (jcol < n_cols - 1 && (float)u_col >= s_cdf[jcol]) ++jcol;
            const float* mcol = (n_cols && jcol < n_stored) ? s_mean + (size_t)jcol * D : nullptr;
            // block b = 0 .. D / 32 - 1 of lane j8 = 0..7 holds elements 32 b + j8 + 8 r (word r) at slot kSlotDir + j8 + 8 b;
            // this thread takes j8 = q4 (8 lanes per chain) or 2 q4, 2 q4 + 1 (4 lanes), all blocks of one j8 advancing
            // round by round together
            R ss = 0;
#pragma unroll 1
            for (int h = 0; h < 8 / LPC; ++h) {
                const int j8 = (8 / LPC) * q4 + h;
                uint32_t wv[2 * NT][4];
                Philox::blockN<2 * NT>(c0, c1, d0, kSlotDir + (uint32_t)j8, 8u, k0s, k1d, wv);
#pragma unroll
                for (int b = 0; b < 2 * NT; ++b) {
                    float z[4];
                    box_muller_f32(wv[b][0], wv[b][1], z[0], z[1]);
                    box_muller_f32(wv[b][2], wv[b][3], z[2], z[3]);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int i = 32 * b + j8 + 8 * rr;
                        const float x = fmaf(s_sd[i], z[rr], mcol ? mcol[i] : 0.0f);
                        xn[i] = x;
                        const R xt = (R)x + tol;
                        ss += xt * xt;
                    }
                }
            }
#pragma unroll
            for (int off = 1; off < LPC; off <<= 1) ss += __shfl_xor_sync(gmask, ss, off);
            if (q4 == 0) var[pc] = R(1) / r_sqrt(ss);  // rho = x / ||x + tol||  (klhr.py:153)
        };

        if (a.n_steps > 0) draw_direction(0);
        for (int step = 0; step < a.n_steps; ++step) {
            __syncthreads();
#ifdef KLHR_DENSE_TIMING
            const long long pq0 = clock64();
#endif
            // under the tensor warps' product on X(step): the move of draw step - 1 (its direction sits in the buffer that
            // X(step + 1) is about to overwrite -- same lanes, program order), then X(step + 1)
            if (step > 0) move(step - 1);
            __syncwarp();                              // a lane's move reads elements that another lane of its chain redraws below
#ifdef KLHR_DENSE_TIMING
            const long long pq1 = clock64();
#endif
            if (step + 1 < a.n_steps) draw_direction(step + 1);
#ifdef KLHR_DENSE_TIMING
            const long long pq2 = clock64();
            if (blockIdx.x == 0 && (ptid & 31) == 0 && step == 40) printf("ws producer warp %d: move %lld | direction %lld\n", ptid >> 5, pq1 - pq0, pq2 - pq1);
#endif
            if (fitter) {
                ws_bar_sync(1, kWsTensor + 32);        // sums of draw `step`
                const R* var = s_var + (size_t)(step & 1) * 5 * CH;
                R cmove = 0;
                if (c_fit < a.B) {
                    double sA = 0, sB = 0;
#pragma unroll
                    for (int w8 = 0; w8 < 8; ++w8) {   // fixed order: deterministic
                        sA += red[(w8 * CH + ptid) * 2 + 0];
                        sB += red[(w8 * CH + ptid) * 2 + 1];
                    }
                    const R inv = var[ptid];
                    typename Model::Coef cf;
                    cf.A = __dmul_rn(__dmul_rn(sA, inv), inv);
                    cf.Bq = __dmul_rn(-sB, inv);
                    const R z_init = var[1 * CH + ptid], z_prop = var[2 * CH + ptid];
                    StepOut<R> so;
                    if (!quad_fit_closed(cf.A, cf.Bq, z_init, z_prop, var[3 * CH + ptid], a.fp, a.tr.eta != nullptr, so)) {
                        OrCtx<R> oc;
                        oc.K = 0; oc.inject = false; oc.r = 0; oc.v = 1;
                        fit_and_propose<1, R, Model, 2>(cf, a.fp, 0, 0u, z_init, R(0), R(0), z_prop, var[4 * CH + ptid], so, oc);
                    }
                    cmove = so.accept ? __dmul_rn(so.zp, inv) : R(0);
                    n_acc += so.accept ? 1 : 0;
                    n_evals += (unsigned long long)so.evals;
                    const long long trow = (long long)step * a.B + c_fit;
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
                        reinterpret_cast<R*>(a.tr.u)[trow] = var[4 * CH + ptid];
                    }
                }
                s_c[(step & 1) * CH + ptid] = cmove;
                ws_bar_arrive(2, kWsTensor + 32);      // c(step) is out
            }
        }
        __syncthreads();
        if (a.n_steps > 0) move(a.n_steps - 1);        // the move of the last draw
        if (fitter) {
            const bool mine = c_fit < a.B;
            if (mine && a.acc.accept_count) a.acc.accept_count[c_fit] += n_acc;
            if (a.acc.evals_total) {
                unsigned long long tot = mine ? n_evals : 0ull;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
                if (ptid == 0 && tot) atomicAdd(a.acc.evals_total, tot);
            }
        }
    }
}

__host__ inline size_t densews_smem_bytes(const StepArgs& a) {
    const int D = a.mp.D, S = D + 4, SW = D + 8, NT = D / 64;
    const int n_cols = a.dir.mean_cols ? a.dir.n_cols : 0;
    const int n_stored = n_cols ? n_cols - a.dir.n_zero_cols : 0;
    size_t b = (size_t)32 * SW * 8;                                            // w = L' theta
    b += (size_t)8 * kWsDepth * NT * 64 * 8;                                   // cp.async rings
    b += (size_t)(8 * 32 * 2 + 2 * 32 + kWsBuf * 5 * 32) * 8;                  // red, c (two draws), variates (two buffers)
    b += (size_t)kWsBuf * 32 * S * 4;                                          // X, two fp32 buffers (= the theta staging tile)
    b += (size_t)(D + (size_t)n_stored * D + ((n_cols + 3) & ~3)) * 4;         // sd, mean columns, cdf
    return b;
}

template <int NT>
int launch_densews_nt(const StepArgs& a, cudaStream_t st, LaunchInfo* info) {
    const size_t smem = densews_smem_bytes(a);
    const void* fn = (const void*)dense_ws_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (info) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, fn);
        if (e != cudaSuccess) return (int)e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kWsThreads, smem);
        if (e != cudaSuccess) return (int)e;
        info->threads = kWsThreads;
        info->smem = (int)smem;
        info->regs = fa.numRegs;
        info->ctas_per_sm = nb;
        return 0;
    }
    const long long grid = (a.B + 31) / 32;
    if (grid <= 0) return 0;
    void* kargs[] = {(void*)&a};
    e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kWsThreads), kargs, smem, st);
    return (int)e;
}

}  // namespace klhr
