// bayesn_kernels.cuh - __global__ kernels: fused log p + gradient, forward flux, device NUTS.
#pragma once
#include "bayesn_device.cuh"

namespace bayesn {

// ------------------------------------------------------------------------------------------
// (1) fused log-density + gradient for every (SN, chain)
// ------------------------------------------------------------------------------------------
// LSMEM: the evaluator's shared-memory L_Sigma path (what the sampler runs); used for the latency-mode launches so that
// the parity tests of K > 1 exercise the sampler's code, while the one-shot throughput launch walks the global tables.
template <int LP, int TP, int VPL, bool RVFREE, bool STAGE, bool LSMEM = false>
__global__ void __launch_bounds__(WPB * 32) logp_grad_kernel(ModelDev M, DataDev Dd, int n_chains, int ntile,
                                                             const float* __restrict__ q_in,
                                                             float* __restrict__ logp_out,
                                                             float* __restrict__ grad_out, int persist) {
    extern __shared__ __align__(16) float smem[];
    WarpCtx c;
    int sn, chain;
    const int lane = threadIdx.x & 31;
    if constexpr (LSMEM) if (persist) {
        // Persistent form for big batches (one resident wave of CTAs, every warp strides over the (SN, chain) units):
        // the model tables - J_l, dust basis, L_Sigma - are staged ONCE per CTA instead of once per eight evaluations,
        // and each warp stages its unit's observations / supports (/ tile) into its own slot like the sampler's work queue.
        cta_setup<LP, TP, RVFREE, STAGE, false, LSMEM>(M, Dd, n_chains, ntile, smem, c, sn, chain, true);
        const SmemPlan P = make_plan(LP, TP, M.n_eps, Dd.n_b, Dd.N_obs, Dd.wt_stride, ntile, RVFREE, STAGE, false, M.kw, LSMEM);
        const long long total = (long long)Dd.N * n_chains;
        for (long long u = (long long)blockIdx.x * WPB + (threadIdx.x >> 5); u < total; u += (long long)gridDim.x * WPB) {
            sn = (int)(u / n_chains);
            warp_bind<STAGE>(M, Dd, P, smem, c, sn, lane);
            float q[VPL], g[VPL];
#pragma unroll
            for (int s = 0; s < VPL; ++s) {
                int i = lane + 32 * s;
                q[s] = (i < M.D) ? q_in[(size_t)u * M.D + i] : 0.f;
                if (i == M.n_eps + 3) q[s] -= c.muhat;
            }
            double lp;
            eval_logp_grad<LP, TP, VPL, RVFREE, STAGE, false, LSMEM>(M, c, M.hs, q, lp, g, lane);
            if (lane == 0) logp_out[u] = (float)lp;
#pragma unroll
            for (int s = 0; s < VPL; ++s) {
                int i = lane + 32 * s;
                if (i < M.D) grad_out[(size_t)u * M.D + i] = g[s];
            }
        }
        return;
    }
    const bool active = cta_setup<LP, TP, RVFREE, STAGE, false, LSMEM>(M, Dd, n_chains, ntile, smem, c, sn, chain);
    if (!active) return;
    const size_t unit = (size_t)sn * n_chains + chain;
    float q[VPL], g[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        int i = lane + 32 * s;
        q[s] = (i < M.D) ? q_in[unit * M.D + i] : 0.f;
        if (i == M.n_eps + 3) q[s] -= c.muhat;     // the evaluator works with Ds - muhat
    }
    double lp;
    eval_logp_grad<LP, TP, VPL, RVFREE, STAGE, false, LSMEM>(M, c, M.hs, q, lp, g, lane);
    if (((threadIdx.x >> 5) & (M.kw - 1)) != 0) return;       // latency mode: the team's warps hold identical results
    if (lane == 0) logp_out[unit] = (float)lp;
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        int i = lane + 32 * s;
        if (i < M.D) grad_out[unit * M.D + i] = g[s];
    }
}

// ------------------------------------------------------------------------------------------
// (2) forward only: get_flux_batch / get_mag_batch with the reference's argument layouts
//     (bayesn/bayesn_model.py:645-759).  One warp per (observation, SN); lanes over bins.
// ------------------------------------------------------------------------------------------
struct FluxArgs {
    int N, N_obs, n_b, L, T, mag;
    float M0;
    const float *theta, *AV, *W0, *W1, *eps, *Ds, *RV;
    const int* band;
    const uint8_t* mask;
    const float *J_t, *hsi, *weights;
    const double* bscale;   // [n_b] log2 of the band scale (double: the exponent is ~+60)
    const float *Jl, *hs, *Mf;   // model tables: [JROWS][RS], [NPH][BP], [9][BP]
    int LPs;                // row stride RS of Jl
    float* out;
};

// SPECTRA = false: the band-pass integral of every (observation, SN) -> out (N_obs, N), flux or magnitude.
// SPECTRA = true : SEDmodel.get_spectra (bayesn/bayesn_model.py:556-643), the host-dust-extinguished
//                  rest-frame SED on the 300-bin model grid, out [N][300][N_obs].
template <bool SPECTRA>
__global__ void __launch_bounds__(128) flux_kernel(FluxArgs A) {
    __shared__ float su[4][BAYESN_MAX_L];
    __shared__ float syk[4][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long item = (long long)blockIdx.x * 4 + warp;
    const long long total = (long long)A.N * A.N_obs;
    if (item >= total) return;
    const int s = (int)(item / A.N_obs), o = (int)(item - (long long)s * A.N_obs);
    const size_t on = (size_t)o * A.N + s;
    const int L = A.L, T = A.T;
    // u[l] = -c2 * sum_m (W0 + theta W1 + eps)[l][m] J_t[s][m][o]
    if (lane < L) {
        float th = A.theta[s], acc = 0.f;
        for (int m = 0; m < T; ++m) {
            float w = A.W0[lane * T + m] + th * A.W1[lane * T + m] + A.eps[((size_t)s * L + lane) * T + m];
            acc = fmaf(w, A.J_t[((size_t)s * T + m) * A.N_obs + o], acc);
        }
        su[warp][lane] = -C2 * acc;
    }
    const float RV = A.RV[s];
    if (lane == 0) {
        float yk[9], dyk[9];
        f99_yk(RV, yk, dyk);
#pragma unroll
        for (int j = 0; j < 9; ++j) syk[warp][j] = yk[j];
    }
    __syncwarp();
    const float AVc = -C2 * A.AV[s];
    const float iR = 1.f / RV;
    int i0 = (int)A.hsi[on], i1 = (int)A.hsi[(size_t)A.N_obs * A.N + on];
    const float f = A.hsi[2 * (size_t)A.N_obs * A.N + on];
    i0 = min(max(i0, 0), NPH - 1);
    i1 = min(max(i1, 0), NPH - 1);
    const int k = SPECTRA ? 0 : A.band[on];
    const float* wcol = SPECTRA ? nullptr : A.weights + (size_t)s * NB * A.n_b + k;
    float F = 0.f;
    for (int b = lane; b < NB; b += 32) {
        float g = 0.f;
        for (int l = 0; l < L; ++l) g = fmaf(A.Jl[b * A.LPs + l], su[warp][l], g);
        float my = 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) my = fmaf(A.Mf[j * BP + b], syk[warp][j], my);
        const float a = 1.f + my * iR;
        const float h0 = A.hs[i0 * BP + b], h1 = A.hs[i1 * BP + b];
        const float H = fmaf(f, h1 - h0, h0);
        const float sed = H * ex2_approx(fmaf(AVc, a, g));
        if (SPECTRA) A.out[((size_t)s * NB + b) * A.N_obs + o] = sed;
        else F = fmaf(wcol[(size_t)b * A.n_b], sed, F);
    }
    if (SPECTRA) return;
    F = wsum(F);
    if (lane == 0) {
        const bool msk = A.mask[on] != 0;
        // scale = 10^(-0.4 (M0 + Ds)) * band scale, assembled in the log2 domain
        float flux = F * (float)exp2(A.bscale[k] - 1.3287712379549449 * ((double)A.M0 + (double)A.Ds[s]));
        flux = msk ? flux : 0.f;
        float outv = flux;
        if (A.mag) {
            float fm = flux + (msk ? 0.f : 0.01f);
            outv = msk ? (-2.5f * log10f(fm) + 27.5f) : 0.f;
        }
        A.out[on] = outv;
    }
}

// ------------------------------------------------------------------------------------------
// (2c) SEDmodel.get_flux_from_chains (bayesn/bayesn_model.py:3426-3524): model photometry of posterior
//      draws on a phase grid, ONE launch over (SN x draw x phase).  The reference loops over SNe and
//      runs simulate_light_curve per SN (4.91 s per SN in example_fits.ipynb); here a warp takes one
//      (SN, draw, phase): it builds the extinguished SED on the 300-bin grid once (lanes over bins, staged
//      in shared memory) and integrates it through every band the SN was observed in, reading the SN's
//      compact weight tile (the same tiles + supports the fit kernels read, built by bayesn_tiles_*).
//      out [N][S][max_bands][n_t], flux (FLUXCAL) or magnitudes.
// ------------------------------------------------------------------------------------------
struct ChainFluxArgs {
    int N, S, n_t, max_bands, n_b, L, T, n_eps, LP, TP, RS, mag, rvfree;
    const float* t;          // [n_t] rest-frame phases
    const float* theta;      // [N][S]
    const float* AV;         // [N][S]
    const float* tmax;       // [N][S]
    const float* dD;         // [N][S]  (mu + delM) - fold_s: the distance relative to the exponent folded into the tile
    const float* RV;         // [N][S] when rvfree
    const float* eps;        // [N][S][n_eps], eps = L_Sigma eps_tform, element (l-1) + (L-2) m
    const int* bands;        // [N][max_bands] column of the SN's tile (index into its support rows), -1 = none
    const float* wt;         // [N][wt_stride]
    const int4* sup;         // [N][n_b]
    int wt_stride;
    const float *Jl, *hs, *Mf, *ad, *W0, *W1, *KD, *tau;      // model tables (layouts of ModelDev)
    float* out;
};

template <int TP>
__global__ void __launch_bounds__(128) chain_flux_kernel(ChainFluxArgs A) {
    __shared__ float su[4][BAYESN_MAX_L];
    __shared__ float syk[4][12];
    __shared__ float ssed[4][BP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long item = (long long)blockIdx.x * 4 + warp;
    const long long total = (long long)A.N * A.S * A.n_t;
    if (item >= total) return;
    const int it = (int)(item % A.n_t);
    const long long ns = item / A.n_t;            // (SN, draw)
    const int sn = (int)(ns / A.S);
    const int L = A.L, T = A.T;
    const float t = A.t[it] - A.tmax[ns];
    const float theta = A.theta[ns];
    // u[l] = -c2 sum_m (W0 + theta W1 + eps)[l][m] J_t(t)[m]; every lane l < L builds the spline row itself
    if (lane < L) {
        float J[TP], dJ[TP];
        spline_row<TP>(t, T, A.tau, A.KD, J, dJ);
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < TP; ++m) {
            if (m < T) {
                float w = fmaf(theta, A.W1[lane * A.TP + m], A.W0[lane * A.TP + m]);
                if (lane >= 1 && lane <= L - 2) w += A.eps[(size_t)ns * A.n_eps + (lane - 1) + (L - 2) * m];
                acc = fmaf(w, J[m], acc);
            }
        }
        su[warp][lane] = -C2 * acc;
    }
    float iR = 0.f;
    if (A.rvfree) {
        const float RV = A.RV[ns];
        iR = 1.f / RV;
        if (lane == 0) {
            float yk[9], dyk[9];
            f99_yk(RV, yk, dyk);
#pragma unroll
            for (int j = 0; j < 9; ++j) syk[warp][j] = yk[j];
        }
    }
    __syncwarp();
    const float AVc = -C2 * A.AV[ns];
    const float fl = floorf(t);
    int i0 = 19 + (int)fl, i1 = 19 + (int)ceilf(t);
    i0 = min(max(i0, 0), NPH - 1);
    i1 = min(max(i1, 0), NPH - 1);
    const float f = t - fl;
    for (int b = lane; b < BP; b += 32) {
        float sed = 0.f;
        if (b < NB) {
            float g = 0.f;
            for (int l = 0; l < L; ++l) g = fmaf(A.Jl[b * A.RS + l], su[warp][l], g);
            float a;
            if (A.rvfree) {
                float my = 0.f;
#pragma unroll
                for (int j = 0; j < 9; ++j) my = fmaf(A.Mf[j * BP + b], syk[warp][j], my);
                a = 1.f + my * iR;
            } else {
                a = A.ad[b];
            }
            const float h0 = A.hs[i0 * BP + b], h1 = A.hs[i1 * BP + b];
            sed = fmaf(f, h1 - h0, h0) * ex2_approx(fmaf(AVc, a, g));
        }
        ssed[warp][b] = sed;
    }
    __syncwarp();
    const float S = exp2f(-C2 * A.dD[ns]);
    const float* tile = A.wt + (size_t)sn * A.wt_stride;
    for (int j = 0; j < A.max_bands; ++j) {
        const int k = A.bands[(size_t)sn * A.max_bands + j];
        float F = 0.f;
        if (k >= 0) {
            const int4 sp = A.sup[(size_t)sn * A.n_b + k];
            for (int b = sp.x + lane; b < sp.y; b += 32) F = fmaf(tile[sp.z + b - sp.x], ssed[warp][b], F);
            F = wsum(F);
        }
        if (lane == 0) {
            float outv = 0.f;
            if (k >= 0) {
                const float flux = S * F;
                outv = A.mag ? (-2.5f * log10f(flux) + 27.5f) : flux;
            }
            A.out[(((size_t)ns) * A.max_bands + j) * A.n_t + it] = outv;
        }
    }
}

// ------------------------------------------------------------------------------------------
// (3) device-resident NUTS: one warp per chain, iterative tree doubling, on-device adaptation.
//     Semantics follow numpyro.infer.NUTS (third-party, not in the reference tree; restated from
//     SURVEY.md Appendix C): multinomial sampling, checkpointed U-turn test, dual averaging,
//     Stan-style windowed diagonal mass matrix.
// ------------------------------------------------------------------------------------------
#ifndef NUTS_MIN_CTAS
#define NUTS_MIN_CTAS (16 / BAYESN_WPB)
#endif
enum { V_ZL = 0, V_RL, V_GL, V_ZR, V_RR, V_GR, V_ZP, V_GP, V_RSUM, V_SZP, V_SGP, V_WMEAN, V_WM2, V_CKPT };
__host__ __device__ inline int nuts_num_vectors(int max_depth) { return V_CKPT + 2 * max_depth; }

template <int VPL>
struct VecIO {
    float* base;   // this chain's workspace: [nvec][32*VPL]
    int lane;
    __device__ __forceinline__ void st(int v, const float (&x)[VPL]) const {
#pragma unroll
        for (int s = 0; s < VPL; ++s) base[(v * VPL + s) * 32 + lane] = x[s];
    }
    __device__ __forceinline__ void ld(int v, float (&x)[VPL]) const {
#pragma unroll
        for (int s = 0; s < VPL; ++s) x[s] = base[(v * VPL + s) * 32 + lane];
    }
};

template <int VPL>
__device__ __forceinline__ float vdot_im(const float (&im)[VPL], const float (&a)[VPL], const float (&b)[VPL]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s = fmaf(im[i] * a[i], b[i], s);
    return wsum(s);
}

// generalised U-turn test (numpyro _is_turning): r_sum - (r_left + r_right)/2 against both ends
template <int VPL>
__device__ __forceinline__ bool is_turning(const float (&im)[VPL], const float (&rl)[VPL], const float (&rr)[VPL],
                                           const float (&rsum)[VPL]) {
    float sl = 0.f, sr = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        float rs = rsum[i] - 0.5f * (rl[i] + rr[i]);
        sl = fmaf(im[i] * rl[i], rs, sl);
        sr = fmaf(im[i] * rr[i], rs, sr);
    }
    sl = wsum(sl);
    sr = wsum(sr);
    return (sl <= 0.f) || (sr <= 0.f);
}

__device__ __forceinline__ float logaddexpf(float a, float b) {
    float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + log1pf(expf(-fabsf(a - b)));
}

template <int LP, int TP, int VPL, bool RVFREE, bool STAGE>
__global__ void __launch_bounds__(WPB * 32, NUTS_MIN_CTAS) nuts_kernel(ModelDev M, DataDev Dd, NutsDev cfg, int ntile,
                                                        const float* __restrict__ q_init, float* __restrict__ samples,
                                                        float* __restrict__ stats, float* __restrict__ chain_info,
                                                        float* __restrict__ work, unsigned long long* queue) {
    extern __shared__ __align__(16) float smem[];
    WarpCtx c;
    int sn, chain;
    const int C = cfg.num_chains;
    // queue != nullptr: work-queue mode.  The grid is one resident wave; every warp takes the next
    // (SN, chain) unit from an atomic counter when its chain is done and stages that SN's data itself,
    // so no warp idles while the slowest chain of its CTA finishes and the last wave fills evenly
    // (C3: +8 % at 1000 SNe, 309 SNe/s at 2368 SNe).  Draws do not depend on the schedule: random
    // streams and workspace are keyed by the unit.  A variant that kept the chains of an SN on one CTA
    // (CTA-level current SN + global SN counter) was 2 % slower: locality of the SN's Hsiao rows in L1
    // does not matter, balance does.  queue == nullptr: CTA b runs units [b WPB, (b+1) WPB), their SN
    // data staged once per CTA (used when a per-warp slot of SN data would cost occupancy).
    const bool dyn = queue != nullptr;
    const bool active = cta_setup<LP, TP, RVFREE, STAGE, false, true>(M, Dd, C, ntile, smem, c, sn, chain, dyn);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    // Latency mode (M.kw > 1): the kw warps of a team run this whole state machine redundantly on the same unit -
    // identical inputs, identical decisions - and only split the observation pairs inside eval_logp_grad; each
    // keeps its own copy of the tree vectors, the team's first warp writes the results.
    const int tk = (threadIdx.x >> 5) & (M.kw - 1);
    const bool writer = tk == 0;
    const unsigned long long total_units = (unsigned long long)Dd.N * C;
    for (bool first_unit = true;; first_unit = false) {
        size_t unit;
        if (dyn) {
            unsigned long long u = 0;
            if (tk == 0) {
                if (lane == 0) u = atomicAdd(queue, 1ULL);
                u = __shfl_sync(FULL, u, 0);
            }
            if (M.kw > 1) {
                // the team's first warp fetched the unit; it hands it to the others through its scalar slot
                // (rewritten only a whole chain - many team barriers - later)
                if (tk == 0 && lane == 0) *reinterpret_cast<unsigned long long*>(c.sc + 2) = u;
                team_sync((threadIdx.x >> 5) >> M.kwl, M.kw);
                u = *reinterpret_cast<const unsigned long long*>(c.sc - tk * M.xstride + 2);
            }
            if (u >= total_units) break;
            sn = (int)(u / (unsigned)C);
            chain = (int)(u - (unsigned long long)sn * C);
            const SmemPlan P = make_plan(LP, TP, M.n_eps, Dd.n_b, Dd.N_obs, Dd.wt_stride, ntile, RVFREE, STAGE, false, M.kw, true);
            warp_bind<STAGE>(M, Dd, P, smem, c, sn, lane);
        } else {
            if (!first_unit) break;
        }
        unit = (size_t)sn * C + chain;
        const int D = M.D;
        const int nvec = nuts_num_vectors(cfg.max_depth);
        VecIO<VPL> io{work + ((unit << M.kwl) + tk) * (size_t)nvec * 32 * VPL, lane};
        const uint2 key = make_uint2(cfg.seed_lo, cfg.seed_hi);
        const uint32_t gunit = (uint32_t)unit + cfg.unit_offset;      // random streams do not depend on the sharding
        bool own[VPL];
#pragma unroll
        for (int s = 0; s < VPL; ++s) own[s] = (lane + 32 * s) < D;

        float z[VPL], r[VPL], g[VPL], im[VPL], rsum[VPL], tmp[VPL], tmp2[VPL];
#pragma unroll
        for (int s = 0; s < VPL; ++s) {
            z[s] = own[s] ? q_init[unit * D + lane + 32 * s] : 0.f;
            if (lane + 32 * s == M.n_eps + 3) z[s] -= c.muhat;   // sample Ds - muhat (float32 resolution)
            im[s] = own[s] ? 1.f : 0.f;
            tmp[s] = 0.f;
        }
        io.st(V_WMEAN, tmp);
        io.st(V_WM2, tmp);
        // adaptation state (numpyro warmup_adapter / dual_averaging / welford_covariance)
        double pe = 0.0;   // potential energy = -log p (double: see eval_logp_grad)
        float step = cfg.init_step;
        float da_x = 0.f, da_xavg = 0.f, da_gavg = 0.f, da_prox = logf(10.f * step);
        int da_t = 0, wf_n = 0, window = 0;
        float mean_acc = 0.f;
        long long total_steps = 0;
        int n_div = 0;
        const int n_iter = cfg.num_warmup + cfg.num_samples;

        // per-transition / per-doubling state
        int it = 0;
        uint32_t rk = 0;
        double E0 = 0.0, prop_pe = 0.0, sprop_pe = 0.0;
        float w_main = 0.f, sum_acc = 0.f;
        int depth = 0, n_prop = 0;
        bool turning = false, diverging = false;
        bool going_right = false;
        float u_main = 0.f, eps_dir = 0.f;
        int n_max = 1, n_sub = 0;
        bool s_turn = false, s_div = false;
        float w_sub = -INFINITY, s_acc = 0.f;

        auto begin_transition = [&]() {
            rk = 0;
            // momentum refresh r ~ N(0, M)
            uint4 ra = philox4x32(make_uint4((uint32_t)it, 0x40000000u, (uint32_t)lane, gunit), key);
            float n0 = sqrtf(-2.f * logf(u01(ra.x) + 5.9604644775390625e-8f));
            float n1 = sqrtf(-2.f * logf(u01(ra.z) + 5.9604644775390625e-8f));
            float sn0, cs0, sn1, cs1;
            sincospif(2.f * u01(ra.y), &sn0, &cs0);
            sincospif(2.f * u01(ra.w), &sn1, &cs1);
            float nrm[4] = {n0 * cs0, n0 * sn0, n1 * cs1, n1 * sn1};
#pragma unroll
            for (int s = 0; s < VPL; ++s) r[s] = own[s] ? nrm[s & 3] * rsqrtf(im[s]) : 0.f;
            E0 = pe + 0.5 * (double)vdot_im<VPL>(im, r, r);
            io.st(V_ZL, z); io.st(V_RL, r); io.st(V_GL, g);
            io.st(V_ZR, z); io.st(V_RR, r); io.st(V_GR, g);
            io.st(V_ZP, z); io.st(V_GP, g); io.st(V_RSUM, r);
            prop_pe = pe;
            w_main = 0.f; sum_acc = 0.f;
            depth = 0; n_prop = 0;
            turning = false; diverging = false;
        };
        auto begin_doubling = [&]() {
            uint4 rd = philox4x32(make_uint4((uint32_t)it, rk++, 0xffffffffu, gunit), key);
            going_right = (rd.x >> 31) != 0;
            u_main = u01(rd.y);
            if (going_right) { io.ld(V_ZR, z); io.ld(V_RR, r); io.ld(V_GR, g); }
            else { io.ld(V_ZL, z); io.ld(V_RL, r); io.ld(V_GL, g); }
            eps_dir = going_right ? step : -step;
            n_max = 1 << depth;
            n_sub = 0;
            s_turn = false; s_div = false;
            w_sub = -INFINITY; s_acc = 0.f; sprop_pe = 0.0;
        };

        // Flat state machine: every pass through this loop does exactly one log p + gradient
        // evaluation (the very first one initialises the chain, all others are leapfrog steps).
        // mode 0: leapfrog steps.  mode 1: evaluate the initial point.  modes 2, 3: distance rescue (below).
        int mode = 1;
#ifdef BAYESN_PHASE_TIMING
        long long t_prev_eval = 0;
#endif
        int init_tries = 0, n_rescue = 0;
        float rs_S1 = 0.f, rs_h1 = 0.f;
        const int iD = M.n_eps + 3;             // element of the (shifted) distance
        while (true) {
            const bool first = mode != 0;
            if (!first) {
#pragma unroll
                for (int s = 0; s < VPL; ++s) {
                    r[s] = fmaf(-0.5f * eps_dir, g[s], r[s]);
                    z[s] = fmaf(eps_dir * im[s], r[s], z[s]);
                }
            }
            double lp;
#ifdef BAYESN_PHASE_TIMING
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                const long long t_now = clock64();
                if (t_prev_eval) atomicAdd(&g_phase[6], (unsigned long long)(t_now - t_prev_eval));
                atomicAdd(&g_phase[7], 1ULL);
            }
#endif
            eval_logp_grad<LP, TP, VPL, RVFREE, STAGE, false, true>(M, c, M.hs, z, lp, g, lane);
#ifdef BAYESN_PHASE_TIMING
            t_prev_eval = clock64();
#endif
            const double pe_new = -lp;
#pragma unroll
            for (int s = 0; s < VPL; ++s) g[s] = -g[s];
            if (first) {
                // numpyro rejects initial points with a non-finite potential or gradient and retries
                // with init_to_uniform(radius=2) (infer/util.py find_valid_initial_params); same here.
                bool ok = isfinite(pe_new);
#pragma unroll
                for (int s = 0; s < VPL; ++s) ok = ok && isfinite(g[s]);
                ok = __all_sync(FULL, ok);
                if (!ok && init_tries < 100) {
                    uint4 ra = philox4x32(make_uint4((uint32_t)init_tries, 0x20000000u, (uint32_t)lane, gunit), key);
                    float uu[4] = {u01(ra.x), u01(ra.y), u01(ra.z), u01(ra.w)};
#pragma unroll
                    for (int s = 0; s < VPL; ++s) z[s] = own[s] ? (4.f * uu[s & 3] - 2.f) : 0.f;
                    ++init_tries;
                    continue;
                }
                if (mode == 2) {
                    // Second point of the rescue: log L is quadratic in the flux scale S = 10^(-0.4 dD), so
                    // h = (dU/d dD) / (kappa S) = a - b S at two distances gives the optimum S* = a / b.
                    const float dD2 = vget<VPL>(z, iD), S2 = exp2f(-C2 * dD2);
                    const float h2 = vget<VPL>(g, iD) / (KAPPA * S2);
                    const float b = (rs_h1 - h2) / (S2 - rs_S1), a = rs_h1 + b * rs_S1;
                    float dDs = -log2f(a / b) * (1.f / C2);
                    if (!(b > 0.f) || !(a > 0.f) || !isfinite(dDs)) dDs = 0.f;       // fall back to Ds = muhat
#pragma unroll
                    for (int s = 0; s < VPL; ++s)
                        if (lane + 32 * s == iD) z[s] = dDs;
                    mode = 3;
                    continue;
                }
                if (mode == 3) {
                    // restart step-size adaptation from the rescued point (the mass matrix is kept)
                    step = cfg.init_step;
                    da_x = 0.f; da_xavg = 0.f; da_gavg = 0.f; da_t = 0;
                    da_prox = logf(10.f * step);
                }
                mode = 0;
                pe = pe_new;
                begin_transition();
                begin_doubling();
                continue;
            }
            // ---- finish the leapfrog, account for the new leaf --------------------------------------
#pragma unroll
            for (int s = 0; s < VPL; ++s) r[s] = fmaf(-0.5f * eps_dir, g[s], r[s]);
            {
                const double E_new = pe_new + 0.5 * (double)vdot_im<VPL>(im, r, r);
                float dE = (float)(E_new - E0);
                if (!(dE == dE)) dE = INFINITY;
                const float leaf_w = -dE;
                s_div = dE > 1000.f;
                s_acc += fminf(1.f, expf(-dE));
                if (n_sub == 0) {
#pragma unroll
                    for (int s = 0; s < VPL; ++s) rsum[s] = r[s];
                    w_sub = leaf_w;
                    io.st(V_SZP, z); io.st(V_SGP, g);
                    sprop_pe = pe_new;
                } else {
#pragma unroll
                    for (int s = 0; s < VPL; ++s) rsum[s] += r[s];
                    uint4 rl = philox4x32(make_uint4((uint32_t)it, rk++, 0xffffffffu, gunit), key);
                    const float tp = 1.f / (1.f + expf(-(leaf_w - w_sub)));   // uniform transition kernel
                    if (u01(rl.x) < tp) {
                        io.st(V_SZP, z); io.st(V_SGP, g);
                        sprop_pe = pe_new;
                    }
                    w_sub = logaddexpf(w_sub, leaf_w);
                }
                const int leaf = n_sub;
                ++n_sub;
                // checkpoint indices (numpyro _leaf_idx_to_ckpt_idxs)
                const int idx_max = __popc(leaf >> 1);
                const int n_sub_trees = __ffs(~leaf) - 1;   // trailing ones
                const int idx_min = idx_max - n_sub_trees + 1;
                if ((leaf & 1) == 0) {
                    io.st(V_CKPT + 2 * idx_max, r);
                    io.st(V_CKPT + 2 * idx_max + 1, rsum);
                }
                for (int i = idx_max; i >= idx_min && !s_turn; --i) {
                    io.ld(V_CKPT + 2 * i, tmp);          // r_ckpt
                    io.ld(V_CKPT + 2 * i + 1, tmp2);     // r_sum_ckpt
#pragma unroll
                    for (int s = 0; s < VPL; ++s) tmp2[s] = rsum[s] - tmp2[s] + tmp[s];
                    s_turn = is_turning<VPL>(im, tmp, r, tmp2);
                }
            }
            if (n_sub < n_max && !s_turn && !s_div) continue;

            // ---- subtree finished: merge into the trajectory (biased progressive sampling) ----------
            total_steps += n_sub;
            if (going_right) { io.st(V_ZR, z); io.st(V_RR, r); io.st(V_GR, g); }
            else { io.st(V_ZL, z); io.st(V_RL, r); io.st(V_GL, g); }
            io.ld(V_RSUM, tmp);
#pragma unroll
            for (int s = 0; s < VPL; ++s) tmp[s] += rsum[s];
            io.st(V_RSUM, tmp);
            {
                const float tp = (s_turn || s_div) ? 0.f : fminf(1.f, expf(w_sub - w_main));
                if (u_main < tp) {
                    io.ld(V_SZP, tmp2); io.st(V_ZP, tmp2);
                    io.ld(V_SGP, tmp2); io.st(V_GP, tmp2);
                    prop_pe = sprop_pe;
                }
            }
            if (s_turn) turning = true;
            else {
                io.ld(V_RL, tmp2);
                float rr_[VPL];
                io.ld(V_RR, rr_);
                turning = is_turning<VPL>(im, tmp2, rr_, tmp);
            }
            w_main = logaddexpf(w_main, w_sub);
            diverging = s_div;
            sum_acc += s_acc;
            n_prop += n_sub;
            ++depth;
            if (depth < cfg.max_depth && !turning && !diverging) {
                begin_doubling();
                continue;
            }

            // ---- transition finished ---------------------------------------------------------------
            const float accept = sum_acc / (float)max(n_prop, 1);
            io.ld(V_ZP, z); io.ld(V_GP, g);
            pe = prop_pe;
            if (lane == 0 && writer) {
                float* srow = stats + (unit * (size_t)n_iter + it) * 5;
                srow[0] = accept;
                srow[1] = (float)n_prop;
                srow[2] = diverging ? 1.f : 0.f;
                srow[3] = (float)pe;
                srow[4] = step;     // step size this transition was run with
            }
            if (it < cfg.num_warmup) {
                // warm-up adaptation (numpyro warmup_adapter.update_fn)
                if (cfg.adapt_step) {
                    ++da_t;
                    const float gg = cfg.target_accept - accept;
                    da_gavg = (1.f - 1.f / (da_t + 10.f)) * da_gavg + gg / (da_t + 10.f);
                    da_x = da_prox - sqrtf((float)da_t) / 0.05f * da_gavg;
                    const float wt = powf((float)da_t, -0.75f);
                    da_xavg = (1.f - wt) * da_xavg + wt * da_x;
                    step = (it == cfg.num_warmup - 1) ? expf(da_xavg) : expf(da_x);
                    step = fmaxf(step, 1.17549435e-38f);
                }
                const bool middle = (window > 0) && (window < cfg.n_windows - 1);
                if (cfg.adapt_mass && middle) {
                    ++wf_n;
                    io.ld(V_WMEAN, tmp); io.ld(V_WM2, tmp2);
#pragma unroll
                    for (int s = 0; s < VPL; ++s) {
                        float dpre = z[s] - tmp[s];
                        tmp[s] += dpre / (float)wf_n;
                        tmp2[s] = fmaf(dpre, z[s] - tmp[s], tmp2[s]);
                    }
                    io.st(V_WMEAN, tmp); io.st(V_WM2, tmp2);
                }
                const bool at_end = (it == cfg.win_end[window]);
                if (at_end) ++window;
                if (at_end && middle) {
                    if (cfg.adapt_mass) {
                        io.ld(V_WM2, tmp2);
#pragma unroll
                        for (int s = 0; s < VPL; ++s) {
                            float cov = tmp2[s] / (float)(wf_n - 1);
                            if (cfg.regularize) cov = (wf_n / (wf_n + 5.f)) * cov + 1e-3f * (5.f / (wf_n + 5.f));
                            im[s] = own[s] ? cov : 0.f;
                            tmp[s] = 0.f;
                        }
                        io.st(V_WMEAN, tmp); io.st(V_WM2, tmp);
                        wf_n = 0;
                    }
                    if (cfg.adapt_step) {
                        da_prox = logf(10.f * step);
                        da_x = 0.f; da_xavg = 0.f; da_gavg = 0.f; da_t = 0;
                    }
                }
            } else {
                const int k = it - cfg.num_warmup;
                const size_t row = unit * cfg.num_samples + k;
#pragma unroll
                for (int s = 0; s < VPL; ++s)
                    if (own[s] && writer) samples[row * D + lane + 32 * s] = z[s] + ((lane + 32 * s == M.n_eps + 3) ? c.muhat : 0.f);
                mean_acc += (accept - mean_acc) / (float)(k + 1);
                n_div += diverging ? 1 : 0;
            }
            if (++it >= n_iter) break;
            // FP32 safeguard (not in numpyro, which runs in float64): a chain whose initial distance is ~2 mag
            // off starts at log p ~ -2e6, where the rounding noise of a float32 log L (relative 1e-7..1e-6) is
            // an energy noise of order 1 per leapfrog; no step size reaches the target acceptance, dual
            // averaging drives the step to ~1e-8, the position stops changing and every transition runs the
            // full 2^max_depth - 1 leapfrogs (1 chain in 9 472 of the C5 share: 498 k leapfrogs, a third of the
            // kernel time).  Healthy chains never go below ~1e-3 during warm-up.  Rescue: solve the distance
            // coordinate in closed form (two evaluations), restart the step-size adaptation, carry on.
            if (it < cfg.num_warmup && cfg.adapt_step && step < 1e-6f && n_rescue < 3) {
                ++n_rescue;
                // recorded with the transition that ended here: stats[2] = diverging + 2 (warm-up rows only)
                if (lane == 0 && writer) stats[(unit * (size_t)n_iter + (it - 1)) * 5 + 2] += 2.f;
                const float dD1 = vget<VPL>(z, iD);
                rs_S1 = exp2f(-C2 * dD1);
                rs_h1 = vget<VPL>(g, iD) / (KAPPA * rs_S1);
#pragma unroll
                for (int s = 0; s < VPL; ++s)
                    if (lane + 32 * s == iD) z[s] = dD1 + 0.5f;
                mode = 2;
                continue;
            }
            begin_transition();
            begin_doubling();
        }
        float* ci = chain_info + unit * (4 + D);
        if (!writer) continue;
        if (lane == 0) {
            ci[0] = step;
            ci[1] = mean_acc;
            ci[2] = (float)total_steps;
            ci[3] = (float)n_div;
        }
#pragma unroll
        for (int s = 0; s < VPL; ++s)
            if (own[s]) ci[4 + lane + 32 * s] = im[s];
    }
}


// ------------------------------------------------------------------------------------------
// (4) numpyro init_to_median(num_samples=15): per chain and site, the median of 15 prior draws,
//     mapped to the unconstrained space (reference call sites :1757, :2110).
// ------------------------------------------------------------------------------------------
struct InitArgs {
    int N, C, D, n_eps, rv_mode;
    float tauA, sdD;
    const float* muhat;
    unsigned seed_lo, seed_hi, unit_offset;
    float* q;
};

__device__ __forceinline__ float median15(float (&v)[15]) {
    // selection of the 8th smallest by partial insertion sort
#pragma unroll
    for (int i = 1; i < 15; ++i) {
        float x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
    return v[7];
}

__global__ void init_median_kernel(InitArgs A) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)A.N * A.C * A.D;
    if (idx >= total) return;
    const int i = (int)(idx % A.D);
    const long long unit = idx / A.D;
    const int sn = (int)(unit / A.C);
    const uint2 key = make_uint2(A.seed_lo ^ 0x5bd1e995u, A.seed_hi + 0x1234567u);
    float v[15];
    const int kind = (i < A.n_eps) ? 0 : (i - A.n_eps + 1);   // 0 eps, 1 theta, 2 AV, 3 tmax, 4 Ds, 5 RV
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint4 ra = philox4x32(make_uint4((uint32_t)b, (uint32_t)i, 0x80000000u, (uint32_t)unit + A.unit_offset), key);
        float u[4] = {u01(ra.x), u01(ra.y), u01(ra.z), u01(ra.w)};
        float n0 = sqrtf(-2.f * logf(u[0] + 5.9604644775390625e-8f));
        float n1 = sqrtf(-2.f * logf(u[2] + 5.9604644775390625e-8f));
        float sn0, cs0, sn1, cs1;
        sincospif(2.f * u[1], &sn0, &cs0);
        sincospif(2.f * u[3], &sn1, &cs1);
        float nr[4] = {n0 * cs0, n0 * sn0, n1 * cs1, n1 * sn1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int j = 4 * b + k;
            if (j < 15) v[j] = (kind == 2 || kind == 3 || kind == 5) ? u[k] : nr[k];
        }
    }
    float med = median15(v);
    float out;
    if (kind == 0 || kind == 1) out = med;
    else if (kind == 2) out = logf(-A.tauA * logf(fmaxf(1.f - med, 1e-7f)));       // Exponential draw, then log
    else if (kind == 4) out = A.muhat[sn] + A.sdD * med;
    else { float u = fminf(fmaxf(med, 1e-6f), 1.f - 1e-6f); out = logf(u) - log1pf(-u); }   // uniform -> logit
    A.q[idx] = out;
}

// Test aid: leave every SM's shared memory full of NaN bit patterns, as an unrelated earlier kernel might.  The
// evaluator's lanes that run past a band's support read such words and multiply them by an exact zero; cta_setup clears
// the CTA's allocation so that they can never be NaN (tests/test_gpu_round2.py::test_stale_shared_memory_is_harmless).
__global__ void poison_smem_kernel(int nfloats) {
    extern __shared__ __align__(16) float psm[];
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) psm[i] = __int_as_float(0x7fc00000 | (i & 0xffff));
    __syncthreads();
    if (psm[(threadIdx.x * 7919) % nfloats] == 0.f) psm[0] = 1.f;      // keep the stores
}

// ------------------------------------------------------------------------------------------
// (5) FP32 FMA peak microbenchmark: the roofline denominator for the SIMT kernels above
//     (MEASURED_PEAKS.json has no FP32 figure).  8 independent FFMA chains per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
          x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace bayesn
