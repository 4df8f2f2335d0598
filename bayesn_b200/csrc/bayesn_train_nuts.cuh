// bayesn_train_nuts.cuh - NUTS for the population-training posterior (train_model_globalRV,
// reference bayesn/bayesn_model.py:1115-1182 sampled by numpyro NUTS at :1767-1770, :1913-1919),
// device resident, with the state SHARDED: every (SN, chain) warp owns that SN's latent slice of
// the position / momentum / tree vectors, one CTA per chain owns the ~1k global coordinates and
// every scalar decision of the sampler.  One "pass" = one leapfrog step of every chain:
//
//   train_nuts_sn_kernel    per (SN, chain) warp: apply the vector actions decided at the end of the
//                           previous pass to the latent slice, leapfrog it (its gradient is local),
//                           evaluate, and emit partial sums: global-gradient contributions, log p,
//                           kinetic energy and every U-turn dot product this leaf can need
//   train_reduce_kernel     sum over the SNe of this rank                        (bayesn_train.cuh)
//   [all-reduce of the [C][R] double buffer when SNe are sharded over GPUs: the one exchange per leapfrog]
//   train_nuts_ctl_kernel   per chain CTA: chain rule to the unconstrained globals, finish their
//                           leapfrog, combine the dot products, take the tree / adaptation decisions
//                           (same semantics as nuts_kernel in bayesn_kernels.cuh), apply them to the
//                           global slice, start the next leapfrog of the globals and publish the
//                           constrained tables + action flags for the next pass
//
// Ranks that hold different SN shards run the ctl kernel redundantly on bit-identical inputs, so
// they take identical decisions without any further communication.
#pragma once
#include "bayesn_kernels.cuh"
#include "bayesn_train.cuh"

namespace bayesn {

#ifdef CTL_TIMING
__device__ long long g_ctl_t[16];
#define CTL_T(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_ctl_t[k] = clock64(); } while (0)
#else
#define CTL_T(k)
#endif

enum { TV_ZL = 0, TV_RL, TV_GL, TV_ZR, TV_RR, TV_GR, TV_ZP, TV_GP, TV_RSUM, TV_SZP, TV_SGP, TV_WMEAN, TV_WM2,
       TV_CZ, TV_CR, TV_CG, TV_RSUB, TV_IM, TV_CKPT };
__host__ __device__ inline int train_nuts_num_vectors(int max_depth) { return TV_CKPT + 2 * max_depth; }

enum { ACT_SUB_ACCEPT = 1, ACT_MERGE = 2, ACT_MAIN_ACCEPT = 4, ACT_TRANS_DONE = 8, ACT_WELFORD = 16, ACT_MASS = 32,
       ACT_SAMPLE = 64, ACT_NEW_TRANS = 128, ACT_NEW_DOUBLING = 256, ACT_INIT = 512, ACT_DONE = 1024 };

// what the ctl kernel publishes for the SN kernel (and keeps for itself), one per chain
struct TrainCtl {
    int act, it, leaf, n_max, going_right, merge_right, sample_row, wf_n, wf_act;
    float eps_dir;
    // sampler scalars (ctl kernel only)
    int phase;                 // 0 = waiting for the initial evaluation, 1 = running, 2 = finished
    unsigned rk;
    int depth, n_sub, n_prop, e0_pending, da_t, window, n_div;
    int turning, diverging;
    float step, u_main, w_main, w_sub, sum_acc, s_acc, da_x, da_xavg, da_gavg, da_prox, mean_acc;
    double pe, E0, prop_pe, sprop_pe, ke0_g;
    long long total_steps;
    unsigned xseq;             // passes exchanged so far (fused reduction + peer exchange in the ctl kernel)
};

struct TrainNutsDev {
    NutsDev cfg;
    TrainCtl* ctl;        // [C]
    float* work_l;        // [N][C][nvec][32*VPL] latent slices of the tree vectors
    float* work_g;        // [C][nvec][G]         global slices
    float* gtmp;          // [C][G]
    float* samples_g;     // [C][S][G]
    float* samples_l;     // [N][C][S][Dl]
    float* stats;         // [C][W+S][5] accept, n_leapfrog, diverging, potential energy, step size
    int sn_offset;        // global index of this rank's first SN (keys the latent momentum streams)
    int VPL;
};

// ------------------------------------------------------------------------------------------
// per (SN, chain) warp
// ------------------------------------------------------------------------------------------
template <int LP, int TP, int VPL, bool STAGE>
__global__ void __launch_bounds__(WPB * 32, NUTS_MIN_CTAS) train_nuts_sn_kernel(ModelDev M, DataDev Dd, TrainDev Tr,
                                                                                TrainNutsDev Nt, int ntile) {
    extern __shared__ __align__(16) float smem[];
    WarpCtx c;
    int sn, chain;
    const int C = Tr.C;
    const bool active = cta_setup<LP, TP, false, STAGE, true>(M, Dd, C, ntile, smem, c, sn, chain);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    const size_t unit = (size_t)sn * C + chain;
    const int ne = M.n_eps, Dl = Tr.Dl;
    const TrainCtl ctl = Nt.ctl[chain];
    const int act = ctl.act;
    const int nvec = train_nuts_num_vectors(Nt.cfg.max_depth);
    VecIO<VPL> io{Nt.work_l + unit * (size_t)nvec * 32 * VPL, lane};
    float* part = Tr.part + unit * Tr.PR;
    float* px = part + train_part_base(Tr.LPTP);
    bool own[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) own[s] = (lane + 32 * s) < Dl;
    float z[VPL], r[VPL], g[VPL], im[VPL], t1[VPL], t2[VPL];
    // Latency mode (M.kw > 1, used when the rank's SNe leave warp slots idle - a sharded training run): the
    // team's first warp owns the latent slice (all vector bookkeeping below), hands the new position to its
    // teammates through shared memory, and all kw warps share the observation pairs of the ONE evaluation of
    // this pass (eval_logp_grad exchanges the partial sums); the teammates then leave.
    const int tk = (threadIdx.x >> 5) & (M.kw - 1);
    constexpr int XW = xchg_words(LP, TP);
    const float eps = ctl.eps_dir;
    float ke0 = 0.f;
    if (tk != 0) {
        if (act & ACT_DONE) return;
        team_sync((threadIdx.x >> 5) >> M.kwl, M.kw);
        const float* zs = c.sc - tk * M.xstride + M.xoff + XW;       // the leader's spare exchange buffer
#pragma unroll
        for (int s = 0; s < VPL; ++s) z[s] = own[s] ? zs[lane + 32 * s] : 0.f;
    } else {
    io.ld(TV_IM, im);

    // ---- vector actions decided at the end of the previous pass ---------------------------------
    if (act & ACT_SUB_ACCEPT) { io.ld(TV_CZ, t1); io.st(TV_SZP, t1); io.ld(TV_CG, t1); io.st(TV_SGP, t1); }
    if (act & ACT_MERGE) {
        io.ld(TV_CZ, t1); io.st(ctl.merge_right ? TV_ZR : TV_ZL, t1);
        io.ld(TV_CR, t1); io.st(ctl.merge_right ? TV_RR : TV_RL, t1);
        io.ld(TV_CG, t1); io.st(ctl.merge_right ? TV_GR : TV_GL, t1);
        io.ld(TV_RSUM, t1); io.ld(TV_RSUB, t2);
#pragma unroll
        for (int s = 0; s < VPL; ++s) t1[s] += t2[s];
        io.st(TV_RSUM, t1);
    }
    if (act & ACT_MAIN_ACCEPT) { io.ld(TV_SZP, t1); io.st(TV_ZP, t1); io.ld(TV_SGP, t1); io.st(TV_GP, t1); }
    if (act & ACT_TRANS_DONE) {
        io.ld(TV_ZP, z); io.st(TV_CZ, z);
        io.ld(TV_GP, t1); io.st(TV_CG, t1);
        if (act & ACT_WELFORD) {
            io.ld(TV_WMEAN, t1); io.ld(TV_WM2, t2);
#pragma unroll
            for (int s = 0; s < VPL; ++s) {
                const float dpre = z[s] - t1[s];
                t1[s] += dpre / (float)ctl.wf_act;
                t2[s] = fmaf(dpre, z[s] - t1[s], t2[s]);
            }
            io.st(TV_WMEAN, t1); io.st(TV_WM2, t2);
        }
        if (act & ACT_MASS) {
            io.ld(TV_WM2, t2);
#pragma unroll
            for (int s = 0; s < VPL; ++s) {
                float cov = t2[s] / (float)(ctl.wf_act - 1);
                if (Nt.cfg.regularize) cov = (ctl.wf_act / (ctl.wf_act + 5.f)) * cov + 1e-3f * (5.f / (ctl.wf_act + 5.f));
                im[s] = own[s] ? cov : 0.f;
                t1[s] = 0.f;
            }
            io.st(TV_IM, im); io.st(TV_WMEAN, t1); io.st(TV_WM2, t1);
        }
        if (act & ACT_SAMPLE) {
            const size_t row = unit * Nt.cfg.num_samples + ctl.sample_row;
#pragma unroll
            for (int s = 0; s < VPL; ++s)
                if (own[s]) Nt.samples_l[row * Dl + lane + 32 * s] = z[s] + ((lane + 32 * s == ne + 2) ? c.muhat : 0.f);
        }
    }
    if (act & ACT_DONE) return;
    if (act & ACT_NEW_TRANS) {
        // momentum refresh r ~ N(0, M) for this SN's latents; the stream is keyed by the global SN index
        const uint2 key = make_uint2(Nt.cfg.seed_lo, Nt.cfg.seed_hi);
        const uint32_t gunit = (uint32_t)((Nt.sn_offset + sn) * C + chain);
        uint4 ra = philox4x32(make_uint4((uint32_t)ctl.it, 0x40000000u, (uint32_t)lane, gunit), key);
        float n0 = sqrtf(-2.f * logf(u01(ra.x) + 5.9604644775390625e-8f));
        float n1 = sqrtf(-2.f * logf(u01(ra.z) + 5.9604644775390625e-8f));
        float sn0, cs0, sn1, cs1;
        sincospif(2.f * u01(ra.y), &sn0, &cs0);
        sincospif(2.f * u01(ra.w), &sn1, &cs1);
        const float nrm[4] = {n0 * cs0, n0 * sn0, n1 * cs1, n1 * sn1};
#pragma unroll
        for (int s = 0; s < VPL; ++s) r[s] = own[s] ? nrm[s & 3] * rsqrtf(im[s]) : 0.f;
        ke0 = 0.5f * vdot_im<VPL>(im, r, r);
        io.st(TV_CR, r); io.st(TV_RL, r); io.st(TV_RR, r); io.st(TV_RSUM, r);
        io.ld(TV_CZ, t1); io.st(TV_ZL, t1); io.st(TV_ZR, t1); io.st(TV_ZP, t1);
        io.ld(TV_CG, t1); io.st(TV_GL, t1); io.st(TV_GR, t1); io.st(TV_GP, t1);
    }
    if (act & ACT_NEW_DOUBLING) {
        io.ld(ctl.going_right ? TV_ZR : TV_ZL, t1); io.st(TV_CZ, t1);
        io.ld(ctl.going_right ? TV_RR : TV_RL, t1); io.st(TV_CR, t1);
        io.ld(ctl.going_right ? TV_GR : TV_GL, t1); io.st(TV_CG, t1);
    }
    __syncwarp();

    // ---- one leapfrog step of the latent slice (its gradient only needs this SN) ----------------
    io.ld(TV_CZ, z); io.ld(TV_CR, r); io.ld(TV_CG, g);
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        r[s] = fmaf(-0.5f * eps, g[s], r[s]);
        z[s] = fmaf(eps * im[s], r[s], z[s]);
    }
    if (M.kw > 1) {
        float* zs = c.sc + M.xoff + XW;
#pragma unroll
        for (int s = 0; s < VPL; ++s)
            if (own[s]) zs[lane + 32 * s] = z[s];
        team_sync((threadIdx.x >> 5) >> M.kwl, M.kw);
    }
    }   // tk == 0
    c.sW0 = Tr.W0 + (size_t)chain * Tr.LPTP;
    c.sW1 = Tr.W1 + (size_t)chain * Tr.LPTP;
    c.Ls = Tr.Ls + (size_t)chain * ne * ne;
    c.LsT = Tr.LsT + (size_t)chain * ne * ne;
    c.ad_g = Tr.ad + (size_t)chain * JROWS;
    c.dad_g = Tr.dad + (size_t)chain * JROWS;
    const float* sc = Tr.sc + (size_t)chain * 8;
    c.sigma0 = sc[0];
    c.tauA = sc[2];
    c.muerr2 = Dd.muerr2[sn];
    c.part = part;
    double lp;
    eval_logp_grad<LP, TP, VPL, false, STAGE, true>(M, c, M.hs, z, lp, g, lane);
    if (tk != 0) return;
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        g[s] = -g[s];
        r[s] = fmaf(-0.5f * eps, g[s], r[s]);
    }
    io.st(TV_CZ, z); io.st(TV_CR, r); io.st(TV_CG, g);
    const float ke = 0.5f * vdot_im<VPL>(im, r, r);
    if (lane == 0) { px[0] = ke; px[1] = ke0; }
    if (act & ACT_INIT) return;

    // ---- tree bookkeeping for leaf ctl.leaf (numpyro iterative NUTS: checkpointed U-turn test) -----
    const int leaf = ctl.leaf;
    float rsub[VPL];
    if (leaf == 0) {
#pragma unroll
        for (int s = 0; s < VPL; ++s) rsub[s] = r[s];
    } else {
        io.ld(TV_RSUB, rsub);
#pragma unroll
        for (int s = 0; s < VPL; ++s) rsub[s] += r[s];
    }
    io.st(TV_RSUB, rsub);
    const int idx_max = __popc(leaf >> 1);
    const int n_sub_trees = __ffs(~leaf) - 1;
    const int idx_min = idx_max - n_sub_trees + 1;
    if ((leaf & 1) == 0) { io.st(TV_CKPT + 2 * idx_max, r); io.st(TV_CKPT + 2 * idx_max + 1, rsub); }
    for (int i = idx_max; i >= idx_min; --i) {
        io.ld(TV_CKPT + 2 * i, t1);          // r at the left end of this sub-tree
        io.ld(TV_CKPT + 2 * i + 1, t2);      // r_sum up to and including it
        float sl = 0.f, sr = 0.f;
#pragma unroll
        for (int s = 0; s < VPL; ++s) {
            const float rs = (rsub[s] - t2[s] + t1[s]) - 0.5f * (t1[s] + r[s]);
            sl = fmaf(im[s] * t1[s], rs, sl);
            sr = fmaf(im[s] * r[s], rs, sr);
        }
        sl = wsum(sl); sr = wsum(sr);
        if (lane == 0) { px[2 + 2 * i] = sl; px[3 + 2 * i] = sr; }
    }
    if (leaf + 1 == ctl.n_max) {
        // the sub-tree completes with this leaf: U-turn test of the whole trajectory after the merge
        io.ld(TV_RSUM, t1);
        io.ld(ctl.going_right ? TV_RL : TV_RR, t2);
        float sl = 0.f, sr = 0.f;
#pragma unroll
        for (int s = 0; s < VPL; ++s) {
            const float rl = ctl.going_right ? t2[s] : r[s];
            const float rr = ctl.going_right ? r[s] : t2[s];
            const float rs = (t1[s] + rsub[s]) - 0.5f * (rl + rr);
            sl = fmaf(im[s] * rl, rs, sl);
            sr = fmaf(im[s] * rr, rs, sr);
        }
        sl = wsum(sl); sr = wsum(sr);
        if (lane == 0) { px[TRAIN_NX - 2] = sl; px[TRAIN_NX - 1] = sr; }
    }
}

// ------------------------------------------------------------------------------------------
// per chain CTA (256 threads): globals + every scalar decision
// ------------------------------------------------------------------------------------------
struct GVec {
    float* base;   // [nvec][G]
    int G;
    __device__ __forceinline__ float* v(int k) const { return base + (size_t)k * G; }
};

__device__ __forceinline__ void gcopy(const GVec& W, int dst, int src) {
    for (int i = threadIdx.x; i < W.G; i += blockDim.x) W.v(dst)[i] = W.v(src)[i];
}

// block-wide sums of NV per-thread partials (NV <= 32); every thread gets all totals in out[]
// only the values whose bit is set in `need` are reduced (the others come back as 0)
template <int NV>
__device__ __forceinline__ void block_sums(float (&v)[NV], float* sh /* [8][NV] */, float (&out)[NV],
                                           unsigned need = 0xffffffffu) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if ((need >> k) & 1u) {
            float x = v[k];
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
            if (l == 0) sh[w * NV + k] = x;
        }
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float t = 0.f;
        if ((need >> k) & 1u)
            for (int i = 0; i < nw; ++i) t += sh[i * NV + k];
        out[k] = t;
    }
    __syncthreads();
}

// Fused second reduction stage + exchange (fused != 0): the chain's CTA adds the chunk partials of train_reduce_kernel
// itself (fixed order) and - when SNe are sharded over ranks with the peer-memory exchange - stores its sums into its
// slot of EVERY rank's symmetric buffer over NVLink, publishes the pass number with st.release.sys, waits for the
// other ranks' numbers (ld.acquire.sys) and adds the slots in rank order: the same bits on every rank.  One kernel
// less per pass and no NCCL launch; slots and flags are per (parity, rank, chain), so chains never wait for each other.
// A rank runs at most one pass ahead of its peers (it needs their sums to finish a pass), which the two parities cover.
__device__ __forceinline__ bool ctl_reduce_exchange(const TrainDev& Tr, const XchgDev& X, int c, int nchunk, unsigned seq,
                                                    double* __restrict__ red_s) {
    const int tid = threadIdx.x, R = Tr.R, C = Tr.C;
    const unsigned par = seq & 1u;
    for (int r = tid; r < R; r += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < nchunk; ++k) acc += Tr.rpart[((size_t)k * C + c) * R + r];
        if (X.world > 1) {
            const size_t slot = (((size_t)par * X.world + X.rank) * C + c) * R + r;
            for (int p = 0; p < X.world; ++p) X.peer_buf[p][slot] = acc;
        } else {
            red_s[r] = acc;
        }
    }
    if (X.world <= 1) {
        __syncthreads();
        return true;
    }
    __shared__ int xbad;
    if (tid == 0) xbad = (*reinterpret_cast<volatile unsigned*>(X.err) != 0u);
    __threadfence_system();
    __syncthreads();
    if (tid < X.world) {
        unsigned* f = X.peer_flag[tid] + ((size_t)par * X.world + X.rank) * C + c;          // my flag in rank tid's memory
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        const unsigned* mine = X.my_flag + ((size_t)par * X.world + tid) * C + c;            // rank tid's flag in mine
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (clock64() - t0 > 4000000000LL) { *X.err = 1u; xbad = 1; break; }      // ~2 s: never hang the GPU
        } while (v != seq && !xbad);
    }
    __syncthreads();
    if (xbad) return false;
    const double* src = X.my_buf + ((size_t)par * X.world * C + c) * R;
    for (int r = tid; r < R; r += blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < X.world; ++p) acc += __ldcg(src + (size_t)p * C * R + r);   // L2: the peers wrote it
        red_s[r] = acc;
    }
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(256) train_nuts_ctl_kernel(ModelDev M, TrainDev Tr, TrainNutsDev Nt, int LP, int TP,
                                                             double* __restrict__ red_all, XchgDev X, int fused, int nchunk) {
    __shared__ TrainShared S;
    __shared__ float shsum[8 * TRAIN_NX];
    __shared__ TrainCtl sctl;
    // dynamic shared memory: [R doubles: this chain's reduced sums | n_eps^2 floats: scratch of the
    // stick-breaking row scans | G floats: the unconstrained global position].  Staging the sums and the
    // position once turns the ~25 dependent global-memory phases of this kernel into shared-memory ones.
    extern __shared__ __align__(16) unsigned char ctl_smem[];
    double* red_s = reinterpret_cast<double*>(ctl_smem);
    float* tmat = reinterpret_cast<float*>(red_s + Tr.R);
    float* qg_s = tmat + M.n_eps * M.n_eps;
    const int c = blockIdx.x, tid = threadIdx.x;
    const int G = Tr.G;
    const NutsDev& cfg = Nt.cfg;
    CTL_T(0);
    if (tid == 0) sctl = Nt.ctl[c];
    __syncthreads();
    if (sctl.phase == 2) {
        if (tid == 0) Nt.ctl[c].act = ACT_DONE;
        return;
    }
    const int nvec = train_nuts_num_vectors(cfg.max_depth);
    GVec W{Nt.work_g + (size_t)c * nvec * G, G};
    bool got;
    if (fused) {
        const unsigned seq = sctl.xseq + 1u;
        got = ctl_reduce_exchange(Tr, X, c, nchunk, seq, red_s);
        if (tid == 0) sctl.xseq = seq;
    } else {
        // sums of the other ranks' SN shards (peer-memory exchange after the separate second stage; no-op on one
        // GPU / with NCCL)
        got = xchg_gather(X, Tr, c, red_all);
        if (got)
            for (int r = tid; r < Tr.R; r += blockDim.x) red_s[r] = red_all[(size_t)c * Tr.R + r];
    }
    if (!got) {
        // a rank's sums never arrived: stop this chain here (the host sees the error flag at its next status poll)
        if (tid == 0) { sctl.phase = 2; sctl.act = ACT_DONE; Nt.ctl[c] = sctl; }
        return;
    }
    for (int i = tid; i < G; i += blockDim.x) qg_s[i] = W.v(TV_CZ)[i];
    __syncthreads();
    const double* red = red_s;
    const double* rx = red + train_red_base(Tr.LPTP, M.n_eps);
    float* gnew = Nt.gtmp + (size_t)c * G;
    const uint2 key = make_uint2(cfg.seed_lo, cfg.seed_hi);
    const uint32_t gkey = 0x7fff0000u + (uint32_t)c;          // stream id of the global slice of chain c

    // ---- 1. finish the evaluation made at CZ: gradient of the potential wrt the unconstrained globals
    CTL_T(1);
    train_finish_dev(M, Tr, LP, TP, c, qg_s, red, gnew, -1.f, tmat);
    __syncthreads();
    CTL_T(2);
    const double pe_new = -(red[0] + Tr.lpg[c]);
    const float eps = sctl.eps_dir;
    const int leaf = sctl.leaf;
    const int idx_max = __popc(leaf >> 1);
    const int idx_min = idx_max - (__ffs(~leaf) - 1) + 1;
    const bool running = sctl.phase == 1;
    // ---- 2. second half of the leapfrog of the global slice + its share of every dot product --------
    float pv[TRAIN_NX], tot[TRAIN_NX];
#pragma unroll
    for (int k = 0; k < TRAIN_NX; ++k) pv[k] = 0.f;
    {
        float *cr = W.v(TV_CR), *cg = W.v(TV_CG), *imv = W.v(TV_IM), *rsub = W.v(TV_RSUB);
        for (int i = tid; i < G; i += blockDim.x) {
            const float gi = gnew[i];
            const float ri = fmaf(-0.5f * eps, gi, cr[i]);
            cr[i] = ri;
            cg[i] = gi;
            const float imi = imv[i];
            pv[0] = fmaf(0.5f * imi * ri, ri, pv[0]);
            if (running) {
                const float rs_i = (leaf == 0) ? ri : rsub[i] + ri;
                rsub[i] = rs_i;
                if ((leaf & 1) == 0) { W.v(TV_CKPT + 2 * idx_max)[i] = ri; W.v(TV_CKPT + 2 * idx_max + 1)[i] = rs_i; }
                for (int k = idx_max; k >= idx_min; --k) {
                    const float rl = W.v(TV_CKPT + 2 * k)[i], rsk = W.v(TV_CKPT + 2 * k + 1)[i];
                    const float rs = (rs_i - rsk + rl) - 0.5f * (rl + ri);
                    pv[2 + 2 * k] = fmaf(imi * rl, rs, pv[2 + 2 * k]);
                    pv[3 + 2 * k] = fmaf(imi * ri, rs, pv[3 + 2 * k]);
                }
                if (leaf + 1 == sctl.n_max) {
                    const float other = W.v(sctl.going_right ? TV_RL : TV_RR)[i];
                    const float rl = sctl.going_right ? other : ri, rr = sctl.going_right ? ri : other;
                    const float rs = (W.v(TV_RSUM)[i] + rs_i) - 0.5f * (rl + rr);
                    pv[TRAIN_NX - 2] = fmaf(imi * rl, rs, pv[TRAIN_NX - 2]);
                    pv[TRAIN_NX - 1] = fmaf(imi * rr, rs, pv[TRAIN_NX - 1]);
                }
            }
        }
    }
    CTL_T(3);
    {
        // kinetic energy always; the checkpoint pairs idx_min..idx_max and the whole-tree pair only when this leaf uses them
        unsigned need = 1u;
        if (running) {
            for (int k = idx_min; k <= idx_max; ++k) need |= 3u << (2 + 2 * k);
            if (leaf + 1 == sctl.n_max) need |= 3u << (TRAIN_NX - 2);
        }
        block_sums<TRAIN_NX>(pv, shsum, tot, need);
    }
    CTL_T(4);

    // ---- 3. scalar decisions (thread 0), mirrored from nuts_kernel ------------------------------------
    if (tid == 0) {
        TrainCtl& s = sctl;
        int act = 0;
        const int n_iter = cfg.num_warmup + cfg.num_samples;
        auto begin_doubling = [&]() {
            uint4 rd = philox4x32(make_uint4((uint32_t)s.it, s.rk++, 0xffffffffu, gkey), key);
            s.going_right = (rd.x >> 31) != 0;
            s.u_main = u01(rd.y);
            s.n_max = 1 << s.depth;
            s.n_sub = 0;
            s.w_sub = -INFINITY; s.s_acc = 0.f; s.sprop_pe = 0.0;
            act |= ACT_NEW_DOUBLING;
        };
        auto begin_transition = [&]() {
            s.rk = 0;
            s.e0_pending = 1;
            s.prop_pe = s.pe;
            s.w_main = 0.f; s.sum_acc = 0.f;
            s.depth = 0; s.n_prop = 0;
            s.turning = 0; s.diverging = 0;
            act |= ACT_NEW_TRANS;
        };
        if (s.phase == 0) {
            s.pe = pe_new;
            s.phase = 1;
            begin_transition();
            begin_doubling();
        } else {
            if (s.e0_pending) {
                s.E0 = s.pe + (double)(float)rx[1] + s.ke0_g;
                s.e0_pending = 0;
            }
            const double E_new = pe_new + (double)((float)rx[0] + tot[0]);
            float dE = (float)(E_new - s.E0);
            if (!(dE == dE)) dE = INFINITY;
            const float leaf_w = -dE;
            const bool s_div = dE > 1000.f;
            s.s_acc += fminf(1.f, expf(-dE));
            if (s.n_sub == 0) {
                s.w_sub = leaf_w;
                act |= ACT_SUB_ACCEPT;
                s.sprop_pe = pe_new;
            } else {
                uint4 rl = philox4x32(make_uint4((uint32_t)s.it, s.rk++, 0xffffffffu, gkey), key);
                const float tp = 1.f / (1.f + expf(-(leaf_w - s.w_sub)));
                if (u01(rl.x) < tp) { act |= ACT_SUB_ACCEPT; s.sprop_pe = pe_new; }
                s.w_sub = logaddexpf(s.w_sub, leaf_w);
            }
            ++s.n_sub;
            bool s_turn = false;
            for (int k = idx_max; k >= idx_min; --k) {
                const float sl = (float)rx[2 + 2 * k] + tot[2 + 2 * k], sr = (float)rx[3 + 2 * k] + tot[3 + 2 * k];
                if (sl <= 0.f || sr <= 0.f) s_turn = true;
            }
            if (s.n_sub < s.n_max && !s_turn && !s_div) {
                s.leaf = s.n_sub;                     // next leaf of the same sub-tree
            } else {
                // sub-tree finished: merge (biased progressive sampling)
                s.total_steps += s.n_sub;
                act |= ACT_MERGE;
                s.merge_right = s.going_right;
                const float tp = (s_turn || s_div) ? 0.f : fminf(1.f, expf(s.w_sub - s.w_main));
                if (s.u_main < tp) { act |= ACT_MAIN_ACCEPT; s.prop_pe = s.sprop_pe; }
                if (s_turn) s.turning = 1;
                else {
                    const float sl = (float)rx[TRAIN_NX - 2] + tot[TRAIN_NX - 2];
                    const float sr = (float)rx[TRAIN_NX - 1] + tot[TRAIN_NX - 1];
                    s.turning = (s.n_sub == s.n_max) ? ((sl <= 0.f) || (sr <= 0.f)) : 0;
                }
                s.w_main = logaddexpf(s.w_main, s.w_sub);
                s.diverging = s_div;
                s.sum_acc += s.s_acc;
                s.n_prop += s.n_sub;
                ++s.depth;
                if (s.depth < cfg.max_depth && !s.turning && !s.diverging) {
                    begin_doubling();
                    s.leaf = 0;
                } else {
                    // transition finished
                    const float accept = s.sum_acc / (float)max(s.n_prop, 1);
                    act |= ACT_TRANS_DONE;
                    s.pe = s.prop_pe;
                    float* srow = Nt.stats + ((size_t)c * n_iter + s.it) * 5;
                    srow[0] = accept; srow[1] = (float)s.n_prop; srow[2] = s.diverging ? 1.f : 0.f;
                    srow[3] = (float)s.pe; srow[4] = s.step;
                    if (s.it < cfg.num_warmup) {
                        if (cfg.adapt_step) {
                            ++s.da_t;
                            const float gg = cfg.target_accept - accept;
                            s.da_gavg = (1.f - 1.f / (s.da_t + 10.f)) * s.da_gavg + gg / (s.da_t + 10.f);
                            s.da_x = s.da_prox - sqrtf((float)s.da_t) / 0.05f * s.da_gavg;
                            const float wt = powf((float)s.da_t, -0.75f);
                            s.da_xavg = (1.f - wt) * s.da_xavg + wt * s.da_x;
                            s.step = (s.it == cfg.num_warmup - 1) ? expf(s.da_xavg) : expf(s.da_x);
                            s.step = fmaxf(s.step, 1.17549435e-38f);
                        }
                        const bool middle = (s.window > 0) && (s.window < cfg.n_windows - 1);
                        if (cfg.adapt_mass && middle) { ++s.wf_n; act |= ACT_WELFORD; }
                        const bool at_end = (s.it == cfg.win_end[s.window]);
                        if (at_end) ++s.window;
                        if (at_end && middle) {
                            if (cfg.adapt_mass) act |= ACT_MASS;
                            if (cfg.adapt_step) {
                                s.da_prox = logf(10.f * s.step);
                                s.da_x = 0.f; s.da_xavg = 0.f; s.da_gavg = 0.f; s.da_t = 0;
                            }
                        }
                    } else {
                        const int k = s.it - cfg.num_warmup;
                        act |= ACT_SAMPLE;
                        s.sample_row = k;
                        s.mean_acc += (accept - s.mean_acc) / (float)(k + 1);
                        s.n_div += s.diverging ? 1 : 0;
                    }
                    ++s.it;
                    if (s.it >= n_iter) {
                        s.phase = 2;
                        act |= ACT_DONE;
                    } else {
                        begin_transition();
                        begin_doubling();
                        s.leaf = 0;
                    }
                }
            }
        }
        s.act = act;
        s.wf_act = s.wf_n;
        if (act & ACT_MASS) s.wf_n = 0;
        s.eps_dir = s.going_right ? s.step : -s.step;
    }
    __syncthreads();

    CTL_T(5);
    // ---- 4. apply the vector actions to the global slice (same order as the SN kernel) --------------------
    const int act = sctl.act;
    const int wf_n = sctl.wf_act;
    if (act & ACT_SUB_ACCEPT) { gcopy(W, TV_SZP, TV_CZ); gcopy(W, TV_SGP, TV_CG); }
    if (act & ACT_MERGE) {
        const bool mr = sctl.merge_right;
        gcopy(W, mr ? TV_ZR : TV_ZL, TV_CZ); gcopy(W, mr ? TV_RR : TV_RL, TV_CR); gcopy(W, mr ? TV_GR : TV_GL, TV_CG);
        for (int i = tid; i < G; i += blockDim.x) W.v(TV_RSUM)[i] += W.v(TV_RSUB)[i];
    }
    if (act & ACT_MAIN_ACCEPT) { gcopy(W, TV_ZP, TV_SZP); gcopy(W, TV_GP, TV_SGP); }
    if (act & ACT_TRANS_DONE) {
        gcopy(W, TV_CZ, TV_ZP); gcopy(W, TV_CG, TV_GP);
        for (int i = tid; i < G; i += blockDim.x) {
            const float zi = W.v(TV_ZP)[i];
            if (act & ACT_WELFORD) {
                float mean = W.v(TV_WMEAN)[i], m2 = W.v(TV_WM2)[i];
                const float dpre = zi - mean;
                mean += dpre / (float)wf_n;
                m2 = fmaf(dpre, zi - mean, m2);
                W.v(TV_WMEAN)[i] = mean; W.v(TV_WM2)[i] = m2;
            }
            if (act & ACT_MASS) {
                float cov = W.v(TV_WM2)[i] / (float)(wf_n - 1);
                if (cfg.regularize) cov = (wf_n / (wf_n + 5.f)) * cov + 1e-3f * (5.f / (wf_n + 5.f));
                W.v(TV_IM)[i] = cov;
                W.v(TV_WMEAN)[i] = 0.f; W.v(TV_WM2)[i] = 0.f;
            }
            if (act & ACT_SAMPLE) Nt.samples_g[((size_t)c * cfg.num_samples + sctl.sample_row) * G + i] = zi;
        }
    }
    __syncthreads();
    if (act & ACT_DONE) {
        if (tid == 0) Nt.ctl[c] = sctl;
        return;
    }
    if (act & ACT_NEW_TRANS) {
        float kp[1] = {0.f}, kt[1];
        for (int i = tid; i < G; i += blockDim.x) {
            const int b = i >> 2;
            uint4 ra = philox4x32(make_uint4((uint32_t)sctl.it, 0x40000000u, (uint32_t)b, gkey), key);
            const float n0 = sqrtf(-2.f * logf(u01(ra.x) + 5.9604644775390625e-8f));
            const float n1 = sqrtf(-2.f * logf(u01(ra.z) + 5.9604644775390625e-8f));
            float sn0, cs0, sn1, cs1;
            sincospif(2.f * u01(ra.y), &sn0, &cs0);
            sincospif(2.f * u01(ra.w), &sn1, &cs1);
            const float nrm[4] = {n0 * cs0, n0 * sn0, n1 * cs1, n1 * sn1};
            const float imi = W.v(TV_IM)[i];
            const float ri = nrm[i & 3] * rsqrtf(imi);
            kp[0] = fmaf(0.5f * imi * ri, ri, kp[0]);
            W.v(TV_CR)[i] = ri; W.v(TV_RL)[i] = ri; W.v(TV_RR)[i] = ri; W.v(TV_RSUM)[i] = ri;
            const float zi = W.v(TV_CZ)[i], gi = W.v(TV_CG)[i];
            W.v(TV_ZL)[i] = zi; W.v(TV_ZR)[i] = zi; W.v(TV_ZP)[i] = zi;
            W.v(TV_GL)[i] = gi; W.v(TV_GR)[i] = gi; W.v(TV_GP)[i] = gi;
        }
        block_sums<1>(kp, shsum, kt);
        if (tid == 0) sctl.ke0_g = (double)kt[0];
    }
    __syncthreads();
    if (act & ACT_NEW_DOUBLING) {
        const bool gr = sctl.going_right;
        gcopy(W, TV_CZ, gr ? TV_ZR : TV_ZL); gcopy(W, TV_CR, gr ? TV_RR : TV_RL); gcopy(W, TV_CG, gr ? TV_GR : TV_GL);
    }
    __syncthreads();
    CTL_T(6);
    // ---- 5. first half of the next leapfrog of the global slice, then the constrained tables ------------
    {
        const float e2 = sctl.eps_dir;
        float *cz = W.v(TV_CZ), *cr = W.v(TV_CR), *cg = W.v(TV_CG), *imv = W.v(TV_IM);
        for (int i = tid; i < G; i += blockDim.x) {
            const float ri = fmaf(-0.5f * e2, cg[i], cr[i]);
            cr[i] = ri;
            const float zi = fmaf(e2 * imv[i], ri, cz[i]);
            cz[i] = zi;
            qg_s[i] = zi;
        }
    }
    __syncthreads();
    CTL_T(7);
    train_globals_dev(M, Tr, LP, TP, c, qg_s, S, tmat);
    CTL_T(8);
    if (tid == 0) Nt.ctl[c] = sctl;
}

// initial state: positions from the caller, unit inverse mass, ctl = "evaluate the initial point"
__global__ void train_nuts_init_kernel(ModelDev M, TrainDev Tr, TrainNutsDev Nt, DataDev Dd, int LP, int TP,
                                       const float* __restrict__ qg0, const float* __restrict__ ql0) {
    __shared__ TrainShared S;
    extern __shared__ __align__(16) float tmat[];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int G = Tr.G, Dl = Tr.Dl, C = Tr.C, VPL = Nt.VPL;
    const int nvec = train_nuts_num_vectors(Nt.cfg.max_depth);
    GVec W{Nt.work_g + (size_t)c * nvec * G, G};
    for (int k = 0; k < nvec; ++k)
        for (int i = tid; i < G; i += blockDim.x)
            W.v(k)[i] = (k == TV_CZ) ? qg0[(size_t)c * G + i] : (k == TV_IM ? 1.f : 0.f);
    for (int sn = 0; sn < Dd.N; ++sn) {
        float* wl = Nt.work_l + ((size_t)sn * C + c) * nvec * 32 * VPL;
        for (int e = tid; e < nvec * 32 * VPL; e += blockDim.x) {
            const int k = e / (32 * VPL), i = e - k * 32 * VPL;     // element i lives at [slot i/32][lane i%32]
            float v = 0.f;
            if (i < Dl) {
                if (k == TV_CZ) v = ql0[((size_t)sn * C + c) * Dl + i] - ((i == M.n_eps + 2) ? Dd.muhat[sn] : 0.f);
                if (k == TV_IM) v = 1.f;
            }
            wl[e] = v;
        }
    }
    __syncthreads();
    train_globals_dev(M, Tr, LP, TP, c, W.v(TV_CZ), S, tmat);
    if (tid == 0) {
        TrainCtl s{};
        s.act = ACT_INIT;
        s.phase = 0;
        s.step = Nt.cfg.init_step;
        s.da_prox = logf(10.f * s.step);
        s.eps_dir = 0.f;
        s.n_max = 1;
        Nt.ctl[c] = s;
    }
}

}  // namespace bayesn
