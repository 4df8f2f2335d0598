// bayesn_device.cuh - device code of the BayeSN hot path for sm_100a.
//
// One warp owns one (supernova, chain) pair.  Observations are processed two at a time: 16 lanes
// (even lanes / odd lanes) over the wavelength bins of each, only over the band's non-zero support;
// the per-observation sums (flux and the adjoint moments) of both are combined with one transposed
// warp-shuffle reduction.  A CTA is 4 warps; the warps of a CTA that work on the
// same supernova share its band-weight tile, which is staged in shared memory with a 1-D TMA
// bulk copy (cp.async.bulk + mbarrier) together with the wavelength-spline matrix J_l.
//
// Math (reference bayesn/bayesn_model.py:556-709, :761-816, :909-945):
//   W      = W0 + theta W1 + eps,   eps = reshape_F(L_Sigma eps_tform) with zero end rows
//   u_o[l] = -c2 sum_m W[l][m] Jt_o[m]                     c2 = 0.4 log2(10)
//   P_o[b] = w[k_o][b] H(t_o, b) 2^( sum_l J_l[b][l] u_o[l] - c2 A_V a[b] )
//   flux_o = 2^(-c2 (Ds - muhat)) sum_b P_o[b]             (band scale folded into w)
//   log L  = sum_o -((y_o - flux_o)/sigma_o)^2 / 2  + data constant
// and the analytic reverse mode of it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "bayesn_b200.h"

// -DBAYESN_BOUNDS_CHECK: device-side asserts on every address of the band-pass loop (lanes past the end
// of a band's support deliberately keep reading - and discarding - up to BAYESN_OVERRUN elements; this
// build proves those reads stay inside the allocations; compute-sanitizer is closed on the GPU pool)
#ifdef BAYESN_BOUNDS_CHECK
#include <assert.h>
#define BAYESN_ASSERT(x) assert(x)
#else
#define BAYESN_ASSERT(x)
#endif

namespace bayesn {

// -DBAYESN_PHASE_TIMING: clock64() deltas of the evaluator's phases, accumulated for warp 0 of CTA 0 (diagnostic
// build, scripts/phase_timing.py): 0 scalars, 1 eps + W (+ dust basis), 2 per-observation setup, 3 pair loop,
// 4 combine / team exchange, 5 reverse + priors, 6 sampler bookkeeping between evaluations, 7 evaluations;
// sub-phases: 8 eps = L_Sigma e (then 1 = W assembly), 9 dtheta (then 10 = L_Sigma^T dW, then 5 = priors).
#ifdef BAYESN_PHASE_TIMING
__device__ unsigned long long g_phase[16];
#define BAYESN_PHASE_BEGIN long long _pt = clock64();
#define BAYESN_PHASE(i)                                                          \
    do {                                                                         \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                               \
            const long long _t = clock64();                                      \
            atomicAdd(&g_phase[i], (unsigned long long)(_t - _pt));              \
            _pt = _t;                                                            \
        }                                                                        \
    } while (0)
#else
#define BAYESN_PHASE_BEGIN
#define BAYESN_PHASE(i)
#endif

constexpr int NB = BAYESN_NBINS;
constexpr int BP = BAYESN_BPAD;
constexpr int NPH = BAYESN_NPHASE;
// 8 warps per CTA, 2 CTAs per SM: the CTA-shared tables (J_l, dust basis: 18 KB) are amortised over 8
// warps, which leaves room for a per-warp slot of SN data (work-queue schedule of the fit sampler) and
// keeps the L1 carve-out at >= 60 KB; the training SN kernel gains 13 % (131 -> 114 us per pass at C4),
// logp+grad is unchanged (profiles/README.md).
#ifndef BAYESN_WPB
#define BAYESN_WPB 8
#endif
constexpr int WPB = BAYESN_WPB;                 // warps (= (SN, chain) units) per CTA
constexpr float C2 = 1.3287712379549449f;       // 0.4 * log2(10)
constexpr float KAPPA = 0.9210340371976184f;    // 0.4 * ln(10)
constexpr float LN2 = 0.6931471805599453f;
constexpr unsigned FULL = 0xffffffffu;

struct ModelDev {
    // Latency mode: kw = 2^kwl consecutive warps of a CTA form a team that evaluates ONE (SN, chain) unit, each
    // warp integrating the observation pairs p = rank (mod kw); set per launch by the host (kw = 1: one warp does
    // everything, the throughput layout).  xstride / xoff locate the teammates' exchange slots relative to a
    // warp's own scalar slot (floats); they live here - in the constant bank - so that the team logic costs the
    // one-warp path no registers.
    int kw, kwl, xstride, xoff;
    int L, T, n_eps, D, n_sc, rv_mode;
    int ls_lower;       // L_Sigma is exactly lower triangular (true for every shipped model): halve the mat-vecs
    int fix_tmax, fix_theta, fix_AV;
    float theta_val, AV_val;
    float inv_tauA, inv_sd2, RV, mu_R, sigma_R, phi_alpha;
    const float* Jl;    // [JROWS][RS] wavelength spline matrix, bin-major rows (zero padded)
    const float* hs;    // [NPH][BP] Hsiao template, phase-major
    const float* Mf;    // [9][BP]   F99 spline basis
    const float* ad;    // [JROWS]   dust basis for the fixed global RV (zero padded)
    const float* W0;    // [LP*TP]
    const float* W1;    // [LP*TP]
    const float* KD;    // [TP*TP]
    const float* tau;   // [TP]
    const float* Ls;    // [n_eps][n_eps] row-major L_Sigma
    const float* LsT;   // [n_eps][n_eps] its transpose
};

struct DataDev {
    int N, N_obs, n_b;
    const float4* obs4;     // [N][N_obs]
    const int* n_valid;     // [N]
    int wt_stride;          // floats per SN in the compact weight tiles (multiple of 4)
    const float* wt;        // [N][wt_stride]  band rows packed back to back over their supports
    const int4* sup;        // [N][n_b]        {first bin, one-past-last bin, offset of the row in the tile, 0}
    const float* muhat;     // [N]
    const float* lconst;    // [N]
    const float* muerr2;    // [N] training only: squared prior width of muhat from the redshift error (:1174-1175)
};

struct NutsDev {
    int num_warmup, num_samples, num_chains, max_depth;
    float target_accept, init_step;
    int adapt_step, adapt_mass, regularize;
    unsigned seed_lo, seed_hi;
    unsigned unit_offset;   // index of this launch's first (SN, chain) unit in the whole job: keys the random streams
    int n_windows;
    int win_end[24];
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Transposed reduction of 16 per-lane partials: 16 shuffles instead of 80.  On return every
// lane holds the warp total of value index (lane >> 1).
__device__ __forceinline__ float treduce16(float (&v)[16], int lane) {
    bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float send = up ? v[i] : v[i + 8];
        float keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float send = up ? v[i] : v[i + 4];
        float keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float send = up ? v[i] : v[i + 2];
        float keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    up = lane & 2;
    {
        float send = up ? v[0] : v[1];
        float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    return v[0];
}

// Packed FP32 arithmetic (Blackwell FFMA2 / FADD2: one issue slot for two IEEE fp32 operations on an
// aligned register pair), used for the two knot contractions of the band-pass loop and the adds of the transposed
// reduction: 36 instead of 47 instructions per bin round and bit-identical results (each half is an ordinary
// round-to-nearest FMA in the same order).  Round 1 measured it neutral (the loop was co-limited by the
// shared-memory pipe); on the round-2 build the same-box A/B (scripts/ab_compare.py, profiles/README.md) gives the
// sampler +3.5 % on C3 and +6.3 % on the dense C5 bands, the one-shot log p kernel -1 %: on by default,
// -DBAYESN_FFMA2=0 restores the scalar loop.
#ifndef BAYESN_FFMA2
#define BAYESN_FFMA2 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// The same for two interleaved groups of 16 lanes (even lanes / odd lanes work on two different
// observations): 15 shuffles; on return lane l holds the total of value (l >> 1) of its group.
__device__ __forceinline__ float treduce16h(float (&v)[16], int lane) {
    bool up = lane & 16;
#if BAYESN_FFMA2
    // the adds of each exchange step are done two at a time (FADD2)
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const float s0 = up ? v[i] : v[i + 8], s1 = up ? v[i + 1] : v[i + 9];
        const float k0 = up ? v[i + 8] : v[i], k1 = up ? v[i + 9] : v[i + 1];
        unpack2(fadd2(pack2(k0, k1), pack2(__shfl_xor_sync(FULL, s0, 16), __shfl_xor_sync(FULL, s1, 16))), v[i], v[i + 1]);
    }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        const float s0 = up ? v[i] : v[i + 4], s1 = up ? v[i + 1] : v[i + 5];
        const float k0 = up ? v[i + 4] : v[i], k1 = up ? v[i + 5] : v[i + 1];
        unpack2(fadd2(pack2(k0, k1), pack2(__shfl_xor_sync(FULL, s0, 8), __shfl_xor_sync(FULL, s1, 8))), v[i], v[i + 1]);
    }
    up = lane & 4;
    {
        const float s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];
        const float k0 = up ? v[2] : v[0], k1 = up ? v[3] : v[1];
        unpack2(fadd2(pack2(k0, k1), pack2(__shfl_xor_sync(FULL, s0, 4), __shfl_xor_sync(FULL, s1, 4))), v[0], v[1]);
    }
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float send = up ? v[i] : v[i + 8];
        float keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float send = up ? v[i] : v[i + 4];
        float keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float send = up ? v[i] : v[i + 2];
        float keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
#endif
    up = lane & 2;
    {
        float send = up ? v[0] : v[1];
        float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    return v[0];
}

// named barrier over the kw warps of a team (ids 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void team_sync(int team, int kw) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(kw * 32) : "memory");
}

// mbarrier + 1-D TMA bulk copy (global -> shared)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// Philox4x32-10 counter RNG (Salmon et al. 2011): the same stream is reproduced on the CPU by
// oracle/ref_nuts.py so that GPU and oracle trajectories can be compared draw for draw.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }  // [0,1)

// Fitzpatrick (1999) spline ordinates yk(R_V) - R_V at the nine knots xk (reference :604-623; the constant
// "- R_V" of the optical/IR knots is folded in: 1.00270 R - R = 0.00270 R) and their R_V derivatives.
// One definition for the fit evaluator (per-chain R_V), the forward kernels and the flux-from-chains kernel.
__device__ __forceinline__ void f99_yk(float R, float (&yk)[9], float (&dyk)[9]) {
    const float R2 = R * R, R3 = R2 * R;
    const float c2f = -0.824f + 4.717f / R, c1f = 2.030f - 3.007f * c2f;
    const float dc2 = -4.717f / R2, dc1 = -3.007f * dc2;
    const float x7 = 1e4f / 2700.f, x8 = 1e4f / 2600.f, x0 = 4.596f, gm = 0.99f;
    const float d1 = x7 * x7 / ((x7 * x7 - x0 * x0) * (x7 * x7 - x0 * x0) + gm * gm * x7 * x7);
    const float d2 = x8 * x8 / ((x8 * x8 - x0 * x0) * (x8 * x8 - x0 * x0) + gm * gm * x8 * x8);
    yk[0] = -R;                                   dyk[0] = -1.f;
    yk[1] = 0.26469f * R / 3.1f - R;              dyk[1] = 0.26469f / 3.1f - 1.f;
    yk[2] = 0.82925f * R / 3.1f - R;              dyk[2] = 0.82925f / 3.1f - 1.f;
    yk[3] = -0.422809f + 0.00270f * R + 2.13572e-4f * R2;     dyk[3] = 0.00270f + 2.f * 2.13572e-4f * R;
    yk[4] = -5.13540e-2f + 0.00216f * R - 7.35778e-5f * R2;   dyk[4] = 0.00216f - 2.f * 7.35778e-5f * R;
    yk[5] = 0.700127f + 0.00184f * R - 3.32598e-5f * R2;      dyk[5] = 0.00184f - 2.f * 3.32598e-5f * R;
    yk[6] = 1.19456f + 0.01707f * R - 5.46959e-3f * R2 + 7.97809e-4f * R3 - 4.45636e-5f * R2 * R2;
    dyk[6] = 0.01707f - 2.f * 5.46959e-3f * R + 3.f * 7.97809e-4f * R2 - 4.f * 4.45636e-5f * R3;
    yk[7] = c1f + c2f * x7 + 3.23f * d1;          dyk[7] = dc1 + dc2 * x7;
    yk[8] = c1f + c2f * x8 + 3.23f * d2;          dyk[8] = dc1 + dc2 * x8;
}

// ------------------------------------------------------------------------------------------
// shared-memory plan
// ------------------------------------------------------------------------------------------
template <int LP, int TP, bool TRAIN = false>
struct Dims {
    static constexpr int LP4 = (LP + 3) & ~3;
    // per-observation record in shared memory:
    //   [u (LP4) | J_t (TP) | J_t' (TP) | pad | frac, hsiao idx + flags + band, y, 1/sigma]
    // (training has no t_max: J_t' is dropped and the record is kept as small as possible so that
    //  four CTAs - every SN of a 500-SN training set in ONE wave - fit an SM's shared memory)
    static constexpr int R_JT = LP4, R_DJ = LP4 + TP, R_SC = (LP4 + (TRAIN ? 1 : 2) * TP + 3) & ~3;
    static constexpr int REC_RAW = R_SC + 4;
    static constexpr int REC = (TRAIN || (REC_RAW / 4) % 2 == 1) ? REC_RAW : REC_RAW + 4;   // odd #16B words: fewer conflicts
    static constexpr int NPAIR = (LP * TP + 31) / 32;   // (l, m) pairs per lane
    // J_l is stored bin-major, one row of RS floats per wavelength bin, read as LDS.128; an odd
    // number of 16-byte words per row keeps the 8-lane phases of LDS.128 bank-conflict free
    static constexpr int RS = ((LP4 / 4) % 2 == 1) ? LP4 : LP4 + 4;
};
// Observations are set up and integrated in chunks of 32 (one setup pass with lanes over observations, then
// 16 pairs): the per-warp record region does not grow with the light-curve length, so long light curves keep
// their occupancy (and the C5-sized per-warp slots of the work queue fit twice per SM).
constexpr int OBS_CHUNK = 32;
constexpr int JROWS = BP + 32;   // per-bin tables carry 32 zero rows so the last round may overrun
constexpr int OVERRUN = BAYESN_OVERRUN;     // bins a lane may read past the end of its band's support (values discarded)
// Row stride of the fit model's L_Sigma tables: a compile-time constant per kernel instantiation, so the
// column walks of the two triangular mat-vecs use immediate offsets (a runtime stride cost two integer
// instructions per load: 4 % of the evaluator's instructions in the r01i profile).  Training keeps
// per-chain tables with stride n_eps.
__host__ __device__ constexpr int ls_row_stride(int LP, int TP) { return ((LP - 2) * TP + 3) & ~3; }
// The fit kernels keep ONE copy of L_Sigma in shared memory with an ODD row stride: lanes over rows walking along
// their row (eps = L e) and lanes over columns walking down theirs (L^T dW) are both bank-conflict free.  Rows are
// padded to a multiple of 8 (zeros) so the 8-wide mat-vec loops need no remainder.  Through L1 the same walks cost a
// lone warp ~27 cycles per multiply-add even with eight loads in flight (scripts/phase_timing.py, r02c).
__host__ __device__ constexpr int ls_smem_stride(int LP, int TP) { return ((LP - 2) * TP) | 1; }
__host__ __device__ constexpr int ls_smem_rows(int LP, int TP) { return ((LP - 2) * TP + 7) & ~7; }
__host__ __device__ constexpr int jl_row_stride(int LP) { return ((((LP + 3) & ~3) / 4) % 2 == 1) ? ((LP + 3) & ~3) : ((LP + 3) & ~3) + 4; }
__host__ __device__ constexpr int rec_stride(int LP, int TP, bool train = false) {
    int raw = ((((LP + 3) & ~3) + (train ? 1 : 2) * TP + 3) & ~3) + 4;
    return (train || (raw / 4) % 2 == 1) ? raw : raw + 4;
}

struct SmemPlan {
    // CTA-shared (float offsets)
    int oJl, oAd, oMf, oW0, oW1, oKD, oTau, oLs, oWt, oSup, oObs, oBar;
    int oWarp, warpStride;     // per-warp region
    // inside the per-warp region
    int wWs, wEv, wEps, wRec, wSc, wAd, wDad, wX;
    int totalFloats;
};

// floats one warp of a team publishes per evaluation: its dW rows + {log L (double), gDs, gAV, gT, gRV}
__host__ __device__ constexpr int xchg_words(int LP, int TP) { return ((LP * TP + 6) + 3) & ~3; }

__host__ __device__ inline SmemPlan make_plan(int LP, int TP, int n_eps, int n_b, int N_obs, int wt_stride, int ntile,
                                              bool rvfree, bool stage, bool train = false, int kw = 1, bool ls_smem = false) {
    SmemPlan p;
    const int rec = rec_stride(LP, TP, train);
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    p.oJl = take(JROWS * jl_row_stride(LP));
    p.oAd = take(rvfree ? 0 : JROWS);
    p.oMf = take(rvfree ? 9 * BP : 0);
    p.oW0 = take(LP * TP);
    p.oW1 = take(LP * TP);
    p.oKD = take(TP * TP);
    p.oTau = take(TP);
    p.oLs = take((train || !ls_smem) ? 0 : ls_smem_rows(LP, TP) * ls_smem_stride(LP, TP) + 8);      // L_Sigma, odd row stride
    p.oWt = take(stage ? ntile * wt_stride : 0);
    p.oSup = take(ntile * n_b * 4);
    p.oObs = take(ntile * N_obs * 4);
    p.oBar = take(4);
    p.oWarp = o;
    int w = 0;
    auto wtake = [&](int n) { int r = w; w += (n + 3) & ~3; return r; };
    p.wWs = wtake(LP * TP);
    p.wEv = wtake((n_eps + 7) & ~7);     // zero beyond n_eps: the 8-wide mat-vec loops read the padding
    p.wEps = wtake(n_eps);
    p.wRec = wtake((((N_obs + 1) & ~1) < OBS_CHUNK ? ((N_obs + 1) & ~1) : OBS_CHUNK) * rec);    // one chunk of observations (pairs)
    p.wSc = wtake(8);
    p.wX = wtake(kw > 1 ? 2 * xchg_words(LP, TP) : 0);      // double-buffered exchange slot (latency mode)
    p.wAd = wtake(rvfree ? JROWS : 0);
    p.wDad = wtake(rvfree ? JROWS : 0);
    p.warpStride = w;
    // Lanes that have run past their band's support keep reading (and discard) up to OVERRUN bins
    // beyond it while the other observation of the pair finishes a wider band: keep those reads inside the CTA's
    // allocation (only the per-chain dust arrays of the sampled-RV fits sit at its very end).
    p.totalFloats = o + WPB * w + (rvfree ? OVERRUN : 0);
    const int jl_reach = p.oJl + (NB + OVERRUN) * jl_row_stride(LP);
    if (p.totalFloats < jl_reach) p.totalFloats = jl_reach;
    return p;
}

// Everything one warp needs to evaluate log p and its gradient.
struct WarpCtx {
    const float *sJl, *sAd, *sMf, *sW0, *sW1, *sKD, *sTau, *sLs;
    const float* wt_s;    // this SN's weight tile staged in shared memory (STAGE) ...
    const float* wt_g;    // ... or in global memory (read through L1)
    const int4* sup;      // this SN's band supports (shared)
    const float4* obs;    // this SN's observations (shared)
    float *Ws, *ev, *eps, *rec, *sc, *wad, *wdad;
    int nobs;
    float muhat, lconst;
    const float *Ls, *LsT;          // L_Sigma (row-major) and its transpose: model tables (fit) or this chain's (training)
    // training (train_model_globalRV): this chain's global parameters and where the per-SN
    // contributions to their gradient go
    const float *ad_g, *dad_g;      // dust basis a[b](RV) and da/dRV, [JROWS] each, global memory
    float sigma0, tauA, muerr2;
    float* part;                    // [LP*TP + 5]: dW, d/dsigma0, d/dRV, d/dtauA, log p (two floats of a double)
    const float *smem_lo, *smem_hi, *wt_lo, *wt_hi, *hs_hi;   // allocation bounds (BAYESN_BOUNDS_CHECK builds)
};

// ------------------------------------------------------------------------------------------
// time-spline row and its derivative (reference spline_coeffs_irr_step, :761-816)
// ------------------------------------------------------------------------------------------
template <int TP>
__device__ __forceinline__ void spline_row(float t, int T, const float* tau, const float* KD, float (&J)[TP],
                                           float (&dJ)[TP]) {
    const float tl = tau[T - 1], tf = tau[0];
    int q;
    float a, b, c, d, da, dc, dd;   // row = a e_q + b e_{q+1} + c KD[q] + d KD[q+1]
    if (t > tl) {
        q = T - 2;
        float h = tl - tau[T - 2];
        a = (tl - t) / h;
        b = 1.f - a;
        c = (t - tl) * h * (1.f / 6.f);
        d = 0.f;
        da = -1.f / h;
        dc = h * (1.f / 6.f);
        dd = 0.f;
    } else if (t < tf) {
        q = 0;
        float h = tau[1] - tf;
        b = (t - tf) / h;
        a = 1.f - b;
        c = 0.f;
        d = -(t - tf) * h * (1.f / 6.f);
        da = -1.f / h;
        dc = 0.f;
        dd = -h * (1.f / 6.f);
    } else {
        q = 0;
#pragma unroll
        for (int m = 1; m < TP - 1; ++m)
            if (m < T - 1 && tau[m] <= t) q = m;
        float h = tau[q + 1] - tau[q];
        a = (tau[q + 1] - t) / h;
        b = 1.f - a;
        float h2 = h * h * (1.f / 6.f);
        c = (a * a * a - a) * h2;
        d = (b * b * b - b) * h2;
        da = -1.f / h;
        dc = -(3.f * a * a - 1.f) * h * (1.f / 6.f);
        dd = (3.f * b * b - 1.f) * h * (1.f / 6.f);
    }
    const float* k0 = KD + q * TP;
    const float* k1 = KD + (q + 1) * TP;
#pragma unroll
    for (int m = 0; m < TP; ++m) {
        float v = fmaf(c, k0[m], d * k1[m]);
        float dv = fmaf(dc, k0[m], dd * k1[m]);
        if (m == q) { v += a; dv += da; }
        if (m == q + 1) { v += b; dv -= da; }
        J[m] = v;
        dJ[m] = dv;
    }
}

// sum_{j0 <= j < j1} col[(j - j0) * stride] * v[j]: one lane's share of a triangular L_Sigma mat-vec (lanes over
// rows, the walk down a column is coalesced across the lanes).  Eight loads in flight and four independent FMA
// chains: a lone warp (latency mode, small batches) no longer waits a full load latency per term - the single
// chain cost ~30 cycles per multiply-add, 2 x 2 300 cycles per evaluation of a W22 fit (scripts/phase_timing.py).
// j0 is a multiple of 4 and v 16-byte aligned.
__device__ __forceinline__ float column_dot(const float* __restrict__ col, int stride, const float* __restrict__ v, int j0, int j1) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = j0;
    for (; j + 8 <= j1; j += 8, col += 8 * stride) {
        const float l0 = __ldg(col), l1 = __ldg(col + stride), l2 = __ldg(col + 2 * stride), l3 = __ldg(col + 3 * stride);
        const float l4 = __ldg(col + 4 * stride), l5 = __ldg(col + 5 * stride), l6 = __ldg(col + 6 * stride), l7 = __ldg(col + 7 * stride);
        const float4 e0 = *reinterpret_cast<const float4*>(v + j), e1 = *reinterpret_cast<const float4*>(v + j + 4);
        a0 = fmaf(l0, e0.x, a0); a1 = fmaf(l1, e0.y, a1); a2 = fmaf(l2, e0.z, a2); a3 = fmaf(l3, e0.w, a3);
        a0 = fmaf(l4, e1.x, a0); a1 = fmaf(l5, e1.y, a1); a2 = fmaf(l6, e1.z, a2); a3 = fmaf(l7, e1.w, a3);
    }
    for (; j < j1; ++j, col += stride) a0 = fmaf(__ldg(col), v[j], a0);
    return (a0 + a1) + (a2 + a3);
}

// The same walk over the shared-memory copy of L_Sigma: n8 terms (a multiple of 8; table and vector are zero padded).
template <int STRIDE>
__device__ __forceinline__ float smem_dot(const float* col, const float* v, int n8) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = 0; j < n8; j += 8, col += 8 * STRIDE) {
        const float l0 = col[0], l1 = col[STRIDE], l2 = col[2 * STRIDE], l3 = col[3 * STRIDE];
        const float l4 = col[4 * STRIDE], l5 = col[5 * STRIDE], l6 = col[6 * STRIDE], l7 = col[7 * STRIDE];
        const float4 e0 = *reinterpret_cast<const float4*>(v + j), e1 = *reinterpret_cast<const float4*>(v + j + 4);
        a0 = fmaf(l0, e0.x, a0); a1 = fmaf(l1, e0.y, a1); a2 = fmaf(l2, e0.z, a2); a3 = fmaf(l3, e0.w, a3);
        a0 = fmaf(l4, e1.x, a0); a1 = fmaf(l5, e1.y, a1); a2 = fmaf(l6, e1.z, a2); a3 = fmaf(l7, e1.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// element i of a lane-strided vector (element i lives in lane i%32, slot i/32), broadcast
template <int VPL>
__device__ __forceinline__ float vget(const float (&v)[VPL], int i) {
    float x = v[0];
#pragma unroll
    for (int s = 1; s < VPL; ++s)
        if ((i >> 5) == s) x = v[s];
    return __shfl_sync(FULL, x, i & 31);
}

// ------------------------------------------------------------------------------------------
// log p and gradient for one (SN, chain): the unit of work of one leapfrog step
// ------------------------------------------------------------------------------------------
// TRAIN = false: fit_model_globalRV / popRV / uniformRV (reference :883-945, :996-1060, :818-881),
//   q = [eps_tform, theta, log A_V, logit tmax, Ds - muhat, (RV transform)].
// TRAIN = true: the per-SN part of train_model_globalRV (:1153-1182) for this chain's globals
//   (W0, W1, L_Sigma, sigma0, R_V, tau_A), q = [eps_tform, theta, log A_V, Ds - muhat]; t_max is
//   fixed at 0, the likelihood is Gaussian in magnitudes, and besides the latent gradient the
//   un-reduced contributions to the global gradient are written to c.part.
// LSMEM: L_Sigma is staged in shared memory (the persistent sampler); a one-shot evaluation and the training kernels
//   (per-chain tables) walk it in global memory through L1.
template <int LP, int TP, int VPL, bool RVFREE, bool STAGE, bool TRAIN = false, bool LSMEM = false>
__device__ __forceinline__ void eval_logp_grad(const ModelDev& M, const WarpCtx& c, const float* __restrict__ hs,
                                               const float (&q)[VPL], double& logp_out, float (&grad)[VPL], int lane) {
    using DM = Dims<LP, TP, TRAIN>;
    constexpr int LP4 = DM::LP4, REC = DM::REC, RS = DM::RS;
    constexpr bool DRV = RVFREE || TRAIN;     // the derivative with respect to R_V is needed
    static_assert(!(RVFREE && TRAIN), "training samples one global R_V");
    const int L = M.L, T = M.T, ne = M.n_eps;
    const int lss = TRAIN ? ne : ls_row_stride(LP, TP);     // row stride of c.Ls / c.LsT
    BAYESN_PHASE_BEGIN

    // ---- scalars ------------------------------------------------------------------------
    const float theta_s = vget<VPL>(q, ne + 0);
    const float uAV = vget<VPL>(q, ne + 1);
    const float ut = TRAIN ? 0.f : vget<VPL>(q, ne + 2);
    // the distance element is the SHIFTED distance dD = Ds - muhat: with Ds ~ 35 the float32 spacing
    // (3.8e-6) times the gradient far from the mode (1e6) is a potential-energy noise of order 1,
    // enough to stall step-size adaptation; around 0 the spacing is 30x finer.
    const float dD = vget<VPL>(q, ne + (TRAIN ? 2 : 3));
    const float AV_s = expf(uAV);
    const float sg = 1.f / (1.f + expf(-ut));
    const float tmax_s = -10.f + 20.f * sg;
    const float theta = (!TRAIN && M.fix_theta) ? M.theta_val : theta_s;
    const float AV = (!TRAIN && M.fix_AV) ? M.AV_val : AV_s;
    const float tmax = (TRAIN || M.fix_tmax) ? 0.f : tmax_s;
    float RVv = M.RV, dRV_du = 0.f, sgR = 0.f;
    if (RVFREE) {
        const float uR = vget<VPL>(q, ne + 4);
        sgR = 1.f / (1.f + expf(-uR));
        if (M.rv_mode == BAYESN_RV_UNIFORM) {
            RVv = 1.f + 5.f * sgR;
            dRV_du = 5.f * sgR * (1.f - sgR);
        } else {
            float pp = M.phi_alpha + sgR * (1.f - M.phi_alpha);
            float z = normcdfinvf(pp);
            RVv = M.mu_R + M.sigma_R * z;
            float pdf = 0.3989422804014327f * expf(-0.5f * z * z);
            dRV_du = M.sigma_R * (1.f - M.phi_alpha) * sgR * (1.f - sgR) / pdf;
        }
    }
    const float AVc = -C2 * AV;
    const float S = exp2f(-C2 * dD);

    BAYESN_PHASE(0);
    // ---- eps = L_Sigma eps_tform, W = W0 + theta W1 + eps ---------------------------------
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        int i = lane + 32 * s;
        if (i < ne) c.ev[i] = q[s];
    }
    __syncwarp();
    // (Splitting the two L_Sigma products over the team as well - a slice of the terms per warp, partials exchanged behind
    // one more barrier each - was measured slower than the replicated products: +0.7 k cycles per evaluation at K = 4,
    // profiles/README.md r02k.  A team barrier costs more than the ~50 multiply-adds it saves.)
    for (int i0 = 0; i0 < ne; i0 += 32) {
        const int i = i0 + lane;
        const int jend = (TRAIN || M.ls_lower) ? min(ne, i0 + 32) : ne;      // eps_i = sum_{j <= i} L[i][j] e_j
        if (i < ne) {
            if constexpr (TRAIN || !LSMEM) {
                // tables in global memory (per chain in training); pointer walk down the column, coalesced over the
                // lanes, immediate offsets thanks to the compile-time stride
                c.eps[i] = column_dot(c.LsT + i, lss, c.ev, 0, jend);
            } else {
                // lane i walks along row i of the shared-memory copy (odd row stride: conflict free)
                c.eps[i] = smem_dot<1>(c.sLs + i * ls_smem_stride(LP, TP), c.ev, (jend + 7) & ~7);
            }
        }
    }
    __syncwarp();
    BAYESN_PHASE(8);
    for (int p = lane; p < LP * TP; p += 32) {
        int l = p / TP, m = p - l * TP;
        float w = 0.f;
        if (l < L && m < T) {
            w = fmaf(theta, c.sW1[p], c.sW0[p]);
            if (l >= 1 && l <= L - 2) w += c.eps[(l - 1) + (L - 2) * m];
        }
        c.Ws[p] = w;
    }
    const float* adv = TRAIN ? c.ad_g : c.sAd;
    if (RVFREE) {
        // per-chain dust basis a[b] = 1 + (M_fitz yk(RV))/RV and its RV derivative (reference :604-625)
        float yk[9], dyk[9];
        f99_yk(RVv, yk, dyk);
        const float iR = 1.f / RVv;
        for (int b = lane; b < JROWS; b += 32) {
            float my = 0.f, dmy = 0.f;
            if (b < NB) {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    float mf = c.sMf[j * BP + b];
                    my = fmaf(mf, yk[j], my);
                    dmy = fmaf(mf, dyk[j], dmy);
                }
            }
            c.wad[b] = (b < NB) ? 1.f + my * iR : 0.f;
            c.wdad[b] = (b < NB) ? (dmy - my * iR) * iR : 0.f;
        }
        adv = c.wad;
    }
    __syncwarp();

    // ---- per-observation setup, lanes over observations -------------------------------------
    // Observations are processed in pairs (16 interleaved lanes each); an odd count is completed by a
    // dummy copy of the last observation with 1/sigma = 0, which contributes exactly nothing.
    // Latency mode (M.kw > 1): this warp is rank tk of a team of kw warps that all hold the same position; it
    // integrates the pairs p = tk (mod kw) and the teams' partial sums are exchanged after the loop.  Pairs are
    // sorted by band width, so the ranks get equal work.  kw = 1: nlp = npair and every index below is the identity.
    const int tk = (threadIdx.x >> 5) & (M.kw - 1);
    const int npair_all = (c.nobs + 1) >> 1;
    const int npair = (npair_all + M.kw - 1 - tk) >> M.kwl;      // pairs of this warp
    if (lane == 0) c.sc[0] = S;
    // log L is accumulated in double: far from the mode |log L| reaches 1e6 while NUTS needs energy
    // differences to ~1e-2, which float32 partial sums cannot hold (ulp(7.5e5) = 0.06)
    double logL = 0.0;
    float gDs = 0.f, gAV = 0.f, gT = 0.f, gRV = 0.f;
    const float* wtile = STAGE ? c.wt_s : c.wt_g;
    const int hl = lane >> 1, hb = lane & 1;
    // lane hl = 4 + l ends every reduction holding Q_l; it also keeps row l of W for the d/dtmax term
    float wrow[TP];
#pragma unroll
    for (int m = 0; m < TP; ++m) wrow[m] = (hl >= 4 && hl < 4 + LP) ? c.Ws[(hl - 4) * TP + m] : 0.f;
    float gTW = 0.f;
    // dW[l][.] = sum_o dU_o[l] J_t,o[.] accumulates in the lane that ends each pair's reduction holding dU_o[l]
    // (one row of 6 per lane and half, added across the halves after the loop) - a separate pass over all
    // observations with lanes over (l, m) cost 5 % of the evaluation in the r01i profile
    float dWr[TP];
#pragma unroll
    for (int m = 0; m < TP; ++m) dWr[m] = 0.f;
    const bool qlane = hl >= 4 && hl < 4 + LP;
    BAYESN_PHASE(1);
    for (int o0 = 0; o0 < 2 * npair; o0 += OBS_CHUNK) {
        const int o = o0 + lane;
        const unsigned act = __ballot_sync(FULL, o < 2 * npair);     // both lanes of a pair are in or out together
        if (o < 2 * npair) {
            const int go = ((((o >> 1) << M.kwl) + tk) << 1) | (o & 1);      // observation index in the light curve
            const bool dummy = go >= c.nobs;
            const float4 ob = c.obs[dummy ? c.nobs - 1 : go];
            float t = ob.x - tmax;
            float J[TP], dJ[TP];
            spline_row<TP>(t, T, c.sTau, c.sKD, J, dJ);
            float* rc = c.rec + lane * REC;
            // the record is assembled in registers and written with 128-bit stores: the record stride (an odd
            // number of 16-byte words) makes those conflict-free, while 32-bit stores at a stride of REC words
            // were 4-way bank-conflicted (r01i profile)
            float rv[DM::R_SC];
#pragma unroll
            for (int i = 0; i < DM::R_SC; ++i) rv[i] = 0.f;
#pragma unroll
            for (int m = 0; m < TP; ++m) {
                rv[DM::R_JT + m] = J[m];
                if (!TRAIN) rv[DM::R_DJ + m] = dJ[m];
            }
#pragma unroll
            for (int l = 0; l < LP4; ++l) {
                float u = 0.f;
                if (l < LP) {
                    if constexpr (TP % 2 == 0) {
                        // rows of W are 8-byte aligned: half as many (broadcast) shared-memory loads
                        const float2* wr = reinterpret_cast<const float2*>(c.Ws + l * TP);
#pragma unroll
                        for (int m = 0; m < TP; m += 2) {
                            const float2 w2 = wr[m >> 1];
                            u = fmaf(w2.x, J[m], u);
                            u = fmaf(w2.y, J[m + 1], u);
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < TP; ++m) u = fmaf(c.Ws[l * TP + m], J[m], u);
                    }
                }
                rv[l] = -C2 * u;
            }
#pragma unroll
            for (int i = 0; i < DM::R_SC; i += 4)
                *reinterpret_cast<float4*>(rc + i) = make_float4(rv[i], rv[i + 1], rv[i + 2], rv[i + 3]);
            float fl = floorf(t);
            int i0 = 19 + (int)fl, i1 = 19 + (int)ceilf(t);
            i0 = min(max(i0, 0), NPH - 1);
            i1 = min(max(i1, 0), NPH - 1);
            // the upper Hsiao row is read at i0 + 1; when ceil == floor (integer phase) or the index
            // is clamped the interpolation degenerates to row i0 and d/dt of it vanishes
            const bool interp = i1 != i0;
            // rounds of the band-pass loop for this PAIR (the wider of its two supports), so that the pair
            // prologue does not wait for a support load and a shuffle before it knows its trip count
            const int band = __float_as_int(ob.w);
            const int4 spo = c.sup[band];
            int nro = (spo.y - spo.x + 15) >> 4;
            nro = max(nro, __shfl_xor_sync(act, nro, 1));
            // Mixed-band pairs: the even lanes read J_l rows first_A + hl, the odd lanes first_B + hl, and a row is three
            // 16-byte words, so within each 8-lane phase of the LDS.128 the two sets of four rows hit the same bank
            // groups unless first_B - first_A = 4 (mod 8) (39 % extra wavefronts on these loads in the r02i profile).
            // The second observation therefore starts its lanes `rot` bins into its support (lane hl takes bin
            // first + ((hl + rot) & 15)): the same bins, another lane order.  Same-band pairs keep rot = 0 (broadcast).
            const int pband = __shfl_xor_sync(act, band, 1), pfirst = __shfl_xor_sync(act, spo.x, 1);
            const int rot = ((lane & 1) && pband != band) ? ((4 - (spo.x - pfirst)) & 7) : 0;
            *reinterpret_cast<float4*>(rc + DM::R_SC) =
                make_float4(interp ? (t - fl) : 0.f,
                            __int_as_float(i0 | (interp ? 256 : 0) | (nro << 9) | (band << 16) | (rot << 24)),
                            dummy ? 0.f : ob.y, dummy ? 0.f : ob.z);
        }
        __syncwarp();
        BAYESN_PHASE(2);

        // ---- band-pass integrals: two observations per pass, 16 lanes over wavelength bins each ---------
        // The lanes of the two observations are INTERLEAVED (even lanes: first, odd lanes: second): when both
        // are in the same band - the common case after sorting - neighbouring lanes read the same J_l row
        // and weight, which the shared-memory pipe serves as one broadcast, halving its traffic (the LSU
        // pipe, not the FMA pipe, is the busiest unit of this loop).
        const int p_end = min(npair, (o0 + OBS_CHUNK) >> 1);
        for (int p = o0 >> 1; p < p_end; ++p) {
            float* rc = c.rec + (2 * p - o0 + hb) * REC;
            float u[LP4];
            {
                const float4* up = reinterpret_cast<const float4*>(rc);
#pragma unroll
                for (int j = 0; j < LP4 / 4; ++j) {
                    float4 a = up[j];
                    u[4 * j] = a.x; u[4 * j + 1] = a.y; u[4 * j + 2] = a.z; u[4 * j + 3] = a.w;
                }
            }
            const float4 scal = *reinterpret_cast<const float4*>(rc + DM::R_SC);
            const float f = scal.x;
            const int pk = __float_as_int(scal.y);
            const float yo = scal.z, iso = scal.w;
            const int4 sp = c.sup[(pk >> 16) & 0xff];
            const int nr = (pk >> 9) & 0x7f;          // <= 19 rounds of 16 bins
            // running pointers: one add per array per round instead of rebuilding addresses
            const int boff = (hl + (pk >> 24)) & 15;  // this lane's first bin in the support (rotated for mixed pairs)
            const int b_first = sp.x + boff;
            const float* p0 = hs + (pk & 0xff) * BP + b_first;
            const float* pw = wtile + sp.z + boff;
            const float4* pj = reinterpret_cast<const float4*>(c.sJl) + b_first * (RS / 4);
            const float* pa = adv + b_first;
            const float* pda = RVFREE ? (c.wdad + b_first) : (TRAIN ? c.dad_g + b_first : nullptr);
            int rem = sp.y - b_first;                // this lane is inside the support while rem > 0
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#if BAYESN_FFMA2
            // knot pairs (l, l+1) share one packed FMA: u and the J_l row arrive as aligned pairs from LDS.128,
            // the even / odd halves are exactly the two FFMA chains of the scalar loop
            f32x2 u2[LP4 / 2], q2[LP4 / 2];
#pragma unroll
            for (int k = 0; k < LP4 / 2; ++k) { u2[k] = pack2(u[2 * k], u[2 * k + 1]); q2[k] = pack2(0.f, 0.f); }
#endif
            for (int it = 0; it < nr; ++it) {
#if BAYESN_FFMA2
                f32x2 jl2[LP4 / 2];
#pragma unroll
                for (int j = 0; j < LP4 / 4; ++j) {
                    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(pj)[j];
                    jl2[2 * j] = v.x; jl2[2 * j + 1] = v.y;
                }
#else
                float jl[LP4];
#pragma unroll
                for (int j = 0; j < LP4 / 4; ++j) {
                    const float4 v = pj[j];
                    jl[4 * j] = v.x; jl[4 * j + 1] = v.y; jl[4 * j + 2] = v.z; jl[4 * j + 3] = v.w;
                }
#endif
                BAYESN_ASSERT((const float*)pj >= c.smem_lo && (const float*)(pj + LP4 / 4) <= c.smem_hi);
                BAYESN_ASSERT(STAGE ? (pw >= c.smem_lo && pw < c.smem_hi) : (pw >= c.wt_lo && pw < c.wt_hi));
                BAYESN_ASSERT(p0 >= hs && p0 + BP < c.hs_hi);
                BAYESN_ASSERT(TRAIN || (pa >= c.smem_lo && pa < c.smem_hi));
                const float wl = STAGE ? *pw : __ldg(pw);
                const float h0 = __ldg(p0), h1 = __ldg(p0 + BP);
                const float a = TRAIN ? __ldg(pa) : *pa;
                // two independent FFMA chains (4-cycle dependent-issue latency)
                float g0 = 0.f, g1 = 0.f;
#if BAYESN_FFMA2
                {
                    f32x2 g2 = pack2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < (LP + 1) / 2; ++k) g2 = ffma2(jl2[k], u2[k], g2);
                    unpack2(g2, g0, g1);
                }
#else
#pragma unroll
                for (int l = 0; l < LP; l += 2) {
                    g0 = fmaf(jl[l], u[l], g0);
                    if (l + 1 < LP) g1 = fmaf(jl[l + 1], u[l + 1], g1);
                }
#endif
                const float dh = h1 - h0;
                const float H = fmaf(f, dh, h0);
                const float e = ex2_approx(fmaf(AVc, a, g0 + g1));
                const float we = (rem > 0) ? wl * e : 0.f;     // select, not multiply: e may be garbage past the support
                const float P = we * H;
                acc[0] += P;
                acc[1] = fmaf(P, a, acc[1]);
                acc[2] = fmaf(we, dh, acc[2]);
                if (DRV) { acc[3] = fmaf(P, TRAIN ? __ldg(pda) : *pda, acc[3]); pda += 16; }
#if BAYESN_FFMA2
                {
                    const f32x2 P2 = pack2(P, P);
#pragma unroll
                    for (int k = 0; k < (LP + 1) / 2; ++k) q2[k] = ffma2(jl2[k], P2, q2[k]);
                }
#else
#pragma unroll
                for (int l = 0; l < LP; ++l) acc[4 + l] = fmaf(jl[l], P, acc[4 + l]);
#endif
                p0 += 16; pw += 16; pj += 16 * (RS / 4); pa += 16; rem -= 16;
            }
#if BAYESN_FFMA2
#pragma unroll
            for (int k = 0; k < (LP + 1) / 2; ++k) {
                float lo, hi;
                unpack2(q2[k], lo, hi);
                acc[4 + 2 * k] = lo;
                if (4 + 2 * k + 1 < 16 && 2 * k + 1 < LP) acc[4 + 2 * k + 1] = hi;
            }
#endif
            const float val = treduce16h(acc, lane);
            const float F = __shfl_sync(FULL, val, hb);
            const float FA = __shfl_sync(FULL, val, hb | 2);
            const float FT = __shfl_sync(FULL, val, hb | 4);
            const float Sv = c.sc[0];
            const float flux = Sv * F;
            float chi, r;                      // r = d log L / d flux
            if (TRAIN) {
                // magnitude-space likelihood (get_mag_batch :711-759, :1178-1182): m = 27.5 - 2.5 log10 flux.
                // The host folds 10^(0.4 (ybar_s - 27.5)) into the weights and stores y - ybar_s, so that
                // both magnitudes are O(1) and their float32 difference keeps ~1e-7 mag (the gradient
                // multiplies it by n_obs / sigma^2)
                const float mg = -0.7525749891599529f * log2f(flux);
                chi = (yo - mg) * iso;
                r = chi * iso * (-1.0857362047581296f) / flux;
            } else {
                chi = (yo - flux) * iso;
                r = chi * iso;
            }
            logL += (double)(-0.5f * chi * chi);
            const float coef = r * Sv;
            gDs = fmaf(-KAPPA * r, flux, gDs);
            gAV = fmaf(-KAPPA * coef, FA, gAV);
            gT = fmaf((pk & 256) ? coef : 0.f, FT, gT);
            if (DRV) gRV = fmaf(-KAPPA * AV * coef, __shfl_sync(FULL, val, hb | 6), gRV);
            // d logL / d(sum_m W[l][m] Jt[m]) for this observation: its outer product with J_t goes into dW,
            // its product with (W J_t'(t))[l] into the gradient through the time spline
            const float dU = qlane ? -KAPPA * coef * val : 0.f;
            if (!TRAIN) {
                float du = 0.f;
#pragma unroll
                for (int m = 0; m < TP; ++m) du = fmaf(wrow[m], rc[DM::R_DJ + m], du);
                gTW = fmaf(dU, du, gTW);
            }
#pragma unroll
            for (int m = 0; m < TP; ++m) dWr[m] = fmaf(dU, rc[DM::R_JT + m], dWr[m]);
        }
        __syncwarp();      // the next chunk overwrites the records
        BAYESN_PHASE(3);
    }
#pragma unroll
    for (int m = 0; m < TP; ++m) dWr[m] += __shfl_xor_sync(FULL, dWr[m], 1);
    // the two halves hold the sums over their own observations
    logL += __shfl_xor_sync(FULL, logL, 1);
    gDs += __shfl_xor_sync(FULL, gDs, 1);
    gAV += __shfl_xor_sync(FULL, gAV, 1);
    gT += __shfl_xor_sync(FULL, gT, 1);
    if (DRV) gRV += __shfl_xor_sync(FULL, gRV, 1);
    __syncwarp();

    gT += wsum(gTW);

    if (M.kw > 1) {
        // ---- team exchange: every warp publishes its partial sums, all add them in rank order ----------
        // (identical additions in every warp -> the replicated sampler states stay bit-identical).  The slots
        // are double-buffered by the evaluation's parity, so ONE barrier per evaluation is enough: a warp can
        // only overwrite a slot after passing the next barrier, which every teammate reaches after its reads.
        constexpr int XW = xchg_words(LP, TP);
        const int par = __float_as_int(c.sc[1]);
        if (lane == 0) c.sc[1] = __int_as_float(par ^ XW);
        float* mine = c.sc + M.xoff + par;
        if (qlane && hb == 0) {
#pragma unroll
            for (int m = 0; m < TP; ++m) mine[(hl - 4) * TP + m] = dWr[m];
        }
        if (lane == 0) {
            float* ps = mine + LP * TP;
            ps[0] = __int_as_float(__double2loint(logL));
            ps[1] = __int_as_float(__double2hiint(logL));
            ps[2] = gDs; ps[3] = gAV; ps[4] = gT; ps[5] = DRV ? gRV : 0.f;
        }
        team_sync((threadIdx.x >> 5) >> M.kwl, M.kw);
        const float* base = mine - tk * M.xstride;
        logL = 0.0; gDs = 0.f; gAV = 0.f; gT = 0.f; gRV = 0.f;
#pragma unroll
        for (int m = 0; m < TP; ++m) dWr[m] = 0.f;
        for (int kk = 0; kk < M.kw; ++kk, base += M.xstride) {
            if (qlane && hb == 0) {
#pragma unroll
                for (int m = 0; m < TP; ++m) dWr[m] += base[(hl - 4) * TP + m];
            }
            const float* ps = base + LP * TP;
            logL += __hiloint2double(__float_as_int(ps[1]), __float_as_int(ps[0]));
            gDs += ps[2]; gAV += ps[3]; gT += ps[4];
            if (DRV) gRV += ps[5];
        }
    }

    BAYESN_PHASE(4);
    // ---- reverse: theta, eps_tform -----------------------------------------------------------
    // lane (hl = 4 + l, hb = 0) owns row l of dW
    float gth = 0.f;
    {
        const int l = hl - 4;
        if (qlane && hb == 0) {
#pragma unroll
            for (int m = 0; m < TP; ++m) {
                gth = fmaf(dWr[m], c.sW1[l * TP + m], gth);
                if (l >= 1 && l <= L - 2 && m < T) c.ev[(l - 1) + (L - 2) * m] = dWr[m];
            }
        }
    }
    gth = wsum(gth);
    __syncwarp();
    BAYESN_PHASE(9);

    // ---- priors, Jacobians, assemble gradient in the unconstrained space ---------------------
    double lp = logL + (double)c.lconst;
    float ss = 0.f;
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        int i = lane + 32 * s;
        float gi = 0.f;
        if (i < ne) {
            const int j0 = (TRAIN || M.ls_lower) ? 32 * s : 0;
            if constexpr (TRAIN || !LSMEM) gi = column_dot(c.Ls + i + (size_t)j0 * lss, lss, c.ev, j0, ne) - q[s];
            else gi = smem_dot<ls_smem_stride(LP, TP)>(c.sLs + i + j0 * ls_smem_stride(LP, TP), c.ev + j0, (ne - j0 + 7) & ~7) - q[s];
            ss = fmaf(q[s], q[s], ss);
        }
        grad[s] = gi;
    }
    ss = wsum(ss);
    BAYESN_PHASE(10);
    lp += -0.5f * ss;
    // theta ~ N(0,1)
    lp += -0.5f * theta_s * theta_s;
    if (TRAIN) {
        const float inv_tau = 1.f / c.tauA;
        // A_V = exp(u) ~ Exponential(1/tauA) (+ log|J| = u); the -log tauA term depends on a sampled global
        lp += -AV_s * inv_tau - logf(c.tauA) + uAV;
        const float g_uAV = (gAV - inv_tau) * AV_s + 1.f;
        // Ds ~ N(muhat, sqrt(muhat_err^2 + sigma0^2)) (:1174-1177)
        const float isd2 = 1.f / (c.muerr2 + c.sigma0 * c.sigma0);
        lp += -0.5f * dD * dD * isd2 + 0.5f * logf(isd2);
        const float g_Ds = gDs - dD * isd2;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < VPL; ++s) {
            int i = lane + 32 * s;
            if (i == ne + 0) grad[s] = gth - theta_s;
            if (i == ne + 1) grad[s] = g_uAV;
            if (i == ne + 2) grad[s] = g_Ds;
        }
        // un-reduced contributions to the gradient of the globals (constrained space); one writer per team
        if (qlane && hb == 0 && tk == 0) {
#pragma unroll
            for (int m = 0; m < TP; ++m) c.part[(hl - 4) * TP + m] = dWr[m];
        }
        if (lane == 0 && tk == 0) {
            float* ps = c.part + LP * TP;
            ps[0] = (dD * dD * isd2 - 1.f) * isd2 * c.sigma0;          // d/dsigma0
            ps[1] = gRV;                                                 // d/dRV
            ps[2] = (AV_s * inv_tau - 1.f) * inv_tau;                    // d/dtauA
            *reinterpret_cast<double*>(ps + 3 + ((LP * TP + 3) & 1)) = lp;
        }
        logp_out = lp;
        return;
    }
    float g_theta = (M.fix_theta ? 0.f : gth) - theta_s;
    // A_V = exp(u): Exponential(1/tauA) prior + log|J| = u
    lp += -AV_s * M.inv_tauA + uAV;
    float g_uAV = ((M.fix_AV ? 0.f : gAV) - M.inv_tauA) * AV_s + 1.f;
    // tmax = -10 + 20 sigmoid(u): uniform prior cancels the affine Jacobian
    lp += logf(sg) + logf(1.f - sg);
    float g_ut = (M.fix_tmax ? 0.f : -gT) * 20.f * sg * (1.f - sg) + (1.f - 2.f * sg);
    // Ds ~ N(muhat, sd)
    lp += -0.5f * dD * dD * M.inv_sd2;
    float g_Ds = gDs - dD * M.inv_sd2;
    float g_uR = 0.f;
    if (RVFREE) {
        lp += logf(sgR) + logf(1.f - sgR);
        g_uR = gRV * dRV_du + (1.f - 2.f * sgR);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        int i = lane + 32 * s;
        if (i == ne + 0) grad[s] = g_theta;
        if (i == ne + 1) grad[s] = g_uAV;
        if (i == ne + 2) grad[s] = g_ut;
        if (i == ne + 3) grad[s] = g_Ds;
        if (RVFREE && i == ne + 4) grad[s] = g_uR;
    }
    logp_out = lp;
    BAYESN_PHASE(5);
}

// ------------------------------------------------------------------------------------------
// CTA prologue shared by the log p kernel and the NUTS kernel: stage tables with TMA, carve smem
// ------------------------------------------------------------------------------------------
template <int LP, int TP, bool RVFREE, bool STAGE, bool TRAIN = false, bool LSMEM = false>
__device__ __forceinline__ bool cta_setup(const ModelDev& M, const DataDev& Dd, int n_chains, int ntile, float* smem,
                                          WarpCtx& c, int& sn, int& chain, bool per_warp = false) {
    const SmemPlan P = make_plan(LP, TP, M.n_eps, Dd.n_b, Dd.N_obs, Dd.wt_stride, ntile, RVFREE, STAGE, TRAIN, M.kw, LSMEM);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int upc = WPB >> M.kwl;                  // (SN, chain) units per CTA: one per team of kw warps
    const int team = warp >> M.kwl;
    const long long unit0 = (long long)blockIdx.x * upc;
    const long long total = (long long)Dd.N * n_chains;
    const int s_first = (int)(unit0 / n_chains);
    long long last_unit = unit0 + upc - 1;
    if (last_unit >= total) last_unit = total - 1;
    const int s_last = (int)(last_unit / n_chains);
    // per_warp: only the model tables are staged here; every warp stages the data of the SN it
    // takes from the work queue itself (warp_bind)
    const int nt = per_warp ? 0 : s_last - s_first + 1;

    float* sJl = smem + P.oJl;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P.oBar);
    // Lanes that have run past their band's support keep reading (and discarding) words beyond it: table padding,
    // neighbouring regions, slots of SN data that are staged later.  The products they form are multiplied by an
    // exact zero, so the only requirement is that no such word is a NaN / Inf pattern left behind by an earlier
    // kernel: clear the CTA's whole allocation once (generic-proxy stores, fenced before the TMA writes).
    {
        float4* z4 = reinterpret_cast<float4*>(smem);
        for (int i = tid; i < P.totalFloats / 4; i += blockDim.x) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = (P.totalFloats & ~3) + tid; i < P.totalFloats; i += blockDim.x) smem[i] = 0.f;
    }
    __syncthreads();
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        constexpr uint32_t kJlBytes = JROWS * Dims<LP, TP>::RS * 4;
        uint32_t bytes = kJlBytes + (RVFREE ? 9 * BP * 4 : JROWS * 4);
        bytes += nt * Dd.N_obs * 16;
        if (STAGE) bytes += nt * Dd.wt_stride * 4;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(sJl, M.Jl, kJlBytes, bar);
        if (RVFREE) tma_load_1d(smem + P.oMf, M.Mf, 9 * BP * 4, bar);
        else tma_load_1d(smem + P.oAd, M.ad, JROWS * 4, bar);
        if (nt) tma_load_1d(smem + P.oObs, Dd.obs4 + (size_t)s_first * Dd.N_obs, nt * Dd.N_obs * 16, bar);
        if (STAGE && nt) tma_load_1d(smem + P.oWt, Dd.wt + (size_t)s_first * Dd.wt_stride, nt * Dd.wt_stride * 4, bar);
    }
    // small tables with plain loads while the bulk copies are in flight
    for (int i = tid; i < LP * TP; i += blockDim.x) {
        smem[P.oW0 + i] = M.W0[i];
        smem[P.oW1 + i] = M.W1[i];
    }
    for (int i = tid; i < TP * TP; i += blockDim.x) smem[P.oKD + i] = M.KD[i];
    for (int i = tid; i < TP; i += blockDim.x) smem[P.oTau + i] = M.tau[i];
    if constexpr (!TRAIN && LSMEM) {
        const int ne = M.n_eps;
        constexpr int SS = ls_smem_stride(LP, TP), LSS = ls_row_stride(LP, TP);
        for (int e = tid; e < ne * ne; e += blockDim.x) {
            const int i = e / ne, j = e - i * ne;
            smem[P.oLs + i * SS + j] = M.Ls[i * LSS + j];
        }
    }
    int4* sSup = reinterpret_cast<int4*>(smem + P.oSup);
    for (int i = tid; i < nt * Dd.n_b; i += blockDim.x) sSup[i] = Dd.sup[(size_t)s_first * Dd.n_b + i];
    mbar_wait(bar, 0);
    __syncthreads();

    const long long unit = unit0 + team;
    const bool active = per_warp || unit < total;
    sn = (!per_warp && active) ? (int)(unit / n_chains) : s_first;
    chain = (!per_warp && active) ? (int)(unit - (long long)sn * n_chains) : 0;
    const int slot = per_warp ? team : sn - s_first;
    // Offsets are passed through an empty asm so that ptxas keeps each one in a register instead
    // of re-deriving the whole shared-memory plan from kernel parameters inside the hot loops
    // (seen in the round-1 profile: ~60 integer instructions per observation).
    auto at = [&](int off) -> float* {
        asm volatile("" : "+r"(off));
        return smem + off;
    };
    const int wbase = P.oWarp + warp * P.warpStride;
    c.sJl = at(P.oJl);
    c.sAd = at(P.oAd);
    c.sMf = smem + P.oMf;
    c.sW0 = smem + P.oW0;
    c.sW1 = smem + P.oW1;
    c.sKD = smem + P.oKD;
    c.sTau = smem + P.oTau;
    c.sLs = at(P.oLs);
    c.wt_s = at(P.oWt + slot * Dd.wt_stride);
    c.wt_g = Dd.wt + (size_t)sn * Dd.wt_stride;
    c.sup = reinterpret_cast<const int4*>(at(P.oSup + slot * Dd.n_b * 4));
    c.obs = reinterpret_cast<const float4*>(smem + P.oObs) + slot * Dd.N_obs;
    c.Ws = smem + wbase + P.wWs;
    c.ev = smem + wbase + P.wEv;
    c.eps = smem + wbase + P.wEps;
    c.rec = at(wbase + P.wRec);
    c.sc = at(wbase + P.wSc);
    if ((tid & 31) == 0) c.sc[1] = 0.f;          // parity of the team exchange slots (latency mode)
    c.wad = at(wbase + P.wAd);
    c.wdad = at(wbase + P.wDad);
    c.nobs = Dd.n_valid[sn];
    c.muhat = Dd.muhat[sn];
    c.lconst = Dd.lconst[sn];
    c.Ls = M.Ls;
    c.LsT = M.LsT;
    c.ad_g = c.dad_g = nullptr;
    c.part = nullptr;
    c.sigma0 = c.tauA = c.muerr2 = 0.f;
    c.smem_lo = smem;
    c.smem_hi = smem + P.totalFloats;
    c.wt_lo = Dd.wt;
    c.wt_hi = Dd.wt + (size_t)Dd.N * Dd.wt_stride + OVERRUN + 32;
    c.hs_hi = M.hs + (NPH + 1) * BP + OVERRUN + 32;
    return active;
}

// Work-queue mode (cta_setup with per_warp = true, plan made with ntile = WPB): the warp stages the
// observations, band supports and - when STAGE - the weight tile of SN `sn` into its own slot and
// points its context at them.  A few KB of coalesced loads per chain, i.e. per ~10^5 evaluations.
template <bool STAGE>
__device__ __forceinline__ void warp_bind(const ModelDev& M, const DataDev& Dd, const SmemPlan& P, float* smem, WarpCtx& c,
                                          int sn, int lane) {
    const int warp = threadIdx.x >> 5;
    const int slot = warp >> M.kwl;                // one slot of SN data per team
    __syncwarp();
    if ((warp & (M.kw - 1)) == 0) {                // the team's first warp stages, the others wait at the barrier
        // Plain coalesced 128-bit loads: a few KB once per ~10^5 evaluations.  (A per-slot cp.async.bulk + mbarrier
        // version of this staging faulted on B200 with staged tiles - r02f - and was backed out; the CTA-level
        // tables and the static schedule's SN data are staged by TMA in cta_setup.)
        float4* so = reinterpret_cast<float4*>(smem + P.oObs) + slot * Dd.N_obs;
        const float4* go = Dd.obs4 + (size_t)sn * Dd.N_obs;
        for (int i = lane; i < Dd.N_obs; i += 32) so[i] = go[i];
        int4* ss = reinterpret_cast<int4*>(smem + P.oSup) + slot * Dd.n_b;
        const int4* gs = Dd.sup + (size_t)sn * Dd.n_b;
        for (int i = lane; i < Dd.n_b; i += 32) ss[i] = gs[i];
        if (STAGE) {
            float4* st = reinterpret_cast<float4*>(smem + P.oWt + slot * Dd.wt_stride);
            const float4* gt = reinterpret_cast<const float4*>(Dd.wt + (size_t)sn * Dd.wt_stride);
            for (int i = lane; i < Dd.wt_stride / 4; i += 32) st[i] = gt[i];
        }
    }
    __syncwarp();
    if (M.kw > 1) team_sync(slot, M.kw);
    c.wt_g = Dd.wt + (size_t)sn * Dd.wt_stride;
    c.nobs = Dd.n_valid[sn];
    c.muhat = Dd.muhat[sn];
    c.lconst = Dd.lconst[sn];
}

}  // namespace bayesn
