// bayesn_train.cuh - population training: log joint density of train_model_globalRV
// (reference bayesn/bayesn_model.py:1115-1182) and its gradient in the unconstrained space.
//
// One evaluation is four kernels:
//   train_globals_kernel   one CTA per chain: unconstrained global vector -> W0, W1, L_Sigma =
//                          diag(tan) * CorrCholesky(y), sigma0, R_V (+ dust basis), tau_A, and the log
//                          prior + log|det J| of the global sites
//   train_sn_kernel        one warp per (SN, chain): the fit evaluator in TRAIN mode - latent
//                          gradient + this SN's contributions to the global gradient
//   train_reduce_kernel    sums the contributions over the SNe of this rank (double), incl. the
//                          outer-product sum dL_Sigma = sum_s dW_s eps_tform_s^T
//   [NCCL all-reduce of the [C][R] double buffer when SNe are sharded over ranks - host side]
//   train_finish_kernel    one CTA per chain: chain rule through tan / stick breaking, priors
#pragma once
#include "bayesn_device.cuh"

namespace bayesn {

struct TrainDev {
    int C, G, Dl, n_corr, R, PR, LPTP;
    // per chain, written by train_globals_kernel
    float* W0;      // [C][LP*TP]
    float* W1;      // [C][LP*TP]
    float* Ls;      // [C][ne*ne]
    float* LsT;     // [C][ne*ne]
    float* ad;      // [C][JROWS]
    float* dad;     // [C][JROWS]
    float* sc;      // [C][8]  sigma0, RV, tauA
    float* sig;     // [C][ne] sigma_epsilon
    float* Lo;      // [C][ne*ne] L_Omega
    float* tm;      // [C][ne*ne] tanh(y) laid out as a strict lower triangle
    float* cum;     // [C][ne*ne] cumulative products c_ij = prod_{k<=j} (1 - t_ik^2)
    float* clog;    // [C][ne*ne] their logarithms
    float* scr;     // [C][ne*ne] scratch of the reverse sweep
    double* lpg;    // [C] log prior + log|det J| of the global sites
    float* part;    // [N][C][PR] per-SN contributions
    double* rpart;  // [ceil(N/32)][C][R] chunk partial sums of the reduction
};

// Per-SN rows and the reduced buffer end with TRAIN_NX sampler partial sums (used by the NUTS
// kernels below, zero otherwise): [kinetic energy, kinetic energy of a fresh momentum draw,
// (left, right) U-turn dot products for up to 12 checkpoints, (left, right) for the whole tree].
constexpr int TRAIN_NX = 2 + 2 * 12 + 2;
__host__ __device__ inline int train_part_base(int LPTP) { return LPTP + 3 + ((LPTP + 3) & 1) + 2; }
__host__ __device__ inline int train_part_stride(int LPTP) { return train_part_base(LPTP) + TRAIN_NX; }
__host__ __device__ inline int train_red_base(int LPTP, int ne) { return 4 + 2 * LPTP + ne * ne; }
__host__ __device__ inline int train_red_len(int LPTP, int ne) { return train_red_base(LPTP, ne) + TRAIN_NX; }

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;
}

// In-place running sum of row[0..n) by one warp (reverse = suffix sums): coalesced loads and a
// shuffle scan instead of one thread walking the row through global memory.
__device__ __forceinline__ void warp_row_cumsum(float* row, int n, int lane, bool reverse) {
    float carry = 0.f;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const int idx = reverse ? n - 1 - j : j;
        float v = (j < n) ? row[idx] : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(FULL, v, o);
            if (lane >= o) v += t;
        }
        v += carry;
        if (j < n) row[idx] = v;
        carry = __shfl_sync(FULL, v, 31);
    }
}

__device__ __forceinline__ double log_sigmoid_pair(double u) {   // log s(u) + log s(-u)
    const double a = fabs(u);
    return -a - 2.0 * log1p(exp(-a));
}

// ------------------------------------------------------------------------------------------
// (T1) globals: transforms + priors.  Layout of qg: [W0 vec_F, W1 vec_F, u_sigmaepsilon, y_L_Omega
//      (strict lower triangle, row-major), u_sigma0, u_RV, u_tauA]; interval supports use the
//      affine-sigmoid bijection, L_Omega numpyro's CorrCholeskyTransform (tanh + signed stick
//      breaking) with the LKJCholesky(eta = 1) density (upstream numpyro; restated in oracle/).
// ------------------------------------------------------------------------------------------
struct TrainShared {
    double sh[8];
    float sig[BAYESN_MAX_NEPS];
    float yk[9], dyk[9];
};

// `mat`: n_eps^2 floats of shared memory (scratch of the row scans)
__device__ __forceinline__ void train_globals_dev(const ModelDev& M, const TrainDev& Tr, int LP, int TP, int c,
                                                  const float* __restrict__ qg, TrainShared& S, float* mat) {
    double* sh = S.sh;
    float* s_sig = S.sig;
    float* s_yk = S.yk;
    float* s_dyk = S.dyk;
    const int tid = threadIdx.x;
    const int L = M.L, T = M.T, ne = M.n_eps;
    const int oW1 = L * T, oSe = 2 * L * T, oY = oSe + ne, oSc = oY + Tr.n_corr;
    double lp = 0.0;
    const double HALF_PI = 1.5707963267948966;
    // W0, W1 ~ N(0, I), reshaped order='F' into the padded [LP][TP] layout of the evaluator
    float* W0 = Tr.W0 + (size_t)c * Tr.LPTP;
    float* W1 = Tr.W1 + (size_t)c * Tr.LPTP;
    for (int p = tid; p < LP * TP; p += blockDim.x) {
        const int l = p / TP, m = p - l * TP;
        float a = 0.f, b = 0.f;
        if (l < L && m < T) {
            a = qg[l + L * m];
            b = qg[oW1 + l + L * m];
            lp += -0.5 * ((double)a * a + (double)b * b) - 1.8378770664093453;   // 2 * (1/2) log 2 pi
        }
        W0[p] = a;
        W1[p] = b;
    }
    // sigma_epsilon = tan(pi/2 sigmoid(u))
    for (int i = tid; i < ne; i += blockDim.x) {
        const double u = qg[oSe + i];
        const double sgm = 1.0 / (1.0 + exp(-u));
        const float se = (float)tan(HALF_PI * sgm);
        s_sig[i] = se;
        Tr.sig[(size_t)c * ne + i] = se;
        lp += log_sigmoid_pair(u);
    }
    __syncthreads();
    // rows of L_Omega by signed stick breaking; L_Sigma = diag(sigma_epsilon) L_Omega.  Row i:
    // t_j = tanh(y_ij), c_j = prod_{k<=j} (1 - t_k^2), L[i][j] = t_j sqrt(c_{j-1}), L[i][i] = sqrt(c_{i-1}).
    // Three sweeps: element-parallel transcendental work, a per-row running sum of log(1 - t^2) in
    // shared memory, element-parallel assembly - instead of one thread walking a row through tanh /
    // log / sqrt and global memory.
    float* Ls = Tr.Ls + (size_t)c * ne * ne;
    float* LsT = Tr.LsT + (size_t)c * ne * ne;
    float* Lo = Tr.Lo + (size_t)c * ne * ne;
    float* tm = Tr.tm + (size_t)c * ne * ne;
    float* cum = Tr.cum + (size_t)c * ne * ne;
    for (int e = tid; e < Tr.n_corr; e += blockDim.x) {
        int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
        while (i * (i - 1) / 2 > e) --i;
        while ((i + 1) * i / 2 <= e) ++i;
        const int j = e - i * (i - 1) / 2;
        const float yv = qg[oY + e], ay = fabsf(yv);
        // log(1 - tanh^2 y) = 2 (ln 2 - |y| - log1p(e^{-2|y|})); single precision suffices: t and the
        // cumulative products are consumed in float32, and the 1e-7 relative error of each term is far
        // below the 1e-5 log p tolerance after summing the n(n-1)/2 terms in double
        const float l1m = 2.f * (0.6931471805599453f - ay - log1pf(__expf(-2.f * ay)));
        tm[i * ne + j] = tanhf(yv);
        mat[i * ne + j] = l1m;
        lp += (double)l1m;                           // tanh log-det
    }
    __syncthreads();
    for (int i = tid; i < ne; i += blockDim.x) {
        float run = 0.f;
        for (int j = 0; j < i; ++j) {
            run += mat[i * ne + j];
            mat[i * ne + j] = run;
        }
    }
    __syncthreads();
    for (int e = tid; e < ne * ne; e += blockDim.x) {
        const int i = e / ne, j = e - i * ne;
        float lo = 0.f, t = 0.f, cj = 1.f;
        if (j <= i) {
            const float clp = (j > 0) ? mat[i * ne + j - 1] : 0.f;        // log c_{j-1}
            if (j < i) {
                t = tm[e];
                const float cl = mat[e];
                cj = expf(cl);
                lo = t * expf(0.5f * clp);
                // stick-breaking log-det (columns 0..i-2) + LKJCholesky(eta = 1) on the diagonal element
                lp += (j <= i - 2) ? 0.5 * (double)cl : 0.5 * (double)(ne - 1 - i) * (double)cl;
            } else {
                lo = expf(0.5f * clp);
            }
        }
        const float se = s_sig[i];
        Lo[e] = lo;
        if (j >= i) tm[e] = 0.f;
        cum[e] = cj;
        Ls[e] = se * lo;
        LsT[j * ne + i] = se * lo;
    }
    // sigma0 = 0.1 tan(.), RV = 1 + 4 sigmoid(.), tauA = tan(.)
    if (tid == 0) {
        const double u0 = qg[oSc], u1 = qg[oSc + 1], u2 = qg[oSc + 2];
        const double s0 = 1.0 / (1.0 + exp(-u0)), s1 = 1.0 / (1.0 + exp(-u1)), s2 = 1.0 / (1.0 + exp(-u2));
        float* sc = Tr.sc + (size_t)c * 8;
        sc[0] = (float)(0.1 * tan(HALF_PI * s0));
        sc[1] = (float)(1.0 + 4.0 * s1);
        sc[2] = (float)tan(HALF_PI * s2);
        lp += log_sigmoid_pair(u0) + log_sigmoid_pair(u1) + log_sigmoid_pair(u2);
        // Fitzpatrick99 knot values and their R_V derivatives (reference :604-623)
        const float R = sc[1], R2 = R * R, R3 = R2 * R;
        const float c2f = -0.824f + 4.717f / R, c1f = 2.030f - 3.007f * c2f;
        const float dc2 = -4.717f / R2, dc1 = -3.007f * dc2;
        const float x7 = 1e4f / 2700.f, x8 = 1e4f / 2600.f, x0 = 4.596f, gm = 0.99f;
        const float d1 = x7 * x7 / ((x7 * x7 - x0 * x0) * (x7 * x7 - x0 * x0) + gm * gm * x7 * x7);
        const float d2 = x8 * x8 / ((x8 * x8 - x0 * x0) * (x8 * x8 - x0 * x0) + gm * gm * x8 * x8);
        s_yk[0] = -R;                                   s_dyk[0] = -1.f;
        s_yk[1] = 0.26469f * R / 3.1f - R;              s_dyk[1] = 0.26469f / 3.1f - 1.f;
        s_yk[2] = 0.82925f * R / 3.1f - R;              s_dyk[2] = 0.82925f / 3.1f - 1.f;
        s_yk[3] = -0.422809f + 0.00270f * R + 2.13572e-4f * R2;     s_dyk[3] = 0.00270f + 2.f * 2.13572e-4f * R;
        s_yk[4] = -5.13540e-2f + 0.00216f * R - 7.35778e-5f * R2;   s_dyk[4] = 0.00216f - 2.f * 7.35778e-5f * R;
        s_yk[5] = 0.700127f + 0.00184f * R - 3.32598e-5f * R2;      s_dyk[5] = 0.00184f - 2.f * 3.32598e-5f * R;
        s_yk[6] = 1.19456f + 0.01707f * R - 5.46959e-3f * R2 + 7.97809e-4f * R3 - 4.45636e-5f * R2 * R2;
        s_dyk[6] = 0.01707f - 2.f * 5.46959e-3f * R + 3.f * 7.97809e-4f * R2 - 4.f * 4.45636e-5f * R3;
        s_yk[7] = c1f + c2f * x7 + 3.23f * d1;          s_dyk[7] = dc1 + dc2 * x7;
        s_yk[8] = c1f + c2f * x8 + 3.23f * d2;          s_dyk[8] = dc1 + dc2 * x8;
    }
    __syncthreads();
    {
        const float R = Tr.sc[(size_t)c * 8 + 1], iR = 1.f / R;
        float* ad = Tr.ad + (size_t)c * JROWS;
        float* dad = Tr.dad + (size_t)c * JROWS;
        for (int b = tid; b < JROWS; b += blockDim.x) {
            float my = 0.f, dmy = 0.f;
            if (b < NB) {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const float mf = M.Mf[j * BP + b];
                    my = fmaf(mf, s_yk[j], my);
                    dmy = fmaf(mf, s_dyk[j], dmy);
                }
            }
            ad[b] = (b < NB) ? 1.f + my * iR : 0.f;
            dad[b] = (b < NB) ? (dmy - my * iR) * iR : 0.f;
        }
    }
    const double tot = block_sum(lp, sh);
    if (tid == 0) Tr.lpg[c] = tot;
}

__global__ void __launch_bounds__(128) train_globals_kernel(ModelDev M, TrainDev Tr, int LP, int TP,
                                                            const float* __restrict__ qg_all) {
    __shared__ TrainShared S;
    extern __shared__ __align__(16) float tmat[];
    train_globals_dev(M, Tr, LP, TP, blockIdx.x, qg_all + (size_t)blockIdx.x * Tr.G, S, tmat);
}

// ------------------------------------------------------------------------------------------
// (T2) per-SN part: the fit evaluator in TRAIN mode
// ------------------------------------------------------------------------------------------
template <int LP, int TP, int VPL, bool STAGE>
__global__ void __launch_bounds__(WPB * 32) train_sn_kernel(ModelDev M, DataDev Dd, TrainDev Tr, int ntile,
                                                            const float* __restrict__ ql, float* __restrict__ grad_l) {
    extern __shared__ __align__(16) float smem[];
    WarpCtx c;
    int sn, chain;
    const int C = Tr.C;
    const bool active = cta_setup<LP, TP, false, STAGE, true>(M, Dd, C, ntile, smem, c, sn, chain);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    const size_t unit = (size_t)sn * C + chain;
    const int ne = M.n_eps, Dl = Tr.Dl;
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
    c.part = Tr.part + unit * Tr.PR;
    float q[VPL], g[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        const int i = lane + 32 * s;
        q[s] = (i < Dl) ? ql[unit * Dl + i] : 0.f;
        if (i == ne + 2) q[s] -= c.muhat;
    }
    double lp;
    eval_logp_grad<LP, TP, VPL, false, STAGE, true>(M, c, M.hs, q, lp, g, lane);
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
        const int i = lane + 32 * s;
        if (i < Dl) grad_l[unit * Dl + i] = g[s];
    }
}

// ------------------------------------------------------------------------------------------
// (T3) reduction over the SNe of this rank.  red[c] = [log p, d/dsigma0, d/dRV, d/dtauA,
//      dW0 (padded [LP][TP]), dW1, dL_Sigma (row-major, lower triangle), sampler partial sums],
//      all double.  dL_Sigma = sum_s vec(dW_s) eps_tform_s^T is a small GEMM (n_eps x N x n_eps):
//      stage 1 tiles it - one CTA per (chunk of 32 SNe, chain) stages the chunk's rows and latents
//      in shared memory once and produces every output's partial sum; stage 2 adds the chunk
//      partials in a fixed order (deterministic: ranks and replays must agree bit for bit).
//      The latents (theta_s, eps_tform_s) of SN s / chain c are read at lat + c * chain_stride +
//      s * sn_stride (the caller's ql array, or the sampler's current-position vectors).
// ------------------------------------------------------------------------------------------
constexpr int TRAIN_RCHUNK = 32;

__global__ void __launch_bounds__(256) train_reduce_kernel(ModelDev M, TrainDev Tr, int N, int LP, int TP,
                                                           const float* __restrict__ lat, size_t chain_stride,
                                                           size_t sn_stride, double* __restrict__ rpart) {
    extern __shared__ __align__(16) float rsm[];
    const int chunk = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int ne = M.n_eps, L = M.L, C = Tr.C, PR = Tr.PR, LPTP = Tr.LPTP, R = Tr.R;
    const int s0 = chunk * TRAIN_RCHUNK;
    const int ns = min(TRAIN_RCHUNK, N - s0);
    const int ES = ne + 2;                                  // eps_tform, theta (+ pad)
    float* A = rsm;                                         // [32][PR]   per-SN rows
    float* E = rsm + TRAIN_RCHUNK * PR;                     // [32][ES]   latents
    for (int e = tid; e < ns * PR; e += blockDim.x) {
        const int s = e / PR, x = e - s * PR;
        A[e] = Tr.part[((size_t)(s0 + s) * C + c) * PR + x];
    }
    for (int e = tid; e < ns * (ne + 1); e += blockDim.x) {
        const int s = e / (ne + 1), x = e - s * (ne + 1);
        E[s * ES + x] = lat[(size_t)c * chain_stride + (size_t)(s0 + s) * sn_stride + x];
    }
    __syncthreads();
    const int off_lp = LPTP + 3 + ((LPTP + 3) & 1);
    const int rbase = train_red_base(LPTP, ne), pbase = train_part_base(LPTP);
    double* out = rpart + ((size_t)chunk * C + c) * R;
    // blockIdx.z splits the outputs: one output per thread, 32 terms each
    for (int r = blockIdx.z * blockDim.x + tid; r < min(R, (int)(blockIdx.z + 1) * (int)blockDim.x); r += blockDim.x) {
        double acc = 0.0;
        if (r == 0) {
            for (int s = 0; s < ns; ++s) acc += *reinterpret_cast<const double*>(A + s * PR + off_lp);
        } else if (r < 4) {
            for (int s = 0; s < ns; ++s) acc += A[s * PR + LPTP + (r - 1)];
        } else if (r < 4 + LPTP) {
            for (int s = 0; s < ns; ++s) acc += A[s * PR + (r - 4)];
        } else if (r < 4 + 2 * LPTP) {
            const int p = r - 4 - LPTP;
            for (int s = 0; s < ns; ++s) acc += (double)A[s * PR + p] * E[s * ES + ne];
        } else if (r >= rbase) {
            const int x = pbase + (r - rbase);
            for (int s = 0; s < ns; ++s) acc += A[s * PR + x];
        } else {
            const int e = r - 4 - 2 * LPTP;
            const int i = e / ne, j = e - i * ne;
            if (j <= i) {
                // vec_F index i = (l-1) + (L-2) m  ->  position of dW[l][m] in the padded layout
                const int m = i / (L - 2), l = 1 + (i - m * (L - 2));
                const int p = l * TP + m;
                for (int s = 0; s < ns; ++s) acc += (double)A[s * PR + p] * E[s * ES + j];
            }
        }
        out[r] = acc;
    }
}

// Exchange of the reduced sums between the ranks that hold different SN shards, fused into the
// second reduction stage: every rank STORES its sums straight into a slot of every peer's buffer
// over NVLink (peer-mapped symmetric memory), the last CTA to finish publishes a sequence number
// to every peer, and the consumer (the control kernel / train_finish) waits for all ranks'
// numbers and adds the slots in rank order - the same bits on every rank, no NCCL launch, one
// NVLink store latency instead of a collective.  Buffers are double-buffered by pass parity.
struct XchgDev {
    int world, rank;
    double* peer_buf[8];          // peer p's buffer: [2][world][C][R] doubles
    unsigned* peer_flag[8];       // peer p's flags:  [2][world]
    double* my_buf;
    unsigned* my_flag;
    unsigned* seq;                // this rank's pass counter (device)
    unsigned* done;               // CTA completion counter of the publishing kernel
    unsigned* err;                // set if a wait timed out
};

__global__ void __launch_bounds__(256) train_reduce_final_kernel(TrainDev Tr, int nchunk, const double* __restrict__ rpart,
                                                                 double* __restrict__ red, XchgDev X) {
    const int c = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned next = X.world > 1 ? *X.seq + 1u : 0u;
    if (r < Tr.R) {
        double acc = 0.0;
        for (int k = 0; k < nchunk; ++k) acc += rpart[((size_t)k * Tr.C + c) * Tr.R + r];
        if (X.world > 1) {
            const size_t slot = (((size_t)(next & 1u) * X.world + X.rank) * Tr.C + c) * Tr.R + r;
            for (int p = 0; p < X.world; ++p) X.peer_buf[p][slot] = acc;
        } else {
            red[(size_t)c * Tr.R + r] = acc;
        }
    }
    if (X.world > 1) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned total = gridDim.x * gridDim.y;
            if (atomicAdd(X.done, 1u) == total - 1) {
                *X.done = 0u;
                __threadfence_system();
                for (int p = 0; p < X.world; ++p) {
                    unsigned* f = X.peer_flag[p] + (next & 1u) * X.world + X.rank;
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(next) : "memory");
                }
                *X.seq = next;
            }
        }
    }
}

// consumer side: wait for every rank's sums of the current pass, add them in rank order into red[c]
// Returns false when a rank's sums did not arrive in time (or an earlier pass already failed): the caller must
// stop the chain - adding stale slots would let the ranks' replicated sampler states diverge.
__device__ __forceinline__ bool xchg_gather(const XchgDev& X, const TrainDev& Tr, int c, double* __restrict__ red) {
    if (X.world <= 1) return true;
    __shared__ int xchg_bad;
    if (threadIdx.x == 0) xchg_bad = (*reinterpret_cast<volatile unsigned*>(X.err) != 0u);
    __syncthreads();
    const unsigned cur = *X.seq;
    if ((int)threadIdx.x < X.world && !xchg_bad) {
        const unsigned* f = X.my_flag + (cur & 1u) * X.world + threadIdx.x;
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (clock64() - t0 > 4000000000LL) { *X.err = 1u; xchg_bad = 1; break; }      // ~2 s: never hang the GPU
        } while (v != cur);
    }
    __syncthreads();
    if (xchg_bad) return false;
    const double* src = X.my_buf + ((size_t)(cur & 1u) * X.world * Tr.C + c) * Tr.R;
    for (int r = threadIdx.x; r < Tr.R; r += blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < X.world; ++p) acc += __ldcg(src + (size_t)p * Tr.C * Tr.R + r);   // L2: peers wrote it
        red[(size_t)c * Tr.R + r] = acc;
    }
    __syncthreads();
    return true;
}

// ------------------------------------------------------------------------------------------
// (T4) finish: log p = reduced per-SN terms + global priors; chain rule to the unconstrained globals
// ------------------------------------------------------------------------------------------
// gg = sign * d log p / d qg (sign = -1 gives the gradient of the potential energy)
__device__ __forceinline__ void train_finish_dev(const ModelDev& M, const TrainDev& Tr, int LP, int TP, int c,
                                                 const float* __restrict__ qg, const double* __restrict__ red,
                                                 float* __restrict__ gg, float sign, float* mat) {
    const int tid = threadIdx.x;
    const int L = M.L, T = M.T, ne = M.n_eps, LPTP = Tr.LPTP;
    const int oW1 = L * T, oSe = 2 * L * T, oY = oSe + ne, oSc = oY + Tr.n_corr;
    const double HALF_PI = 1.5707963267948966;
    for (int p = tid; p < L * T; p += blockDim.x) {
        const int m = p / L, l = p - m * L;          // vec_F index p = l + L m
        gg[p] = sign * (float)(red[4 + l * TP + m] - (double)qg[p]);
        gg[oW1 + p] = sign * (float)(red[4 + LPTP + l * TP + m] - (double)qg[oW1 + p]);
    }
    const double* Gm = red + 4 + 2 * LPTP;            // dL_Sigma [ne][ne]
    const float* Lo = Tr.Lo + (size_t)c * ne * ne;
    const float* tm = Tr.tm + (size_t)c * ne * ne;
    const float* cum = Tr.cum + (size_t)c * ne * ne;
    float* scr = mat;                                  // shared memory
    // Reverse sweep of the stick breaking.  With gL_ij = sigma_i G[i][j] (d/dL_Omega[i][j]) and
    // A_j the weight of log c_j in the density, d/dc_j obeys dc_j = a_j / c_j + (1 - t_{j+1}^2) dc_{j+1},
    // a_j = A_j + sqrt(c_j)/2 * (j == i-1 ? gL_ii : gL_{i,j+1} t_{j+1}); multiplying by c_j turns it
    // into a plain suffix sum S_j = sum_{k >= j} a_k, and
    //   d/dy_ij = gL_ij sqrt(c_{j-1}) (1 - t_j^2) - 2 t_j (S_j + 1).
    for (int e = tid; e < ne * ne; e += blockDim.x) {
        const int i = e / ne, j = e - i * ne;
        float a = 0.f;
        if (j < i) {
            const float se = Tr.sig[(size_t)c * ne + i];
            const float A = (j <= i - 2) ? 0.5f : 0.5f * (float)(ne - 1 - i);
            const float nxt = (j == i - 1) ? (float)Gm[i * ne + i] : (float)Gm[e + 1] * tm[e + 1];
            a = A + 0.5f * sqrtf(cum[e]) * se * nxt;
        }
        scr[e] = a;
    }
    __syncthreads();
    for (int i = tid; i < ne; i += blockDim.x) {
        // d/dsigma_i = sum_j G[i][j] L_Omega[i][j], and the suffix sums of row i (shared memory)
        double dse = 0.0;
        for (int j = 0; j <= i; ++j) dse += Gm[i * ne + j] * (double)Lo[i * ne + j];
        float run = 0.f;
        for (int j = i - 1; j >= 0; --j) {
            run += scr[i * ne + j];
            scr[i * ne + j] = run;
        }
        const double se = Tr.sig[(size_t)c * ne + i];
        const double u = qg[oSe + i];
        const double sgm = 1.0 / (1.0 + exp(-u));
        gg[oSe + i] = sign * (float)(dse * (1.0 + se * se) * HALF_PI * sgm * (1.0 - sgm) + (1.0 - 2.0 * sgm));
    }
    __syncthreads();
    for (int e = tid; e < Tr.n_corr; e += blockDim.x) {
        int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
        while (i * (i - 1) / 2 > e) --i;
        while ((i + 1) * i / 2 <= e) ++i;
        const int j = e - i * (i - 1) / 2;
        const int m = i * ne + j;
        const float t = tm[m];
        const float cjm = (j > 0) ? cum[m - 1] : 1.f;
        const float se = Tr.sig[(size_t)c * ne + i];
        gg[oY + e] = sign * (se * (float)Gm[m] * sqrtf(cjm) * (1.f - t * t) - 2.f * t * (scr[m] + 1.f));
    }
    if (tid == 0) {
        const float* sc = Tr.sc + (size_t)c * 8;
        const double u0 = qg[oSc], u1 = qg[oSc + 1], u2 = qg[oSc + 2];
        const double s0 = 1.0 / (1.0 + exp(-u0)), s1 = 1.0 / (1.0 + exp(-u1)), s2 = 1.0 / (1.0 + exp(-u2));
        const double sig0 = sc[0], tauA = sc[2];
        gg[oSc] = sign * (float)(red[1] * 0.1 * (1.0 + 100.0 * sig0 * sig0) * HALF_PI * s0 * (1.0 - s0) + (1.0 - 2.0 * s0));
        gg[oSc + 1] = sign * (float)(red[2] * 4.0 * s1 * (1.0 - s1) + (1.0 - 2.0 * s1));
        gg[oSc + 2] = sign * (float)(red[3] * (1.0 + tauA * tauA) * HALF_PI * s2 * (1.0 - s2) + (1.0 - 2.0 * s2));
    }
}

__global__ void __launch_bounds__(128) train_finish_kernel(ModelDev M, TrainDev Tr, int LP, int TP,
                                                           const float* __restrict__ qg_all, double* __restrict__ red_all,
                                                           double* __restrict__ logp, float* __restrict__ grad_g, XchgDev X) {
    const int c = blockIdx.x;
    if (!xchg_gather(X, Tr, c, red_all)) {
        if (threadIdx.x == 0) logp[c] = nan("");          // the exchange failed: bayesn_train_finish reports it
        return;
    }
    const double* red = red_all + (size_t)c * Tr.R;
    extern __shared__ __align__(16) float tmat[];
    if (threadIdx.x == 0) logp[c] = red[0] + Tr.lpg[c];
    train_finish_dev(M, Tr, LP, TP, c, qg_all + (size_t)c * Tr.G, red, grad_g + (size_t)c * Tr.G, 1.f, tmat);
}

}  // namespace bayesn
