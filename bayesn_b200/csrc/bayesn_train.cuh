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
    float* cum;     // [C][ne*ne] cumulative products prod_{k<=j} (1 - t_ik^2)
    double* lpg;    // [C] log prior + log|det J| of the global sites
    float* part;    // [N][C][PR] per-SN contributions
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

__device__ __forceinline__ void train_globals_dev(const ModelDev& M, const TrainDev& Tr, int LP, int TP, int c,
                                                  const float* __restrict__ qg, TrainShared& S) {
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
    // rows of L_Omega by signed stick breaking; L_Sigma = diag(sigma_epsilon) L_Omega
    float* Ls = Tr.Ls + (size_t)c * ne * ne;
    float* LsT = Tr.LsT + (size_t)c * ne * ne;
    float* Lo = Tr.Lo + (size_t)c * ne * ne;
    float* tm = Tr.tm + (size_t)c * ne * ne;
    float* cum = Tr.cum + (size_t)c * ne * ne;
    for (int i = tid; i < ne; i += blockDim.x) {
        const float* y = qg + oY + (size_t)i * (i - 1) / 2;
        double cprod = 1.0;
        const float se = s_sig[i];
        for (int j = 0; j < ne; ++j) {
            float lo = 0.f, t = 0.f;
            if (j < i) {
                const double yv = y[j];
                const double td = tanh(yv);
                t = (float)td;
                lo = (float)(td * sqrt(cprod));
                const double om = 1.0 - td * td;
                cprod *= om;
                // tanh log-det + stick-breaking log-det (columns 0..i-2) + LKJ (diagonal element)
                lp += 2.0 * (0.6931471805599453 - fabs(yv) - log1p(exp(-2.0 * fabs(yv))));
                lp += (j <= i - 2) ? 0.5 * log(cprod) : 0.5 * (double)(ne - 1 - i) * log(cprod);
            } else if (j == i) {
                lo = (float)sqrt(cprod);
            }
            Lo[i * ne + j] = lo;
            tm[i * ne + j] = t;
            cum[i * ne + j] = (float)cprod;
            Ls[i * ne + j] = se * lo;
            LsT[j * ne + i] = se * lo;
        }
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
    train_globals_dev(M, Tr, LP, TP, blockIdx.x, qg_all + (size_t)blockIdx.x * Tr.G, S);
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
    const bool active = cta_setup<LP, TP, false, STAGE>(M, Dd, C, ntile, smem, c, sn, chain);
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
//      dW0 (padded [LP][TP]), dW1, dL_Sigma (row-major, lower triangle)], all double.
// ------------------------------------------------------------------------------------------
//      The latents (theta_s, eps_tform_s) of SN s / chain c are read at lat + c * chain_stride +
//      s * sn_stride (the caller's ql array, or the sampler's current-position vectors).
__global__ void __launch_bounds__(256) train_reduce_kernel(ModelDev M, TrainDev Tr, int N, int LP, int TP,
                                                           const float* __restrict__ lat, size_t chain_stride,
                                                           size_t sn_stride, double* __restrict__ red) {
    const int c = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int ne = M.n_eps, L = M.L, C = Tr.C, PR = Tr.PR, LPTP = Tr.LPTP;
    if (r >= Tr.R) return;
    const size_t stride = (size_t)C * PR;
    const float* base = Tr.part + (size_t)c * PR;
    double acc = 0.0;
    if (r == 0) {
        const int off = LPTP + 3 + ((LPTP + 3) & 1);
        for (int s = 0; s < N; ++s) acc += *reinterpret_cast<const double*>(base + s * stride + off);
    } else if (r < 4) {
        for (int s = 0; s < N; ++s) acc += base[s * stride + LPTP + (r - 1)];
    } else if (r < 4 + LPTP) {
        const int p = r - 4;
        for (int s = 0; s < N; ++s) acc += base[s * stride + p];
    } else if (r < 4 + 2 * LPTP) {
        const int p = r - 4 - LPTP;
        const float* th = lat + (size_t)c * chain_stride + ne;       // theta_s
        for (int s = 0; s < N; ++s) acc += (double)base[s * stride + p] * th[(size_t)s * sn_stride];
    } else if (r >= train_red_base(LPTP, ne)) {
        const int x = train_part_base(LPTP) + (r - train_red_base(LPTP, ne));
        for (int s = 0; s < N; ++s) acc += base[s * stride + x];
    } else {
        const int e = r - 4 - 2 * LPTP;
        const int i = e / ne, j = e - i * ne;
        if (j <= i) {
            // vec_F index i = (l-1) + (L-2) m  ->  position of dW[l][m] in the padded layout
            const int m = i / (L - 2), l = 1 + (i - m * (L - 2));
            const int p = l * TP + m;
            const float* ev = lat + (size_t)c * chain_stride + j;    // eps_tform_s[j]
            for (int s = 0; s < N; ++s) acc += (double)base[s * stride + p] * ev[(size_t)s * sn_stride];
        }
    }
    red[(size_t)c * Tr.R + r] = acc;
}

// ------------------------------------------------------------------------------------------
// (T4) finish: log p = reduced per-SN terms + global priors; chain rule to the unconstrained globals
// ------------------------------------------------------------------------------------------
// gg = sign * d log p / d qg (sign = -1 gives the gradient of the potential energy)
__device__ __forceinline__ void train_finish_dev(const ModelDev& M, const TrainDev& Tr, int LP, int TP, int c,
                                                 const float* __restrict__ qg, const double* __restrict__ red,
                                                 float* __restrict__ gg, float sign) {
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
    for (int i = tid; i < ne; i += blockDim.x) {
        const double se = Tr.sig[(size_t)c * ne + i];
        // d/dsigma_i = sum_j G[i][j] L_Omega[i][j];  d/dL_Omega[i][j] = sigma_i G[i][j]
        double dse = 0.0;
        for (int j = 0; j <= i; ++j) dse += Gm[i * ne + j] * (double)Lo[i * ne + j];
        {
            const double u = qg[oSe + i];
            const double sgm = 1.0 / (1.0 + exp(-u));
            gg[oSe + i] = sign * (float)(dse * (1.0 + se * se) * HALF_PI * sgm * (1.0 - sgm) + (1.0 - 2.0 * sgm));
        }
        if (i == 0) continue;
        // reverse sweep of the stick breaking of row i: c_j = c_{j-1} (1 - t_j^2), L[i][j] = t_j sqrt(c_{j-1}),
        // L[i][i] = sqrt(c_{i-1}); log-density terms A_j log c_j + log(1 - t_j^2)
        float* gy = gg + oY + (size_t)i * (i - 1) / 2;
        double dc = 0.0;                               // d/dc_j flowing down from columns > j
        for (int j = i - 1; j >= 0; --j) {
            const double cj = cum[i * ne + j];
            const double cjm = (j > 0) ? (double)cum[i * ne + j - 1] : 1.0;
            const double t = tm[i * ne + j];
            // contributions to d/dc_j
            const double A = (j <= i - 2) ? 0.5 : 0.5 * (double)(ne - 1 - i);
            double dcj = dc + A / cj;
            if (j == i - 1) dcj += se * Gm[i * ne + i] * 0.5 / sqrt(cj);
            else dcj += se * Gm[i * ne + j + 1] * (double)tm[i * ne + j + 1] * 0.5 / sqrt(cj);
            // t_j: direct through L[i][j], through c_j, and the tanh log-det
            const double om = 1.0 - t * t;
            const double dt = se * Gm[i * ne + j] * sqrt(cjm) + dcj * cjm * (-2.0 * t);
            gy[j] = sign * (float)(dt * om - 2.0 * t);
            dc = dcj * om;
        }
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
                                                           const float* __restrict__ qg_all,
                                                           const double* __restrict__ red_all,
                                                           double* __restrict__ logp, float* __restrict__ grad_g) {
    const int c = blockIdx.x;
    const double* red = red_all + (size_t)c * Tr.R;
    if (threadIdx.x == 0) logp[c] = red[0] + Tr.lpg[c];
    train_finish_dev(M, Tr, LP, TP, c, qg_all + (size_t)c * Tr.G, red, grad_g + (size_t)c * Tr.G, 1.f);
}

}  // namespace bayesn
