/*
 * bayesn_b200.h - C ABI of libbayesn_b200.so, the B200 (sm_100a) implementation of BayeSN's
 * data-parallel hot path: SED forward model + analytic gradient inside a device-resident,
 * batched NUTS sampler.
 *
 * The reference (bayesn v0.4.1) has no FFI layer: its seam is the Python method surface of
 * `SEDmodel`.  Each entry point below names the reference code it replaces (file:line under
 * /root/reference).  `bayesn_b200/native.py` binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; bayesn_last_error(ctx) gives the text.
 *  - "host" pointers are read during the call and copied; "device" pointers are CUDA device
 *    memory owned by the caller, must stay valid until the stream work that uses them is done;
 *    the library never frees them.  Dataset pointers are retained until the next
 *    bayesn_bind_dataset / bayesn_ctx_destroy.
 *  - all kernels are launched asynchronously on `stream` (a cudaStream_t passed as void*;
 *    NULL = legacy default stream).  A ctx is bound to one device; do not use one ctx from two
 *    host threads at once.
 *  - FP32 device data, int32 indices.  Wavelength-bin rows are padded to BAYESN_BPAD floats.
 */
#ifndef BAYESN_B200_H
#define BAYESN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAYESN_NBINS 300      /* spectrum_bins, bayesn/bayesn_model.py:295 */
#define BAYESN_BPAD 304       /* row stride of every per-bin table (16-byte aligned rows) */
#define BAYESN_OVERRUN 320    /* floats the kernels may read (and discard) past a band row's support */
#define BAYESN_NPHASE 105     /* Hsiao phase grid -19..85, bayesn/bayesn_model.py:274 */
#define BAYESN_MAX_L 12
#define BAYESN_MAX_T 10
#define BAYESN_MAX_NEPS ((BAYESN_MAX_L - 2) * BAYESN_MAX_T)

typedef struct bayesn_ctx bayesn_ctx;

/* RV treatment of the fit model: which numpyro model function the log-density follows. */
enum { BAYESN_RV_FIXED = 0,    /* fit_model_globalRV  bayesn/bayesn_model.py:883-945  */
       BAYESN_RV_POP = 1,      /* fit_model_popRV     bayesn/bayesn_model.py:996-1060 */
       BAYESN_RV_UNIFORM = 2   /* fit_model_uniformRV bayesn/bayesn_model.py:818-881  */ };

/* Static model tables (all HOST pointers, row-major, float32).  Replaces what
 * SEDmodel.__init__ / _setup_band_weights / _load_hsiao_template keep as attributes
 * (bayesn/bayesn_model.py:236-252, :483-497, :278-285). */
typedef struct {
    int32_t L, T;             /* len(l_knots), len(tau_knots) */
    int32_t rv_mode;          /* BAYESN_RV_* */
    const float* tau_knots;   /* [T] */
    const float* KD_t;        /* [T][T]   invKD_irr(tau_knots) */
    const float* J_l;         /* [300][L] spline_coeffs_irr(model_wave, l_knots) */
    const float* hsiao;       /* [300][105] Hsiao template on the model grid */
    const float* M_fitz;      /* [300][9] F99 basis: k(lambda-V)/E(B-V) = M_fitz yk(R_V).  Spline weights of the nine
                                 anchors (reference M_fitz_block, bayesn_model.py:313-316); for bins below 2700 A the
                                 row must be (0,..,0,alpha,beta) reproducing the FM90 ultraviolet override of
                                 get_spectra (:627-638) - bayesn_b200.static.fitz_matrix builds it */
    const float* a_dust;      /* [300] 1 + (M_fitz yk(RV))/RV for the global RV (RV_FIXED only) */
    const float* W0;          /* [L][T] */
    const float* W1;          /* [L][T] */
    const float* L_Sigma;     /* [n_eps][n_eps], n_eps = (L-2)*T */
    float tauA;               /* A_V ~ Exponential(1/tauA) */
    float Ds_prior_sd;        /* sqrt(5^2 + sigma0^2), bayesn/bayesn_model.py:937-940 */
    float RV, mu_R, sigma_R, trunc_R;
    int32_t fix_tmax, fix_theta, fix_AV;   /* fit() flags, bayesn/bayesn_model.py:914-918 */
    float theta_val, AV_val;
} bayesn_model_desc;

/* One data set of N supernovae (all DEVICE pointers).  This is the packed, float32 form of the
 * reference's `obs` (10, N_obs, N) array and `weights` (N, 300, n_b) array
 * (bayesn/bayesn_model.py:2096-2106, :2711-2730); bayesn_b200/model.py builds it.
 *   obs4[s][o]   = {phase t, FLUXCAL y, 1/sigma, band index as float}; valid observations first
 *   n_valid[s]   = number of unmasked observations of SN s (<= N_obs)
 *   weights[s][wt_stride] = compact weight tile of SN s: for every band k that has observations,
 *                  the non-zero extent [first_k, last_k) of
 *                  band weight * 10^(0.4(27.5-offset_k) - zp_k/2.5 - 0.4(M0+muhat_s)), rows packed back
 *                  to back at support[s][k].offset; wt_stride is a multiple of 4 and the allocation
 *                  must extend BAYESN_OVERRUN floats past the last tile (read, never used)
 *   support[s][k] = {first, one-past-last, offset, 0}: exact non-zero extent of the band weight
 *                  (threshold-free: exact zeros only) and where its row starts inside the tile
 *   muhat[s]     = fiducial distance modulus;  lconst[s] = data-only log-density constant
 *   muhat_err2[s] = (5 / (z ln 10))^2 (z_err^2 + sigma_pec^2), training only (NULL for fits;
 *                  bayesn/bayesn_model.py:1174-1175); training data carry magnitudes in obs4.y */
typedef struct {
    int32_t N, N_obs, n_b, wt_stride;
    const float* obs4;
    const int32_t* n_valid;
    const float* weights;
    const int32_t* support;
    const float* muhat;
    const float* lconst;
    const float* muhat_err2;
} bayesn_dataset_desc;

/* NUTS configuration: numpyro.infer.NUTS / MCMC arguments at the reference call sites
 * bayesn/bayesn_model.py:1798-1802, :1830-1833, :2110-2141. */
typedef struct {
    int32_t num_warmup, num_samples, num_chains, max_tree_depth;
    float target_accept, init_step_size;
    int32_t adapt_step_size, adapt_mass_matrix, regularize_mass_matrix;
    uint64_t seed;
    uint64_t unit_offset;   /* index of this data set's first (SN, chain) unit in the whole job (first SN *
                             * num_chains): random streams are keyed by the global unit, so sharding
                             * SNe over GPUs reproduces the single-GPU draws bit for bit */
    int32_t warps_per_chain; /* latency mode: 0 / 1 = one warp per (SN, chain) (throughput layout); 2, 4, 8 = that
                              * many warps share one chain's observation pairs and exchange their partial sums
                              * through shared memory once per evaluation.  For few chains on a whole GPU (a single
                              * fit(), small shards of a strong-scaled job).  Draws depend on the value (summation
                              * order) but not on the sharding. */
    int32_t reserved;
} bayesn_nuts_cfg;

int bayesn_ctx_create(int device, bayesn_ctx** out);
void bayesn_ctx_destroy(bayesn_ctx* ctx);
const char* bayesn_last_error(bayesn_ctx* ctx);
int bayesn_abi_version(void);

int bayesn_set_model(bayesn_ctx* ctx, const bayesn_model_desc* model);
int bayesn_bind_dataset(bayesn_ctx* ctx, const bayesn_dataset_desc* data);

/* Dimension of the unconstrained per-SN parameter vector: n_eps + 4 (+1 if RV is sampled).
 * Layout: [eps_tform[0..n_eps), theta, log A_V, logit((tmax+10)/20), Ds, (logit RV-transform)]. */
int bayesn_param_dim(bayesn_ctx* ctx);

/* Log joint density (unconstrained space, incl. log|det J|) and its gradient for every
 * (SN, chain): one call = what numpyro evaluates per leapfrog through
 * fit_model_globalRV / popRV / uniformRV + JAX reverse-mode AD.
 *   q    [N][n_chains][D]  device, in
 *   logp [N][n_chains]     device, out
 *   grad [N][n_chains][D]  device, out */
int bayesn_logp_grad_fit(bayesn_ctx* ctx, int n_chains, const float* q, float* logp, float* grad, void* stream);
/* The same evaluation in latency mode (warps_per_chain as in bayesn_nuts_cfg). */
int bayesn_logp_grad_fit_team(bayesn_ctx* ctx, int n_chains, int warps_per_chain, const float* q, float* logp,
                              float* grad, void* stream);

/* Forward model only: SEDmodel.get_flux_batch (bayesn/bayesn_model.py:645-709) when mag == 0,
 * SEDmodel.get_mag_batch (:711-759) when mag != 0.  Same argument meaning and array layouts as
 * the reference (device float32 / int32):
 *   theta, AV, Ds, RV [N]; eps [N][L][T]; W0, W1 [L][T]; band_indices [N_obs][N];
 *   mask [N_obs][N] (uint8); J_t [N][T][N_obs]; hsiao_interp [3][N_obs][N];
 *   weights [N][300][n_b]; band_log10_scale [n_b] (HOST, double) =
 *   0.4*(27.5 - offset_k) - zp_k/2.5; out [N_obs][N]. */
int bayesn_flux_batch(bayesn_ctx* ctx, int N, int N_obs, int n_b, int mag, float M0, const float* theta,
                      const float* AV, const float* W0, const float* W1, const float* eps, const float* Ds,
                      const float* RV, const int32_t* band_indices, const uint8_t* mask, const float* J_t,
                      const float* hsiao_interp, const float* weights, const double* band_log10_scale, float* out,
                      void* stream);

/* SEDmodel.get_spectra (bayesn/bayesn_model.py:556-643): host-dust-extinguished rest-frame SEDs on
 * the 300-bin model grid, same argument meaning as bayesn_flux_batch; out [N][300][N_obs]. */
int bayesn_spectra_batch(bayesn_ctx* ctx, int N, int N_obs, const float* theta, const float* AV, const float* W0,
                         const float* W1, const float* eps, const float* RV, const float* J_t, const float* hsiao_interp,
                         float* out, void* stream);

/* SEDmodel.get_flux_from_chains (bayesn/bayesn_model.py:3426-3524; called by postprocess :2269-2272 for the
 * LCPLOT file): model photometry of posterior draws on a rest-frame phase grid, ONE launch over
 * (SN x draw x phase x band) instead of the reference's per-SN loop through simulate_light_curve.
 *   t        [n_t]            device, in   rest-frame phases
 *   theta, AV, tmax, dD [N][S] device, in  per draw; dD = (mu + delM) - fold_s, the distance relative to the
 *                                          exponent 10^(-0.4 (M0 + fold_s)) folded into the SN's weight tile
 *   RV       [N][S]           device, in   per-draw R_V (rv_per_draw != 0: popRV / uniformRV chains), else NULL
 *   eps      [N][S][n_eps]    device, in   the 'eps' site (L_Sigma eps_tform), element (l-1) + (L-2) m
 *   bands    [N][max_bands]   device, in   for each SN the tile rows (band columns) to integrate, -1 = padding
 *   weights, support, wt_stride            compact band-weight tiles as in bayesn_dataset_desc
 *   out      [N][S][max_bands][n_t] device, out  FLUXCAL, or magnitudes when mag != 0 (0 for padded bands) */
int bayesn_chain_flux(bayesn_ctx* ctx, int N, int S, int n_t, int max_bands, int n_b, int mag, int rv_per_draw,
                      const float* t, const float* theta, const float* AV, const float* tmax, const float* dD,
                      const float* RV, const float* eps, const int32_t* bands, const float* weights,
                      const int32_t* support, int wt_stride, float* out, void* stream);

/* Device-resident batched NUTS over every (SN, chain) of the bound data set: replaces the
 * numpyro MCMC(NUTS(fit_model_*)) runs at bayesn/bayesn_model.py:1830-1837 and :2128-2141.
 *   q_init  [N][C][D]           device, in   initial unconstrained positions
 *   samples [N][C][S][D]        device, out  unconstrained draws after warm-up
 *   stats   [N][C][W+S][5]      device, out  per transition incl. warm-up: {accept_prob, n_leapfrog,
 *                                              diverging, potential_energy, step_size}; warm-up rows carry
 *                                              diverging + 2 when the float32 distance rescue fired after them
 *   chain_info [N][C][4+D]      device, out  {final step size, mean accept (sampling), total
 *                                              leapfrogs, divergences, inverse mass diag[D]}
 *   work, work_bytes            device scratch of at least bayesn_nuts_workspace_bytes(): the tree
 *                               vectors of every chain and the launch's work-queue counter (one
 *                               workspace per concurrent launch)
 * One asynchronous launch on `stream`.  With more than one wave of (SN, chain) units the kernel runs as
 * one resident wave whose warps fetch units from an atomic counter (draws do not depend on the
 * schedule; environment BAYESN_NUTS_SCHED=static|queue|queue-l1 overrides the choice for measurements).
 * Deviation from numpyro (float64): a chain whose adapted step size falls below 1e-6 during warm-up - a
 * float32 energy-noise stall seen for initial distances ~2 mag off - gets its distance coordinate
 * re-solved in closed form and its step-size adaptation restarted (bayesn_kernels.cuh). */
size_t bayesn_nuts_workspace_bytes(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg);
int bayesn_nuts_fit(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg, const float* q_init, float* samples, float* stats,
                    float* chain_info, void* work, size_t work_bytes, void* stream);

/* Population training: log joint density of train_model_globalRV (bayesn/bayesn_model.py:1115-1182)
 * in the unconstrained space and its gradient, for n_chains independent parameter vectors.
 *   qg [C][G]      globals: [W0 vec_F (L T), W1 vec_F (L T), u_sigmaepsilon (n_eps), y_L_Omega
 *                  (n_eps (n_eps-1)/2, strict lower triangle row-major), u_sigma0, u_RV, u_tauA];
 *                  interval sites through the affine-sigmoid bijection, L_Omega through numpyro's
 *                  CorrCholeskyTransform, density LKJCholesky(eta = 1) without its constant
 *   ql [N][C][Dl]  per-SN latents [eps_tform (n_eps), theta, log A_V, Ds], Dl = n_eps + 3
 * bayesn_train_begin evaluates this rank's SNe: grad_l [N][C][Dl] is final, red [C][R] (double,
 * device) holds the SUMS over this rank's SNe that the globals need ([log p, d/dsigma0, d/dRV,
 * d/dtauA, dW0, dW1, dL_Sigma]).  When SNe are sharded over GPUs the caller all-reduces (sum)
 * `red` between the two calls - the one exchange step of the training path.  bayesn_train_finish
 * adds the global priors / Jacobians: logp [C] (double), grad_g [C][G]. */
int bayesn_train_dims(bayesn_ctx* ctx, int* G, int* Dl, int* R);
int bayesn_train_begin(bayesn_ctx* ctx, int n_chains, const float* qg, const float* ql, float* grad_l, double* red,
                       void* stream);
int bayesn_train_finish(bayesn_ctx* ctx, int n_chains, const float* qg, double* red, double* logp,
                        float* grad_g, void* stream);

/* Device-resident NUTS over the joint training posterior (all globals + the latents of every SN),
 * replacing numpyro MCMC(NUTS(train_model_globalRV, step_size=0.1, ...)) at
 * bayesn/bayesn_model.py:1767-1770, :1913-1919.  The sampler state is sharded: (SN, chain) warps own
 * the latent slices, one CTA per chain owns the globals and all decisions.  One PASS = one leapfrog
 * of every chain = bayesn_train_nuts_sn (per-SN work + sums over this rank's SNe into red [C][R],
 * double, device) -> [caller all-reduces red over ranks when SNe are sharded] ->
 * bayesn_train_nuts_ctl.  All calls are asynchronous on `stream` and capturable in a CUDA graph.
 *   cfg             num_chains = C; reference training uses init_step_size 0.1, no regularisation
 *   sn_offset       global index of this rank's first SN (keys the latent momentum streams)
 *   qg0 [C][G], ql0 [N][C][Dl]   initial unconstrained position (device)
 *   samples_g [C][S][G], samples_l [N][C][S][Dl], stats [C][W+S][5]   outputs (device), written as
 *                   the run proceeds; layout of stats rows as in bayesn_nuts_fit
 * bayesn_train_nuts_status synchronises the stream, fills out [C][8] (HOST) = {phase (2 = done),
 * transition index, step size, mean accept, leapfrogs so far, divergences, potential energy, depth}
 * and returns 1 when every chain is done (one more bayesn_train_nuts_sn call then flushes the last
 * draw of the latents), 0 if not, <0 on error. */
int bayesn_train_nuts_init(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg, int sn_offset, const float* qg0,
                           const float* ql0, float* samples_g, float* samples_l, float* stats, void* stream);
int bayesn_train_nuts_sn(bayesn_ctx* ctx, double* red, void* stream);
/* warps per (SN, chain) unit the last bayesn_train_nuts_sn launch used (latency mode of the training SN kernel:
 * chosen from the number of units on this rank, 1 when they fill the GPU) */
int bayesn_train_warps_per_unit(bayesn_ctx* ctx);
/* Training pass with three kernels instead of four (+ no NCCL launch): with `on`, bayesn_train_nuts_sn leaves the sums
 * over this rank's SNe as per-chunk partials and bayesn_train_nuts_ctl adds them in a fixed order itself; with the
 * peer-memory exchange enabled (bayesn_train_set_exchange) the control kernel also stores them into every rank's
 * buffer, publishes / waits for the pass number and adds the ranks' slots in rank order.  `red` is then unused.
 * Switch it off when the caller all-reduces `red` itself (NCCL) between the two calls. */
int bayesn_train_fuse_reduction(bayesn_ctx* ctx, int on);
int bayesn_train_nuts_ctl(bayesn_ctx* ctx, double* red, void* stream);

/* Fused exchange of the per-leapfrog training sums over NVLink peer memory, replacing the caller's
 * all-reduce between bayesn_train_nuts_sn and bayesn_train_nuts_ctl (and between bayesn_train_begin
 * and a consumer): the second reduction stage stores this rank's sums into a slot of EVERY peer's
 * buffer and publishes a pass number; the control kernel waits for all ranks' numbers and adds the
 * slots in rank order (bit-identical on every rank).
 *   peer_bufs[p]  device address, valid in THIS process, of rank p's buffer of 2 * world * C * R doubles
 *   peer_flags[p] same for rank p's 2 * world * C uint32 flags (zero-initialised)
 *   scratch       4 zero-initialised uint32 words in this rank's own memory (pass counter, CTA counter, error)
 * The buffers must be peer-mapped symmetric memory (e.g. torch.distributed._symmetric_memory); all
 * ranks must issue the same sequence of passes.  world = 1 switches the exchange off. */
int bayesn_train_set_exchange(bayesn_ctx* ctx, int world, int rank, const uint64_t* peer_bufs,
                              const uint64_t* peer_flags, void* scratch);
int bayesn_train_nuts_status(bayesn_ctx* ctx, float* out, void* stream);

/* Per-SN band-weight tiles built on the device: SEDmodel._calculate_band_weights
 * (bayesn/bayesn_model.py:499-554: redshifted linear interpolation of the x51-oversampled band master
 * curves onto the 300 model bins, normalisation, Fitzpatrick99 Milky Way dust at R_V = 3.1, 1/(1+z))
 * followed by the scale folding / float32 rounding / support compaction of bayesn_dataset_desc.weights
 * - without ever materialising the (N, 300, n_b) float64 array.  All pointers are DEVICE pointers.
 *   master [n_b][n_bw]  band_interpolate_weights rows of the bands in play (column 0 = NULL band)
 *   z, ebv, fold [N]    redshift, Milky Way E(B-V), and M0 + float32(muhat_s) (training: minus
 *                       (ybar_s - 27.5)); the log10 scale of (s, k) is bscale[k] - 0.4 fold[s]
 *   used [N]            bit k set <=> SN s has a valid observation in band k (other rows stay empty)
 *   f99_xk/yk/mk        knots (1/micron), knot values and second derivatives of the F99 k(lambda-V) spline at R_V,MW
 * bayesn_tiles_support fills support [N][n_b][4] = {first, last, offset, 0} and total_width [N]; the
 * caller allocates weights [N][wt_stride] (+ BAYESN_OVERRUN floats), wt_stride = max total_width
 * rounded up to a multiple of 4 and zero-initialised, and calls bayesn_tiles_fill. */
typedef struct {
    int32_t N, n_b, n_bw;
    const double* master;
    const double* z;
    const double* ebv;
    const double* fold;
    const double* bscale;
    const uint32_t* used;
    const double* model_wave;
    double band_spacing, RV_MW;
    double f99_xk[9], f99_yk[9], f99_mk[9];
} bayesn_tiles_desc;
int bayesn_tiles_support(bayesn_ctx* ctx, const bayesn_tiles_desc* desc, int32_t* support, int32_t* total_width,
                         void* stream);
int bayesn_tiles_fill(bayesn_ctx* ctx, const bayesn_tiles_desc* desc, const int32_t* support, int wt_stride,
                      float* weights, void* stream);

/* numpyro init_to_median(num_samples=15) (reference bayesn/bayesn_model.py:1757, :2110): initial
 * unconstrained positions q [N][n_chains][D] (device, out) for the bound data set. */
int bayesn_init_to_median(bayesn_ctx* ctx, int n_chains, uint64_t seed, uint64_t unit_offset, float* q, void* stream);

/* Test aid: fill every SM's shared memory with NaN bit patterns (what an unrelated earlier kernel may leave behind);
 * the evaluation kernels must be insensitive to it. */
int bayesn_debug_poison_smem(bayesn_ctx* ctx, void* stream);

/* FP32 FMA peak of this device in TFLOP/s, measured with an FFMA-only kernel: the roofline
 * denominator for the SIMT kernels (no reference counterpart; measurement aid). */
int bayesn_ffma_peak(bayesn_ctx* ctx, double* tflops_out);

/* Number of kernel launches this ctx has issued (for bench accounting). */
long long bayesn_launch_count(bayesn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
