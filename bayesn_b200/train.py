"""Population training (``train_model_globalRV``, reference ``bayesn/bayesn_model.py:1115-1182``;
driver ``run`` :1766-1770, :1913-1924; ``initial_guess`` :1568-1654; ``postprocess`` :2186-2242):
host side of the device-resident training sampler.

Unconstrained layouts (``include/bayesn_b200.h``): globals ``[W0 vec_F, W1 vec_F, u_sigmaepsilon,
y_L_Omega, u_sigma0, u_RV, u_tauA]``; per-SN latents ``[eps_tform, theta, log A_V, Ds]``.
"""
import math
import os
import pickle

import numpy as np

from . import datafiles
from .static import inv_kd, spline_matrix

HALF_PI = math.pi / 2


def _logit(x):
    return np.log(x) - np.log1p(-x)


def dims(model):
    L, T = len(model.l_knots), len(model.tau_knots)
    ne = (L - 2) * T
    return dict(L=L, T=T, n_eps=ne, n_corr=ne * (ne - 1) // 2, G=2 * L * T + ne + ne * (ne - 1) // 2 + 3, Dl=ne + 3)


def simulate_training_set(model, N, seed=0, bands=('B_CSP', 'V_CSP', 'r_CSP', 'i_CSP', 'Y_RC', 'J_RC1', 'H_RC'),
                          t=np.arange(-8, 40, 4.0), zrange=(0.015, 0.08), yerr=0.05):
    """``simulate_light_curve(mag=True)`` -> training arrays in the
    reference layout (``obs[1]`` magnitudes, ``obs[6]`` z error 1e-4)."""
    np.random.seed(seed)
    z = np.random.uniform(zrange[0], zrange[1], N)
    data, err, params = model.simulate_light_curve(t, N, list(bands), yerr=yerr, err_type='mag', z=z, zerr=1e-4,
                                                   mu='z', ebv_mw=0, mag=True)
    obs, weights, band_inds = model.simulated_obs_array(data, err, params)
    obs[9] = np.isfinite(obs[1]) & (obs[9] != 0)          # training masks non-positive fluxes (:2714-2718)
    obs[1] = np.where(obs[9] != 0, obs[1], 0.0)
    return obs, weights, band_inds


def initial_values(model, obs, reference_model='M20_model', rng=None, mode='training_globalRV'):
    """``SEDmodel.initial_guess`` (reference :1568-1654): W0 / W1 of a trained model interpolated
    onto this model's knots (cubic in wavelength, linear in time, zero outside), jittered; fixed
    starting values for the other globals; prior draws for the latents."""
    from scipy.interpolate import interp1d
    rng = np.random if rng is None else rng
    if os.path.exists(reference_model):
        params = datafiles.load_yaml(reference_model)
    elif reference_model in datafiles.builtin_models():
        params = datafiles.load_model_yaml(reference_model)
    else:
        raise ValueError("Invalid initialisation method, please choose either 'median' or 'sample', or choose "
                         "either one of the built-in models or a custom model to base the hyperparmeter "
                         "initialisation on")
    W0, W1 = np.array(params['W0'], dtype=float), np.array(params['W1'], dtype=float)
    lk, tk = np.array(params['L_KNOTS'], dtype=float), np.array(params['TAU_KNOTS'], dtype=float)
    tauA_init = float(params['TAUA'])
    W0 = interp1d(lk, W0, kind='cubic', axis=0, fill_value=0, bounds_error=False)(model.l_knots)
    W1 = interp1d(lk, W1, kind='cubic', axis=0, fill_value=0, bounds_error=False)(model.l_knots)
    W0 = interp1d(tk, W0, kind='linear', axis=1, fill_value=0, bounds_error=False)(model.tau_knots)
    W1 = interp1d(tk, W1, kind='linear', axis=1, fill_value=0, bounds_error=False)(model.tau_knots)
    W0, W1 = W0.flatten(order='F'), W1.flatten(order='F')
    d = dims(model)
    ne, n_sne = d['n_eps'], obs.shape[-1]
    tauA_ = tauA_init + rng.normal(0, 0.01)
    while tauA_ < 0:
        tauA_ = tauA_init + rng.normal(0, 0.01)
    sigma0_ = 0.1 + rng.normal(0, 0.01)
    sigmaepsilon_init = 0.1 * np.ones(ne)
    # same keys, same order of random draws as the reference (:1626-1652), so that a seeded numpy RNG gives the
    # same dictionary; the sampler takes the entries that name sample sites of train_model_globalRV
    init = {'W0': W0 + rng.normal(0, 0.01, W0.shape[0]), 'W1': W1 + rng.normal(0, 0.01, W1.shape[0])}
    if 'poprv' in mode.lower():
        init['mu_R'], init['sigma_R'] = 3.0, 0.5
        init['RV_tform'] = rng.uniform(0, 1, n_sne)
    else:
        init['RV'] = 3.0
    init['tauA_tform'] = math.atan(tauA_ / 1.0)
    init['sigma0_tform'] = math.atan(sigma0_ / 0.1)
    init['sigma0'] = sigma0_
    init['theta'] = rng.normal(0, 1, n_sne)
    init['AV'] = rng.exponential(tauA_, n_sne)
    L_Sigma = np.diag(sigmaepsilon_init) @ np.eye(ne)
    init['epsilon_tform'] = np.linalg.inv(L_Sigma) @ rng.normal(0, 1, (ne, n_sne))
    init['epsilon'] = rng.normal(0, 1, (n_sne, ne))
    init['sigmaepsilon_tform'] = np.arctan(sigmaepsilon_init + rng.normal(0, 0.01, ne) / 1.0)
    init['sigmaepsilon'] = sigmaepsilon_init + rng.normal(0, 0.01, ne)
    init['L_Omega'] = np.eye(ne)
    init['Ds'] = rng.normal(obs[-3, 0, :], sigma0_)
    return init


def unconstrain(model, init, n_sne, rng=None):
    """Constrained initial values -> (qg (G,), ql (N, Dl)).  Sites the reference's dictionary
    misses (it spells ``eps_tform`` as ``epsilon_tform``, :1645) fall back to numpyro's
    ``init_to_uniform(radius=2)`` in the unconstrained space, as ``init_to_value`` does."""
    rng = np.random if rng is None else rng
    d = dims(model)
    L, T, ne = d['L'], d['T'], d['n_eps']
    qg = np.zeros(d['G'])
    qg[:L * T] = init['W0']
    qg[L * T:2 * L * T] = init['W1']
    qg[2 * L * T:2 * L * T + ne] = _logit(np.clip(np.asarray(init['sigmaepsilon_tform']) / HALF_PI, 1e-6, 1 - 1e-6))
    Lo = np.asarray(init.get('L_Omega', np.eye(ne)))
    qg[2 * L * T + ne:d['G'] - 3] = corr_cholesky_inverse(Lo)
    qg[-3] = _logit(init['sigma0_tform'] / HALF_PI)
    qg[-2] = _logit((init['RV'] - 1.0) / 4.0)
    qg[-1] = _logit(init['tauA_tform'] / HALF_PI)
    ql = np.zeros((n_sne, d['Dl']))
    ql[:, :ne] = init['eps_tform'] if 'eps_tform' in init else rng.uniform(-2, 2, (n_sne, ne))
    ql[:, ne] = init['theta']
    ql[:, ne + 1] = np.log(np.maximum(init['AV'], 1e-8))
    ql[:, ne + 2] = init['Ds']
    return qg, ql


def prior_draw(model, obs, rng, sigma_pec=150 / 3e5):
    """One draw of every site of ``train_model_globalRV`` from its prior (reference :1131-1177), constrained
    values in the layout of :func:`initial_values` (the AV and Ds priors are conditional on the drawn tau_A /
    sigma0, as a forward run of the numpyro model draws them).  L_Omega ~ LKJCholesky(eta = 1) through the C-vine:
    the partial correlation of column j is 2 Beta(b_j, b_j) - 1 with b_j = 1 + (n - 2 - j) / 2, and those are
    exactly the tanh(y) of the signed stick breaking."""
    d = dims(model)
    L, T, ne, n = d['L'], d['T'], d['n_eps'], obs.shape[-1]
    tauA_t, sigma0_t = rng.uniform(0, HALF_PI), rng.uniform(0, HALF_PI)
    tauA, sigma0 = math.tan(tauA_t), 0.1 * math.tan(sigma0_t)
    ii, jj = np.tril_indices(ne, -1)
    b = 1.0 + (ne - 2 - jj) / 2.0
    r = 2.0 * rng.beta(b, b) - 1.0
    zs, zerr = obs[-5, 0, :], obs[-4, 0, :]
    muerr = 5 / (zs * math.log(10)) * np.sqrt(zerr ** 2 + sigma_pec ** 2)
    return {'W0': rng.normal(0, 1, L * T), 'W1': rng.normal(0, 1, L * T), 'RV': rng.uniform(1, 5),
            'tauA_tform': tauA_t, 'sigma0_tform': sigma0_t, 'theta': rng.normal(0, 1, n),
            'AV': rng.exponential(tauA, n), 'sigmaepsilon_tform': rng.uniform(0, HALF_PI, ne),
            '_y_corr': np.arctanh(np.clip(r, -1 + 1e-7, 1 - 1e-7)), 'eps_tform': rng.normal(0, 1, (n, ne)),
            'Ds': rng.normal(obs[-3, 0, :], np.sqrt(muerr ** 2 + sigma0 ** 2))}


def prior_position(model, obs, rng, how):
    """``initialisation: median | sample`` (reference :1756-1759): numpyro's ``init_to_median`` (median of 15 prior
    draws per site, taken in the constrained space - for L_Omega on the partial correlations, whose prior median is
    the identity) or ``init_to_sample`` (one prior draw) -> (qg (G,), ql (N, Dl))."""
    n = obs.shape[-1]
    draws = [prior_draw(model, obs, rng) for _ in range(15 if how == 'median' else 1)]
    init = {k: np.median(np.stack([np.asarray(dr[k], dtype=np.float64) for dr in draws]), axis=0) for k in draws[0]}
    y = np.arctanh(np.median(np.stack([np.tanh(dr['_y_corr']) for dr in draws]), axis=0))
    init['L_Omega'] = np.eye(dims(model)['n_eps'])
    qg, ql = unconstrain(model, init, n, rng)
    d = dims(model)
    qg[2 * d['L'] * d['T'] + d['n_eps']:d['G'] - 3] = y
    return qg, ql


def initial_position(model, obs, num_chains=4, reference_model='M20_model', seed=0, device='cuda'):
    """Starting point of the training chains: float32 CUDA tensors qg (C, G), ql (N, C, Dl).  A model name / yaml
    path starts every chain from ``initial_guess`` (the reference notes numpyro cannot initialise chains
    separately, :1572-1573); ``median`` / ``sample`` start every chain from its own prior median / draw
    (:1756-1759)."""
    import torch
    rng = np.random.RandomState(seed)
    n = obs.shape[-1]
    if reference_model in ('median', 'sample'):
        pos = [prior_position(model, obs, np.random.RandomState(seed * 1000 + c), reference_model) for c in range(num_chains)]
        qg_t = torch.tensor(np.stack([p[0] for p in pos]), dtype=torch.float32, device=device).contiguous()
        ql_t = torch.tensor(np.stack([p[1] for p in pos], axis=1), dtype=torch.float32, device=device).contiguous()
        return qg_t, ql_t
    qg, ql = unconstrain(model, initial_values(model, obs, reference_model, rng), n, rng)
    qg_t = torch.tensor(np.repeat(qg[None], num_chains, axis=0), dtype=torch.float32, device=device).contiguous()
    ql_t = torch.tensor(np.repeat(ql[:, None], num_chains, axis=1), dtype=torch.float32, device=device).contiguous()
    return qg_t, ql_t


def corr_cholesky(y, n):
    """numpyro ``CorrCholeskyTransform`` forward (tanh + signed stick breaking) in numpy: y (..., n(n-1)/2)."""
    y = np.asarray(y, dtype=np.float64)
    t = np.tanh(y)
    out = np.zeros(y.shape[:-1] + (n, n))
    ii, jj = np.tril_indices(n, -1)
    r = np.zeros_like(out)
    r[..., ii, jj] = t
    cum = np.cumprod(1 - r * r, axis=-1)
    shifted = np.concatenate([np.ones(cum.shape[:-1] + (1,)), np.sqrt(cum[..., :-1])], axis=-1)
    out = (r + np.eye(n)) * shifted
    return out


def corr_cholesky_inverse(Lo):
    """Inverse of :func:`corr_cholesky` for one matrix."""
    n = Lo.shape[0]
    y = np.zeros(n * (n - 1) // 2)
    k = 0
    for i in range(1, n):
        rem = 1.0
        for j in range(i):
            t = Lo[i, j] / math.sqrt(rem) if rem > 0 else 0.0
            t = min(max(t, -1 + 1e-7), 1 - 1e-7)
            y[k] = math.atanh(t)
            rem *= 1 - t * t
            k += 1
    return y


def constrain(model, samples_g, samples_l, muhat=None):
    """Unconstrained draws -> dict with the reference's site names and ``group_by_chain`` shapes
    (chains, draws, ...), as ``mcmc.get_samples(group_by_chain=True)`` returns for
    ``train_model_globalRV``.  ``samples_g`` (C, S, G), ``samples_l`` (N, C, S, Dl) numpy arrays."""
    d = dims(model)
    L, T, ne = d['L'], d['T'], d['n_eps']
    sg = np.asarray(samples_g, dtype=np.float64)
    sl = np.asarray(samples_l, dtype=np.float64)
    sig = lambda u: 1.0 / (1.0 + np.exp(-u))
    out = {'W0': sg[..., :L * T], 'W1': sg[..., L * T:2 * L * T]}
    out['sigmaepsilon_tform'] = HALF_PI * sig(sg[..., 2 * L * T:2 * L * T + ne])
    out['sigmaepsilon'] = np.tan(out['sigmaepsilon_tform'])
    out['L_Omega'] = corr_cholesky(sg[..., 2 * L * T + ne:d['G'] - 3], ne)
    out['sigma0_tform'] = HALF_PI * sig(sg[..., -3])
    out['sigma0'] = 0.1 * np.tan(out['sigma0_tform'])
    out['RV'] = 1 + 4 * sig(sg[..., -2])
    out['tauA_tform'] = HALF_PI * sig(sg[..., -1])
    out['tauA'] = np.tan(out['tauA_tform'])
    lat = sl.transpose(1, 2, 0, 3)                               # (C, S, N, Dl)
    out['eps_tform'] = lat[..., :ne]
    out['theta'] = lat[..., ne]
    out['AV'] = np.exp(lat[..., ne + 1])
    out['Ds'] = lat[..., ne + 2]
    L_Sigma = out['sigmaepsilon'][..., :, None] * out['L_Omega']
    out['eps'] = np.einsum('csij,csnj->csin', L_Sigma, out['eps_tform'])     # (C, S, n_eps, N) as in the reference
    return out


def postprocess_training(model, samples, outputdir=None, mode='training_globalRV'):
    """Reference ``postprocess`` :2186-2242: W1 / theta sign convention and rescaling, posterior-mean
    model written as ``bayesn.yaml`` (+ ``initial_chains.pkl`` / ``fit_summary.csv`` / ``chains.pkl``, :2358-2364).
    Returns the YAML dict and the post-processed samples."""
    import yaml
    samples = dict(samples)
    if outputdir is not None:
        os.makedirs(outputdir, exist_ok=True)
        with open(os.path.join(outputdir, 'initial_chains.pkl'), 'wb') as fh:
            pickle.dump(samples, fh)
    L, T = len(model.l_knots), len(model.tau_knots)
    J_R = spline_matrix(np.array([6200.0]), model.l_knots, inv_kd(model.l_knots))
    J_10 = spline_matrix(np.array([10.0]), model.tau_knots, inv_kd(model.tau_knots))
    J_0 = spline_matrix(np.array([0.0]), model.tau_knots, inv_kd(model.tau_knots))
    W1 = np.reshape(samples['W1'], (samples['W1'].shape[0], samples['W1'].shape[1], L, T), order='F')
    sign = np.zeros(W1.shape[0])
    for chain in range(W1.shape[0]):
        cw = np.mean(W1[chain], axis=0)
        sign[chain] = np.sign(np.squeeze(J_R @ cw @ J_10.T) - np.squeeze(J_R @ cw @ J_0.T))
    samples['W1'] = samples['W1'] * sign[:, None, None]
    samples['theta'] = samples['theta'] * sign[:, None, None]
    theta_std = np.std(samples['theta'], axis=2)
    samples['theta'] = samples['theta'] / theta_std[..., None]
    samples['W1'] = samples['W1'] * theta_std[..., None]
    W0m = np.mean(samples['W0'], axis=(0, 1)).reshape((L, T), order='F')
    W1m = np.mean(samples['W1'], axis=(0, 1)).reshape((L, T), order='F')
    L_Sigma = np.diag(np.mean(samples['sigmaepsilon'], axis=(0, 1))) @ np.mean(samples['L_Omega'], axis=(0, 1))
    yaml_data = {'M0': float(model.M0), 'SIGMA0': float(np.mean(samples['sigma0'])), 'TAUA': float(np.mean(samples['tauA'])),
                 'TAU_KNOTS': model.tau_knots.tolist(), 'L_KNOTS': model.l_knots.tolist(), 'W0': W0m.tolist(),
                 'W1': W1m.tolist(), 'L_SIGMA_EPSILON': L_Sigma.tolist()}
    if 'singlerv' in mode.lower():            # reference quirk (:2235): globalRV runs do not write RV
        yaml_data['RV'] = float(np.mean(samples['RV']))
    if outputdir is not None:
        with open(os.path.join(outputdir, 'bayesn.yaml'), 'w') as fh:
            yaml.safe_dump(yaml_data, fh)
        # convergence table of every parameter, then the post-processed chains (reference :2358-2364)
        from . import summary
        summary.summary_table(samples).to_csv(os.path.join(outputdir, 'fit_summary.csv'))
        with open(os.path.join(outputdir, 'chains.pkl'), 'wb') as fh:
            pickle.dump(samples, fh)
    return yaml_data, samples
