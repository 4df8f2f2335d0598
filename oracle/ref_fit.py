"""ORACLE (test infrastructure): per-object NUTS fits driven by the float64 restatement -
the CPU path that is timed beside the GPU numbers and that pins posterior parity."""
import math
import time

import numpy as np
import torch

from . import ref_nuts


def fit_cpu(ref, obs, weights, n_chains=4, num_warmup=250, num_samples=250, max_depth=10, seed=0, rv_mode='fixed',
            q_init=None, threads=None):
    """Fit every SN in ``obs`` (10, N_obs, N) with ``n_chains`` chains each, all chains in
    lock-step.  Parameter layout is the oracle's ``[theta, u_AV, u_tmax, Ds, (u_RV,) eps_tform]``.
    Returns dict(samples (N, C, S, D), accept, n_steps, diverging, ticks, evals, seconds)."""
    if threads:
        torch.set_num_threads(threads)
    obs = np.asarray(obs)
    N = obs.shape[-1]
    n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
    n_sc = 4 if rv_mode == 'fixed' else 5
    D = n_eps + n_sc
    if q_init is None:
        qk = ref_nuts.init_to_median_device_compat(obs[7, 0, :], n_chains, D, n_eps, ref.tauA,
                                                   math.sqrt(25 + ref.sigma0 ** 2), seed=seed)
        q_init = np.concatenate([qk[..., n_eps:], qk[..., :n_eps]], axis=-1)     # kernel -> oracle layout
    B = N * n_chains
    sn_of = np.repeat(np.arange(N), n_chains)
    q0 = q_init.reshape(B, D)
    obs_t = torch.as_tensor(obs)
    w_t = torch.as_tensor(weights)

    def lg(q, rows):
        idx = sn_of[rows]
        qt = torch.as_tensor(q).clone().requires_grad_(True)
        lp = ref.fit_logp(qt, obs_t[:, :, idx], w_t[idx], rv_mode=rv_mode)
        g, = torch.autograd.grad(lp.sum(), qt)
        return lp.detach().numpy(), g.numpy()

    t0 = time.perf_counter()
    recs, ticks, evals = ref_nuts.run_chains(lg, q0, num_warmup, num_samples, unit_ids=list(range(B)),
                                             max_depth=max_depth, seed=seed)
    dt = time.perf_counter() - t0
    S = num_samples
    samples = np.stack([np.stack(r['z'][num_warmup:]) for r in recs]).reshape(N, n_chains, S, D)
    pack = lambda k: np.array([r[k][num_warmup:] for r in recs]).reshape(N, n_chains, S)
    return dict(samples=samples, accept=pack('accept'), n_steps=pack('n_steps'), diverging=pack('diverging'),
                all_steps=np.array([np.sum(r['n_steps']) for r in recs]).reshape(N, n_chains),
                ticks=ticks, evals=evals, seconds=dt)


def constrain(ref, samples, rv_mode='fixed'):
    """(..., D) oracle-layout unconstrained draws -> dict of constrained arrays."""
    out = {'theta': samples[..., 0], 'AV': np.exp(samples[..., 1]),
           'tmax': -10 + 20 / (1 + np.exp(-samples[..., 2])), 'Ds': samples[..., 3]}
    n_sc = 4 if rv_mode == 'fixed' else 5
    out['eps_tform'] = samples[..., n_sc:]
    if rv_mode == 'uniform':
        out['RV'] = 1 + 5 / (1 + np.exp(-samples[..., 4]))
    elif rv_mode == 'pop':
        out['RV_tform'] = 1 / (1 + np.exp(-samples[..., 4]))
    return out


def train_cpu(ref, obs, weights, qg0, ql0, chain=0, n_chains=1, num_warmup=0, num_samples=5, max_depth=5,
              init_step=0.01, seed=0, adapt_step=False, adapt_mass=False, sn_offset=0):
    """One chain of the joint training posterior (train_model_globalRV) with the float64 oracle
    density and the same sampler semantics / random streams as the device sampler.  ``qg0`` (G,),
    ``ql0`` (N, Dl) in the kernel's latent layout.  Returns dict of per-transition records; ``z``
    rows are [globals, latents...]."""
    qg0, ql0 = np.asarray(qg0, np.float64), np.asarray(ql0, np.float64)
    G, (N, Dl) = qg0.shape[0], ql0.shape
    nrm_fn, unit = ref_nuts.train_momentum_layout(G, N, Dl, n_chains, chain, sn_offset)
    rec = {k: [] for k in ('z', 'accept', 'n_steps', 'diverging', 'pe', 'step')}
    gen = ref_nuts.chain_nuts(np.concatenate([qg0, ql0.ravel()]), unit, num_warmup, num_samples, max_depth=max_depth,
                              init_step=init_step, seed=seed, adapt_step=adapt_step, adapt_mass=adapt_mass,
                              regularize=False, record=rec, nrm_fn=nrm_fn)
    z = next(gen)
    evals = 0
    try:
        while True:
            lp, gg, gl = ref.train_logp_grad(z[:G], z[G:].reshape(N, Dl), obs, weights)
            evals += 1
            z = gen.send((lp, np.concatenate([gg, gl.ravel()])))
    except StopIteration:
        pass
    rec['evals'] = evals
    return rec
