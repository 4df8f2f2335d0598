"""Generates the committed golden fixtures.  The reference (numpyro/JAX) cannot be imported in
this image, so the vectors come from the float64 oracle (``oracle/ref_model.py``), which is
itself pinned against the reference's only published numbers (``tests/test_oracle.py``).

    python tests/golden/make_golden.py            # logp/grad + flux vectors (seconds)
    python tests/golden/make_golden.py posterior  # C1 posterior with the oracle NUTS (~10-20 min)
    python tests/golden/make_golden.py posterior_w22 [threads]   # 16 W22 ugrizYJH SNe (C3 recipe), ~1 h
    python tests/golden/make_golden.py posterior_lsst [threads]  # 8 W22 ugrizy LSST SNe (C5 recipe), ~1 h
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.common import make_case  # noqa: E402

CASES = {
    'T21_griz': dict(model_name='T21_model', bands=['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'], t=np.arange(-8, 40, 4.0),
                     N=6, seed=101),
    'W22_ugrizYJH': dict(model_name='W22_model',
                         bands=['u_LSST', 'g_PS1', 'r_PS1', 'i_PS1', 'z_PS1', 'Y_WFCAM', 'J_WFCAM', 'H_WFCAM'],
                         t=np.array([-8.0, 2.0, 12.0, 22.0, 32.0]), N=6, seed=102),
    'M20_csp': dict(model_name='M20_model', bands=['B_CSP', 'V_CSP', 'r_CSP', 'i_CSP', 'Y_RC', 'J_RC1', 'H_RC'],
                    t=np.arange(-8, 40, 4.0), N=4, seed=103, ebv_mw=0.03),
    'popRV_griz': dict(model_name='pop_RV_model', bands=['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'],
                       t=np.arange(-8, 40, 4.0), N=4, seed=104),
}


def golden_logp():
    import torch
    out = {}
    for name, kw in CASES.items():
        case = make_case(**kw)
        ref = case['ref']
        mode = 'fixed' if ref.model_type == 'fixed_RV' else 'pop'
        n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
        n_sc = 4 if mode == 'fixed' else 5
        rng = np.random.RandomState(kw['seed'])
        N = kw['N']
        q = np.zeros((N, n_sc + n_eps), dtype=np.float32)
        q[:, 0] = rng.randn(N)
        q[:, 1] = np.log(rng.exponential(ref.tauA, N) + 0.02)
        q[:, 2] = 0.3 * rng.randn(N)
        q[:, 3] = case['obs'][7, 0, :] + 0.2 * rng.randn(N)
        if mode == 'pop':
            q[:, 4] = 0.5 * rng.randn(N)
        q[:, n_sc:] = rng.randn(N, n_eps)
        lp, g = ref.fit_logp_grad(q.astype(np.float64), case['obs'], case['weights'], rv_mode=mode)
        _, flux = ref.fit_logp(torch.as_tensor(q.astype(np.float64)), case['obs'], case['weights'], rv_mode=mode,
                               return_flux=True)
        out[name] = dict(obs=case['obs'], weights=case['weights'].astype(np.float64), z=case['z'], ebv=case['ebv'],
                         q=q, logp=lp, grad=g, flux=flux.numpy(), band_names=np.array(case['band_names']))
        np.savez_compressed(os.path.join(HERE, f'logp_{name}.npz'), **out[name])
        print(name, lp)


def golden_posterior():
    from oracle import ref_fit
    kw = dict(model_name='T21_model', bands=['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'], t=np.arange(-8, 40, 4.0), N=10,
              seed=20241019, mask_some=False)
    case = make_case(**kw)
    res = ref_fit.fit_cpu(case['ref'], case['obs'], case['weights'], n_chains=4, num_warmup=250, num_samples=250,
                          seed=0)
    c = ref_fit.constrain(case['ref'], res['samples'])
    np.savez_compressed(os.path.join(HERE, 'c1_oracle_posterior.npz'), obs=case['obs'], weights=case['weights'],
                        z=case['z'], band_names=np.array(case['band_names']),
                        theta=c['theta'], AV=c['AV'], tmax=c['tmax'], Ds=c['Ds'], n_steps=res['n_steps'],
                        diverging=res['diverging'], seconds=res['seconds'], evals=res['evals'], ticks=res['ticks'],
                        true_theta=case['params']['theta'], true_AV=case['params']['AV'],
                        true_mu=case['params']['mu'], true_delM=case['params']['del_M'])
    print('done', res['seconds'], res['evals'])


W22_POSTERIOR = {
    'w22': dict(model_name='W22_model',
                bands=['u_LSST', 'g_PS1', 'r_PS1', 'i_PS1', 'z_PS1', 'Y_WFCAM', 'J_WFCAM', 'H_WFCAM'],
                t=np.array([-8.0, 2.0, 12.0, 22.0, 32.0]), N=16, seed=20241020, mask_some=False),
    'lsst': dict(model_name='W22_model', bands=['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST'],
                 t=np.arange(-8, 40, 5.0), N=8, seed=20241021, mask_some=False, zrange=(0.02, 0.09)),
}


def golden_posterior_w22(which, threads=None):
    """Oracle-sampler posteriors on the W22 model: the C3 recipe (ugrizYJH, 40 obs) and the C5 recipe
    (LSST ugrizy, 60 obs), 4 x (250 + 250) each - the float64 counterpart of the shapes where the FP32
    sampler needed its safeguard."""
    from oracle import ref_fit
    kw = W22_POSTERIOR[which]
    case = make_case(**kw)
    res = ref_fit.fit_cpu(case['ref'], case['obs'], case['weights'], n_chains=4, num_warmup=250, num_samples=250,
                          seed=0, threads=threads)
    c = ref_fit.constrain(case['ref'], res['samples'])
    np.savez_compressed(os.path.join(HERE, f'{which}_oracle_posterior.npz'), obs=case['obs'],
                        weights=case['weights'].astype(np.float32), z=case['z'], band_names=np.array(case['band_names']),
                        theta=c['theta'].astype(np.float32), AV=c['AV'].astype(np.float32),
                        tmax=c['tmax'].astype(np.float32), Ds=c['Ds'], n_steps=res['n_steps'], accept=res['accept'],
                        diverging=res['diverging'], all_steps=res['all_steps'], seconds=res['seconds'],
                        evals=res['evals'], ticks=res['ticks'], threads=threads or 0,
                        true_theta=case['params']['theta'], true_AV=case['params']['AV'],
                        true_mu=case['params']['mu'], true_delM=case['params']['del_M'])
    print('done', which, res['seconds'], res['evals'])


def golden_train():
    """Training log density (train_model_globalRV, unconstrained) and gradient at a seeded point."""
    from tests.test_gpu_train import training_point
    case = make_case('T21_model', ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'], np.arange(-8, 40, 4.0), N=6, seed=31, mag=True,
                     mask_some=True)
    ref = case['ref']
    qg, ql = training_point(ref, case, 1, 8)
    qg32, ql32 = qg[0].astype(np.float32), ql[:, 0].astype(np.float32)
    lp, gg, gl = ref.train_logp_grad(qg32.astype(np.float64), ql32.astype(np.float64), case['obs'], case['weights'])
    np.savez_compressed(os.path.join(HERE, 'train_T21_griz.npz'), obs=case['obs'], weights=case['weights'], z=case['z'],
                        ebv=case['ebv'], qg=qg32, ql=ql32, logp=lp, grad_g=gg, grad_l=gl,
                        band_names=np.array(case['band_names']))
    print('train', lp)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'train':
        golden_train()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] in ('posterior_w22', 'posterior_lsst'):
        golden_posterior_w22(sys.argv[1].split('_')[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'posterior':
        golden_posterior()
    else:
        golden_logp()
