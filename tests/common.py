"""Shared helpers for the parity tests: simulated cases built with the ORACLE (CPU, float64) and
the comparison of the CUDA path against it."""
import numpy as np

from oracle.ref_model import RefSED

_REF_CACHE = {}


def ref_model(model_name, bands):
    key = (model_name, tuple(sorted(bands)))
    if key not in _REF_CACHE:
        _REF_CACHE[key] = RefSED(model_name, bands=list(bands))
    return _REF_CACHE[key]


def make_case(model_name, bands, t, N=8, seed=1, mask_some=True, ebv_mw=0.0, zrange=(0.015, 0.08), mag=False,
              yerr=0.05):
    """simulate_light_curve-style inputs (SURVEY.md section 8d recipe) from the oracle."""
    np.random.seed(seed)
    ref = ref_model(model_name, bands)
    z = np.random.uniform(zrange[0], zrange[1], N)
    ebv = np.full(N, ebv_mw)
    RV = None if ref.model_type == 'fixed_RV' else ref.mu_R
    data, yerr_a, pd, (tt, bidx, bw) = ref.simulate_light_curve(t, N, bands, yerr=yerr, z=z, mu='z', mag=mag,
                                                                ebv_mw=ebv, err_type='mag', RV=RV)
    N_obs = data.shape[0]
    obs = np.zeros((10, N_obs, N))
    obs[0], obs[1], obs[2], obs[4] = tt, data, yerr_a, bidx
    obs[5], obs[6], obs[7], obs[8], obs[9] = z[None], 1e-4, ref.distmod(z)[None], ebv[None], 1
    if mask_some:
        obs[9, ::7, ::2] = 0
    names = [ref.inv_band_dict[k] for k in range(len(ref.band_dict))]
    return dict(ref=ref, model_name=model_name, obs=obs, weights=bw, z=z, ebv=ebv, params=pd, band_names=names)


def product_model(case, **kw):
    from bayesn_b200 import SEDmodel
    m = SEDmodel(load_model=case['model_name'], **kw)
    band_inds = np.array([m.band_dict[n] for n in case['band_names']])
    return m, band_inds


def to_oracle_layout(qk, ne):
    return np.concatenate([qk[..., ne:], qk[..., :ne]], axis=-1)


def to_kernel_layout(qo, n_sc):
    return np.concatenate([qo[..., n_sc:], qo[..., :n_sc]], axis=-1)


def compare_logp_grad(case, n_chains=2, rv=None, seed=3, jitter=0.3, fix=None):
    """Evaluate the CUDA kernel (through the C ABI) and the oracle at the same float32 points."""
    import torch
    from bayesn_b200 import native
    ref = case['ref']
    fix = fix or {}
    m, band_inds = product_model(case)
    z, ebv = case['z'], case['ebv']
    w_prod = m._calculate_band_weights(z, ebv, band_inds)
    w_rel = float(np.max(np.abs(w_prod - case['weights'])) / np.max(np.abs(case['weights'])))
    rv_mode = native.RV_UNIFORM if rv == 'uniform' else (
        native.RV_FIXED if ref.model_type == 'fixed_RV' else native.RV_POP)
    ctx = m.prepare(case['obs'], w_prod, band_inds, rv_mode, **fix)
    ne, D = m.n_eps, ctx.D
    n_sc = D - ne
    rng = np.random.RandomState(seed)
    q0 = ctx.init_to_median(n_chains, seed=seed).cpu().numpy().astype(np.float64)          # device kernel (numpyro init_to_median)
    q0 += jitter * rng.randn(*q0.shape)
    q32 = torch.tensor(q0, dtype=torch.float32, device='cuda').contiguous()
    lp, g = ctx.logp_grad(q32)
    torch.cuda.synchronize()
    lp, g = lp.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64)
    qk = q32.cpu().numpy().astype(np.float64)
    mode = {native.RV_FIXED: 'fixed', native.RV_POP: 'pop', native.RV_UNIFORM: 'uniform'}[rv_mode]
    okw = {}
    if fix:
        okw = dict(fix_tmax=fix.get('fix_tmax', False), fix_theta=fix.get('fix_theta', False),
                   theta_val=fix.get('theta_val', 0.0), fix_AV=fix.get('fix_AV', False), AV_val=fix.get('AV_val', 0.0))
    worst_lp = worst_g = 0.0
    for c in range(n_chains):
        lpo, go = ref.fit_logp_grad(to_oracle_layout(qk[:, c], ne), case['obs'], case['weights'], rv_mode=mode, **okw)
        go = to_kernel_layout(go, n_sc)
        assert np.all(np.isfinite(lpo))
        worst_lp = max(worst_lp, float(np.max(np.abs(lp[:, c] - lpo) / np.abs(lpo))))
        worst_g = max(worst_g, float(np.max(np.max(np.abs(g[:, c] - go), axis=1) / np.max(np.abs(go), axis=1))))
    return dict(logp_rel=worst_lp, grad_rel=worst_g, weights_rel=w_rel, ctx=ctx, model=m, D=D)


def write_uv_model(tmp_path, blue_knot=2300.0, base='T21_model', pop_rv=False):
    """A custom model whose first wavelength knot lies below 2700 A (the FM90 ultraviolet branch of the host-dust law,
    reference get_spectra :627-638): a shipped model with its blue knot moved.  Returns the path of its BAYESN.YAML."""
    import os
    import yaml
    from bayesn_b200 import datafiles
    d = dict(datafiles.load_model_yaml(base))
    d['L_KNOTS'] = [float(blue_knot)] + [float(x) for x in d['L_KNOTS'][1:]]
    if pop_rv:
        d.pop('RV', None)
        d['MUR'], d['SIGMAR'] = 2.8, 0.6
    out = {k: (np.asarray(v).tolist() if not isinstance(v, (int, float, str)) else v) for k, v in d.items()}
    path = os.path.join(str(tmp_path), f'uv{"_pop" if pop_rv else ""}_model.yaml')
    with open(path, 'w') as fh:
        yaml.safe_dump(out, fh)
    return path
