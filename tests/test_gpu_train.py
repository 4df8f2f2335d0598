"""GPU parity of the population-training path (train_model_globalRV, reference
bayesn/bayesn_model.py:1115-1182): log joint density and gradient of the CUDA kernels, called
through the C ABI, against the float64 oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): log p relative 1e-5, gradient 1e-4 norm-wise."""
import os

import numpy as np
import pytest
import torch

from tests.common import make_case, product_model

pytestmark = pytest.mark.gpu
LOGP_TOL, GRAD_TOL = 1e-5, 1e-4
CSP = ['B_CSP', 'V_CSP', 'r_CSP', 'i_CSP', 'Y_RC', 'J_RC1', 'H_RC']
PS1 = ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1']


def training_point(ref, case, C, seed, spread=1.0):
    """Unconstrained points around the trained model: globals (C, G) and latents (N, C, Dl)."""
    d = ref.train_dims()
    L, T, ne, nc = d['L'], d['T'], d['n_eps'], d['n_corr']
    N = case['obs'].shape[2]
    rng = np.random.RandomState(seed)
    qg = np.zeros((C, d['G']))
    qg[:, :L * T] = ref.W0.flatten(order='F')[None] + 0.02 * spread * rng.randn(C, L * T)
    qg[:, L * T:2 * L * T] = ref.W1.flatten(order='F')[None] + 0.02 * spread * rng.randn(C, L * T)
    se = 0.1 + 0.02 * rng.rand(C, ne)
    x = np.arctan(se) / (np.pi / 2)
    qg[:, 2 * L * T:2 * L * T + ne] = np.log(x) - np.log1p(-x)
    qg[:, 2 * L * T + ne:2 * L * T + ne + nc] = 0.15 * spread * rng.randn(C, nc)
    qg[:, -3] = 0.1 * rng.randn(C)
    qg[:, -2] = 0.3 * rng.randn(C)
    qg[:, -1] = -1.5 + 0.2 * rng.randn(C)
    ql = np.zeros((N, C, d['Dl']))
    ql[..., :ne] = 0.7 * rng.randn(N, C, ne)
    ql[..., ne] = case['params']['theta'][:, None] + 0.2 * rng.randn(N, C)
    ql[..., ne + 1] = np.log(case['params']['AV'][:, None] + 0.05) + 0.1 * rng.randn(N, C)
    ql[..., ne + 2] = (case['params']['mu'] + case['params']['del_M'])[:, None] + 0.03 * rng.randn(N, C)
    return qg, ql


def run_case(model, bands, t, N, C, seed, mask_some=True):
    case = make_case(model, bands, t, N=N, seed=seed, mag=True, mask_some=mask_some, yerr=0.05)
    ref = case['ref']
    m, bi = product_model(case)
    ctx = m.prepare_training(case['obs'], m._calculate_band_weights(case['z'], case['ebv'], bi), bi)
    qg, ql = training_point(ref, case, C, seed)
    qg32 = torch.tensor(qg, dtype=torch.float32, device='cuda').contiguous()
    ql32 = torch.tensor(ql, dtype=torch.float32, device='cuda').contiguous()
    lp, gg, gl = ctx.train_logp_grad(qg32, ql32)
    torch.cuda.synchronize()
    lp, gg, gl = lp.cpu().numpy(), gg.cpu().numpy().astype(np.float64), gl.cpu().numpy().astype(np.float64)
    qgk, qlk = qg32.cpu().numpy().astype(np.float64), ql32.cpu().numpy().astype(np.float64)
    out = []
    for c in range(C):
        lpo, ggo, glo = ref.train_logp_grad(qgk[c], qlk[:, c], case['obs'], case['weights'])
        assert np.isfinite(lpo)
        e_lp = abs(lp[c] - lpo) / abs(lpo)
        e_gg = np.max(np.abs(gg[c] - ggo)) / np.max(np.abs(ggo))
        e_gl = np.max(np.max(np.abs(gl[:, c] - glo), axis=1) / np.max(np.abs(glo), axis=1))
        # the blocks of the global gradient separately: they differ by orders of magnitude
        d = ref.train_dims()
        LT, ne = d['L'] * d['T'], d['n_eps']
        blocks = {'W0': slice(0, LT), 'W1': slice(LT, 2 * LT), 'sigeps': slice(2 * LT, 2 * LT + ne),
                  'L_Omega': slice(2 * LT + ne, d['G'] - 3), 'scalars': slice(d['G'] - 3, d['G'])}
        eb = {k: float(np.max(np.abs(gg[c][sl] - ggo[sl])) / max(np.max(np.abs(ggo[sl])), 1e-30)) for k, sl in blocks.items()}
        out.append((e_lp, e_gg, e_gl, eb))
    return out


@pytest.mark.parametrize('name,model,bands,t,N,C', [
    ('C4_M20_csp', 'M20_model', CSP, np.arange(-8, 40, 4.0), 12, 2),
    ('T21_griz_1chain', 'T21_model', PS1, np.arange(-8, 40, 4.0), 9, 1),
    ('W22_4chains', 'W22_model', ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1', 'Y_WFCAM', 'J_WFCAM', 'H_WFCAM'],
     np.array([-8.0, 2.0, 12.0, 22.0, 32.0]), 7, 4),
])
def test_train_logp_grad_parity(name, model, bands, t, N, C):
    for e_lp, e_gg, e_gl, eb in run_case(model, bands, t, N, C, seed=len(name)):
        assert e_lp < LOGP_TOL, (name, e_lp)
        assert e_gl < GRAD_TOL, (name, e_gl)
        assert e_gg < GRAD_TOL, (name, e_gg, eb)
        for k, v in eb.items():
            assert v < 5e-4, (name, k, v)


def test_train_reduction_is_additive_over_sn_shards():
    """The per-SN sums of two disjoint shards add up to the sums of the whole set (the property the
    multi-GPU all-reduce relies on)."""
    case = make_case('T21_model', PS1, np.arange(-8, 40, 4.0), N=8, seed=3, mag=True, mask_some=True)
    ref = case['ref']
    m, bi = product_model(case)
    w = m._calculate_band_weights(case['z'], case['ebv'], bi)
    qg, ql = training_point(ref, case, 2, 5)
    qg32 = torch.tensor(qg, dtype=torch.float32, device='cuda').contiguous()
    full = m.prepare_training(case['obs'], w, bi)
    bufs = {}
    lp, gg, gl = full.train_logp_grad(qg32, torch.tensor(ql, dtype=torch.float32, device='cuda').contiguous(), out=bufs)
    red_full = bufs['red'].clone()
    reds = []
    for sl in (slice(0, 3), slice(3, 8)):
        ctx = m.prepare_training(case['obs'][:, :, sl], w[sl], bi)
        b = {}
        ctx.train_logp_grad(qg32, torch.tensor(ql[sl], dtype=torch.float32, device='cuda').contiguous(), out=b)
        reds.append(b['red'].clone())
    torch.cuda.synchronize()
    tot = reds[0] + reds[1]
    assert torch.allclose(tot, red_full, rtol=1e-9, atol=1e-9 * float(red_full.abs().max()))


def _small_training_problem(N=3, C=2, seed=7):
    case = make_case('T21_model', PS1, np.arange(-8, 40, 4.0), N=N, seed=seed, mag=True, mask_some=True)
    ref = case['ref']
    m, bi = product_model(case)
    ctx = m.prepare_training(case['obs'], m._calculate_band_weights(case['z'], case['ebv'], bi), bi)
    qg, ql = training_point(ref, case, C, seed, spread=0.3)
    qg32 = torch.tensor(qg, dtype=torch.float32, device='cuda').contiguous()
    ql32 = torch.tensor(ql, dtype=torch.float32, device='cuda').contiguous()
    return case, ref, m, ctx, qg32, ql32


def test_train_nuts_trajectories_match_oracle_sampler():
    """Same Philox streams, same start, adaptation off: the sharded device sampler and the float64
    oracle sampler build the same trees (leapfrog counts per transition) and land on the same points
    for the first transitions (float32 chaos separates them eventually)."""
    from oracle import ref_fit
    case, ref, m, ctx, qg32, ql32 = _small_training_problem()
    n_tr = 5
    out = ctx.train_nuts(qg32, ql32, num_warmup=0, num_samples=n_tr, max_tree_depth=5, init_step_size=0.004,
                         adapt_step_size=False, adapt_mass_matrix=False, seed=11, passes_per_graph=8)
    assert out['finished']
    st = out['stats'].cpu().numpy()
    sg = out['samples_g'].cpu().numpy().astype(np.float64)
    sl = out['samples_l'].cpu().numpy().astype(np.float64)
    qg, ql = qg32.cpu().numpy().astype(np.float64), ql32.cpu().numpy().astype(np.float64)
    G = qg.shape[1]
    for c in range(2):
        rec = ref_fit.train_cpu(ref, case['obs'], case['weights'], qg[c], ql[:, c], chain=c, n_chains=2, num_samples=n_tr,
                                max_depth=5, init_step=0.004, seed=11)
        n_gpu, n_cpu = st[c, :, 1], np.array(rec['n_steps'])
        agree = int(np.argmax(n_gpu != n_cpu)) if np.any(n_gpu != n_cpu) else n_tr
        assert agree >= 3, (n_gpu, n_cpu)
        z_cpu = np.array(rec['z'])
        assert np.max(np.abs(sg[c, :agree] - z_cpu[:agree, :G])) < 5e-3
        zl = z_cpu[:agree, G:].reshape(agree, ql.shape[0], -1)
        assert np.max(np.abs(sl[:, c, :agree].transpose(1, 0, 2) - zl)) < 5e-3
        assert np.allclose(st[c, :agree, 0], np.array(rec['accept'])[:agree], atol=5e-3)
        assert np.allclose(st[c, :agree, 3], np.array(rec['pe'])[:agree], rtol=1e-5)


def test_train_nuts_adapts_and_conserves_energy():
    """Short adaptive run on a small simulated training set: finite draws, acceptance near the
    target after warm-up, few divergences, and CUDA-graph replay == eager launches."""
    case, ref, m, ctx, qg32, ql32 = _small_training_problem(N=16, C=2, seed=4)
    kw = dict(num_warmup=60, num_samples=30, max_tree_depth=6, seed=3)
    out = ctx.train_nuts(qg32, ql32, passes_per_graph=16, **kw)
    assert out['finished']
    for k in ('samples_g', 'samples_l', 'stats'):
        assert torch.isfinite(out[k]).all(), k
    st = out['stats'].cpu().numpy()
    assert 0.5 < st[:, 60:, 0].mean() < 0.99
    assert st[:, 60:, 2].sum() <= 3
    out2 = ctx.train_nuts(qg32, ql32, use_graph=False, passes_per_graph=16, **kw)
    assert torch.equal(out['samples_g'], out2['samples_g']) and torch.equal(out['samples_l'], out2['samples_l'])


def test_train_kernel_against_committed_golden():
    """Same inputs as tests/test_oracle.py::test_training_oracle_gradient_finite_difference_and_golden."""
    import os
    from bayesn_b200 import SEDmodel
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'train_T21_griz.npz'))
    m = SEDmodel(load_model='T21_model')
    bi = np.array([m.band_dict[str(n)] for n in d['band_names']])
    ctx = m.prepare_training(d['obs'], None, bi)              # band weights rebuilt on the device from z, E(B-V)
    qg = torch.tensor(d['qg'][None], dtype=torch.float32, device='cuda').contiguous()
    ql = torch.tensor(d['ql'][:, None], dtype=torch.float32, device='cuda').contiguous()
    lp, gg, gl = ctx.train_logp_grad(qg, ql)
    torch.cuda.synchronize()
    assert abs(float(lp[0]) - float(d['logp'])) / abs(float(d['logp'])) < LOGP_TOL
    gg, gl = gg[0].cpu().numpy().astype(np.float64), gl[:, 0].cpu().numpy().astype(np.float64)
    assert np.max(np.abs(gg - d['grad_g'])) / np.max(np.abs(d['grad_g'])) < GRAD_TOL
    assert np.max(np.max(np.abs(gl - d['grad_l']), axis=1) / np.max(np.abs(d['grad_l']), axis=1)) < GRAD_TOL


def test_training_recovers_the_generating_model():
    """End to end: SNe simulated from T21 (simulate_light_curve, mag=True), SEDmodel.train with the
    device-resident sampler (4 chains), posterior of the global parameters against the truth."""
    from bayesn_b200 import SEDmodel, train as btrain, summary
    m = SEDmodel(load_model='T21_model')
    obs, w, bi = btrain.simulate_training_set(m, 100, seed=11, bands=('g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'),
                                              t=np.arange(-8, 40, 3.0), yerr=0.02)
    W, S = 150, 100
    samples, ex = m.train(obs, None, bi, num_chains=4, num_warmup=W, num_samples=S, initialisation='T21_model', seed=5)
    st = ex['stats']
    # a short warm-up (150) occasionally leaves one chain with a step size that is a little too large
    assert ex['info'][:, 5].sum() <= 0.1 * 4 * S              # divergences
    assert 0.6 < st[:, W:, 0].mean() < 0.98
    for k, truth in (('tauA', m.tauA), ('sigma0', m.sigma0), ('RV', m.RV)):
        x = samples[k]
        assert abs(x.mean() - truth) < 4 * x.std() + 1e-3, (k, x.mean(), x.std(), truth)
        assert summary.split_rhat(x) < 1.3, k
    W0 = samples['W0']
    z = (W0.mean(axis=(0, 1)) - m.W0.flatten(order='F')) / W0.std(axis=(0, 1))
    assert np.sqrt((z ** 2).mean()) < 2.5 and np.abs(z).max() < 5, z
    assert samples['theta'].shape == (4, S, 100) and samples['eps'].shape == (4, S, m.n_eps, 100)
    assert samples['L_Omega'].shape == (4, S, m.n_eps, m.n_eps)


@pytest.mark.parametrize('K', [2, 4, 8])
def test_train_sn_kernel_latency_mode_matches_one_warp_per_unit(K, monkeypatch):
    """Latency mode of the training SN kernel (K warps share the observation pairs of a unit's evaluation; used
    when a rank's SNe leave warp slots idle): same trees as one warp per unit for the first transitions, finite
    draws, and the whole run agrees statistically (acceptance)."""
    case, ref, m, ctx, qg32, ql32 = _small_training_problem(N=5, C=2, seed=9)
    kw = dict(num_warmup=0, num_samples=6, max_tree_depth=5, init_step_size=0.004, adapt_step_size=False,
              adapt_mass_matrix=False, seed=11, passes_per_graph=8)
    monkeypatch.setenv('BAYESN_TRAIN_KW', '1')
    o1 = ctx.train_nuts(qg32, ql32, **kw)
    assert o1['warps_per_unit'] == 1
    monkeypatch.setenv('BAYESN_TRAIN_KW', str(K))
    ok = ctx.train_nuts(qg32, ql32, **kw)
    assert ok['warps_per_unit'] == K and ok['finished']
    for k in ('samples_g', 'samples_l', 'stats'):
        assert torch.isfinite(ok[k]).all(), k
    s1, sk = o1['stats'].cpu().numpy(), ok['stats'].cpu().numpy()
    assert np.array_equal(s1[:, :3, 1], sk[:, :3, 1]), (s1[:, :, 1], sk[:, :, 1])       # leapfrogs per transition
    assert np.allclose(s1[:, :3, 3], sk[:, :3, 3], rtol=1e-5)                            # potential energy
    assert float((o1['samples_g'][:, :3] - ok['samples_g'][:, :3]).abs().max()) < 5e-3


@pytest.mark.parametrize('init', ['T21_model', 'median'])
def test_run_training_from_snana_text_directory(tmp_path, init):
    """End to end like ``run_bayesn input.yaml`` with ``mode: training_globalRV`` (reference :1766-1770, :1913-1924,
    postprocess :2186-2242): SNANA text files with magnitudes (written by simulate_light_curve(write_to_files=True))
    -> process_dataset -> device-resident training sampler -> bayesn.yaml / chains.pkl / initial_chains.pkl /
    input.yaml; every chain starts from ``initial_guess`` of a model or from its own prior median."""
    import pickle
    import yaml
    from bayesn_b200 import SEDmodel
    m = SEDmodel(load_model='T21_model')
    np.random.seed(5)
    N = 12
    z = np.random.uniform(0.02, 0.06, N)
    d = tmp_path / 'SIMTRAIN'
    m.simulate_light_curve(np.arange(-8, 36, 4.0), N, PS1, yerr=0.02, err_type='mag', z=z, mu='z', ebv_mw=0.01, mag=True,
                           write_to_files=True, output_dir=str(d))
    files = sorted(os.listdir(d))
    assert len(files) == N
    (d / 'SIMTRAIN.LIST').write_text('\n'.join(files) + '\n')
    out = tmp_path / 'out'
    W, S = 40, 20
    samples = m.run(dict(mode='training_globalRV', version_photometry=str(d), outputdir=str(out), num_warmup=W, num_samples=S,
                         num_chains=2, initialisation=init))
    L, T = len(m.l_knots), len(m.tau_knots)
    assert samples['W0'].shape == (2, S, L * T) and samples['theta'].shape == (2, S, N) and samples['AV'].shape == (2, S, N)
    assert np.all(np.isfinite(samples['W0'])) and np.all(samples['tauA'] > 0) and np.all(samples['sigma0'] > 0)
    y = yaml.safe_load(open(out / 'bayesn.yaml'))
    assert np.array(y['W0']).shape == (L, T) and np.array(y['L_SIGMA_EPSILON']).shape == ((L - 2) * T, (L - 2) * T)
    assert 'RV' not in y                                   # reference quirk (:2235): globalRV runs do not write RV
    ch = pickle.load(open(out / 'chains.pkl', 'rb'))
    assert ch['W1'].shape == (2, S, L * T)
    assert (out / 'initial_chains.pkl').exists() and (out / 'input.yaml').exists() and (out / 'fit_summary.csv').exists()


@pytest.mark.gpu
def test_run_training_on_new_knots(tmp_path):
    """A training run that defines its model on its own knots (``l_knots`` / ``tau_knots`` of the input, reference
    :1721-1727): seven wavelength knots instead of T21's six - another kernel instantiation (L <= 9) - started from
    T21 interpolated onto them; bayesn.yaml comes back on the new knots and loads as a model."""
    import yaml
    from bayesn_b200 import SEDmodel
    m = SEDmodel(load_model='T21_model')
    np.random.seed(6)
    N = 10
    z = np.random.uniform(0.02, 0.06, N)
    d = tmp_path / 'SIMTRAIN'
    m.simulate_light_curve(np.arange(-8, 36, 4.0), N, PS1, yerr=0.02, err_type='mag', z=z, mu='z', ebv_mw=0.01, mag=True,
                           write_to_files=True, output_dir=str(d))
    (d / 'SIMTRAIN.LIST').write_text('\n'.join(sorted(os.listdir(d))) + '\n')
    lk = [3500.0, 4500.0, 5500.0, 6500.0, 7700.0, 8700.0, 9500.0]
    out = tmp_path / 'out'
    W, S = 30, 15
    samples = m.run(dict(mode='training_globalRV', version_photometry=str(d), outputdir=str(out), num_warmup=W, num_samples=S,
                         num_chains=2, initialisation='T21_model', l_knots=lk))
    assert list(m.l_knots) == lk and samples['W0'].shape == (2, S, 42) and samples['sigmaepsilon'].shape == (2, S, 30)
    assert np.all(np.isfinite(samples['W0'])) and np.all(np.isfinite(samples['theta']))
    y = yaml.safe_load(open(out / 'bayesn.yaml'))
    assert y['L_KNOTS'] == lk and np.array(y['W0']).shape == (7, 6) and np.array(y['L_SIGMA_EPSILON']).shape == (30, 30)
    assert np.max(np.abs(np.array(y['W0']))) < 20 and (out / 'fit_summary.csv').exists()
    y['RV'] = 2.6                                      # reference quirk: globalRV runs do not write RV (:2235)
    path = tmp_path / 'trained.yaml'
    path.write_text(yaml.safe_dump(y))
    m2 = SEDmodel(load_model=str(path))
    assert m2.W0.shape == (7, 6) and m2.n_eps == 30
