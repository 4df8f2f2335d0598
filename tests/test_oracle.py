"""CPU tests (no GPU): the oracle against what pins it, the host logic, the file readers."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import ref_model, ref_nuts
from oracle.ref_model import RefSED, invKD_irr, spline_coeffs_irr
from tests.common import make_case, ref_model as get_ref

HERE = os.path.dirname(os.path.abspath(__file__))
PS1 = ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1']


def test_trained_model_constants_match_reference_info():
    # bayesn/model_files/T21_model/BAYESN.INFO:27-31
    ref = get_ref('T21_model', PS1)
    assert ref.M0 == -19.5 and abs(ref.sigma0 - 0.103) < 1e-12
    assert abs(ref.RV - 2.61) < 1e-12 and abs(ref.tauA - 0.194) < 1e-12
    assert ref.W0.shape == (6, 6) and ref.L_Sigma.shape == (24, 24)


def test_spline_matrix_is_natural_cubic_spline():
    from scipy.interpolate import CubicSpline
    x = np.array([3000.0, 4300.0, 4900.0, 5400.0, 6200.0, 7700.0, 8700.0, 10400.0, 12400.0, 16500.0, 18500.0])
    y = np.sin(x / 2500.0)
    xi = np.linspace(3000, 18500, 57)
    J = spline_coeffs_irr(xi, x, invKD_irr(x))
    cs = CubicSpline(x, y, bc_type='natural')
    assert np.max(np.abs(J @ y - cs(xi))) < 1e-12
    # linear extrapolation beyond the knots
    xo = np.array([2500.0, 19000.0])
    Jo = spline_coeffs_irr(xo, x, invKD_irr(x))
    d0, d1 = cs(x[0], 1), cs(x[-1], 1)
    assert np.allclose(Jo @ y, [y[0] + d0 * (xo[0] - x[0]), y[-1] + d1 * (xo[1] - x[-1])], atol=1e-12)


def test_time_spline_step_matches_matrix_builder():
    ref = get_ref('T21_model', PS1)
    t = np.array([-14.2, -10.0, -3.3, 0.0, 7.5, 19.99, 39.5, 44.0])
    a = ref.spline_coeffs_irr_step(torch.as_tensor(t), ref.tau_knots, ref.KD_t).numpy()
    b = spline_coeffs_irr(t, ref.tau_knots, ref.KD_t)
    assert np.max(np.abs(a - b)) < 1e-13


def test_ab_zero_points_ps1():
    # values derived independently during the survey of the reference
    ref = get_ref('T21_model', PS1)
    want = {'g_PS1': -20.8411216, 'r_PS1': -21.37027142, 'i_PS1': -21.79336838, 'z_PS1': -22.09908881}
    for k, v in want.items():
        assert abs(ref.zp_dict[k] - v) < 1e-7


def test_product_static_tables_equal_oracle():
    from bayesn_b200 import SEDmodel
    bands = ['u_LSST', 'g_PS1', 'Y_WFCAM', 'H_WFCAM', 'B_CSP']
    ref = get_ref('W22_model', bands)
    m = SEDmodel(load_model='W22_model')
    assert np.max(np.abs(m.J_l_T - ref.J_l_T)) < 1e-13
    assert np.max(np.abs(m.KD_t - ref.KD_t)) < 1e-13
    assert np.max(np.abs(m.M_fitz_block - ref.M_fitz_block)) < 1e-13
    assert np.max(np.abs(m.hsiao_flux - ref.hsiao_flux)) / ref.hsiao_flux.max() < 1e-13
    names = [ref.inv_band_dict[k] for k in range(len(ref.band_dict))]
    bi = np.array([m.band_dict[n] for n in names])
    assert np.max(np.abs(m.zps[bi] - ref.zps)) < 1e-12          # AB, Vega (FITS reader) and BD17 systems
    z = np.array([0.015, 0.05, 0.08])
    ebv = np.array([0.0, 0.03, 0.1])
    w_o = ref._calculate_band_weights(z, ebv)
    w_p = m._calculate_band_weights(z, ebv, bi)
    assert np.max(np.abs(w_o - w_p)) / np.max(np.abs(w_o)) < 1e-13
    assert np.max(np.abs(m.cosmo.distmod(z) - ref.distmod(z))) < 1e-9


def test_distmod_known_value():
    # flat LCDM H0=73.24 Om0=0.28: d_L(z=0.019253) -> mu = 34.5154 (used by the SN 2016W fit)
    assert abs(ref_model.distmod(0.019253)[0] - 34.51542547) < 1e-6


def test_oracle_gradient_finite_difference():
    case = make_case('T21_model', PS1, np.arange(-8, 40, 4.0), N=2, seed=5)
    ref = case['ref']
    rng = np.random.RandomState(0)
    q = np.zeros((2, 28))
    q[:, 0] = rng.randn(2)
    q[:, 1] = np.log(0.2)
    q[:, 2] = 0.1
    q[:, 3] = case['obs'][7, 0, :] + 0.05
    q[:, 4:] = rng.randn(2, 24)
    lp, g = ref.fit_logp_grad(q, case['obs'], case['weights'])
    for i in [0, 1, 2, 3, 4, 17, 27]:
        h = 1e-6
        qp, qm = q.copy(), q.copy()
        qp[:, i] += h
        qm[:, i] -= h
        fd = (ref.fit_logp(torch.as_tensor(qp), case['obs'], case['weights'])
              - ref.fit_logp(torch.as_tensor(qm), case['obs'], case['weights'])).numpy() / (2 * h)
        assert np.allclose(fd, g[:, i], rtol=2e-5, atol=1e-4)


@pytest.mark.parametrize('name', ['T21_griz', 'W22_ugrizYJH', 'M20_csp', 'popRV_griz'])
def test_oracle_reproduces_committed_golden(name):
    d = np.load(os.path.join(HERE, 'golden', f'logp_{name}.npz'))
    model = {'T21_griz': 'T21_model', 'W22_ugrizYJH': 'W22_model', 'M20_csp': 'M20_model',
             'popRV_griz': 'pop_RV_model'}[name]
    ref = get_ref(model, [b for b in d['band_names'] if b != 'NULL_BAND'])
    mode = 'pop' if name.startswith('popRV') else 'fixed'
    # the band weights are recomputed from the filter files: pins the static pipeline too
    w = ref._calculate_band_weights(d['z'], d['ebv'])
    assert np.max(np.abs(w - d['weights'])) / np.max(np.abs(w)) < 1e-12
    lp, g = ref.fit_logp_grad(d['q'].astype(np.float64), d['obs'], w, rv_mode=mode)
    assert np.allclose(lp, d['logp'], rtol=1e-10)
    assert np.allclose(g, d['grad'], rtol=1e-8, atol=1e-8)


def test_sn2016w_forward_tracks_data_at_notebook_posterior():
    from bayesn_b200 import snana, datafiles
    ref = get_ref('T21_model', PS1)
    meta, lc = snana.read_snana_ascii(datafiles.EXAMPLE_LC)
    assert meta['REDSHIFT_HELIO'] == 0.019253 and meta['MWEBV'] == 0.0599 and len(lc['MJD']) == 29
    fm = {'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'}
    data, bw = ref.prepare_fit_data(lc['MJD'], lc['FLUXCAL'], lc['FLUXCALERR'], lc['FLT'], meta['REDSHIFT_HELIO'],
                                    ebv_mw=meta['MWEBV'], peak_mjd=meta['SEARCH_PEAKMJD'], filt_map=fm)
    assert data.shape == (10, 28, 1)      # one epoch falls outside (-10, 40) d
    gold = json.load(open(os.path.join(HERE, 'golden', 'sn2016w_notebook.json')))['globalRV']
    q = np.zeros((1, 28))
    q[0, 0] = gold['theta'][0]
    q[0, 1] = math.log(gold['AV'][0])
    p = (gold['tmax'][0] + 10) / 20
    q[0, 2] = math.log(p / (1 - p))
    q[0, 3] = gold['Ds'][0]
    q[0, 4:] = [e[0] for e in gold['eps_tform']]
    lp, flux = ref.fit_logp(torch.as_tensor(q), data, bw, return_flux=True)
    ratio = flux.numpy()[:, 0] / data[1, :, 0]
    assert abs(np.median(ratio) - 1) < 0.02
    chi2 = np.sum(((flux.numpy()[:, 0] - data[1, :, 0]) / data[2, :, 0]) ** 2)
    assert chi2 < 3 * 28        # the notebook's posterior mean fits the light curve


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10
    assert ref_nuts.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert ref_nuts.philox4x32((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6,
                                                                              0x6d5451fd)


def test_adaptation_windows_match_stan_schedule():
    assert ref_nuts.adaptation_window_ends(250) == [74, 99, 199, 249]
    assert ref_nuts.adaptation_window_ends(500) == [74, 99, 149, 249, 449, 499]
    assert ref_nuts.adaptation_window_ends(10) == [9]


def test_oracle_nuts_on_gaussian():
    rng = np.random.RandomState(1)
    D = 5
    sd = np.array([0.1, 1.0, 3.0, 0.5, 10.0])
    mu = np.array([1.0, -2.0, 0.0, 5.0, 30.0])

    def lg(q, rows):
        d = (q - mu) / sd
        return -0.5 * np.sum(d * d, axis=1), -d / sd

    q0 = mu + sd * rng.randn(4, D)
    recs, ticks, evals = ref_nuts.run_chains(lg, q0, 300, 400, seed=3)
    z = np.concatenate([np.array(r['z'][300:]) for r in recs])
    assert np.all(np.abs(z.mean(0) - mu) < 0.15 * sd)
    assert np.all(np.abs(z.std(0) / sd - 1) < 0.12)
    assert sum(sum(r['diverging'][300:]) for r in recs) == 0


def test_hsiao_and_fits_readers():
    from bayesn_b200 import datafiles
    ph, wv, fl = datafiles.load_hsiao()
    assert ph[0] == -19 and ph[-1] == 85 and wv[0] == 1000 and wv[-1] == 25000 and fl.shape == (105, 2401)
    assert ph[19] == 0.0
    tab = datafiles.read_fits_bintable(os.path.join(datafiles.FILTER_DIR, 'standards', 'alpha_lyr_stis_010.fits'))
    assert tab['WAVELENGTH'].shape == (9192,) and np.all(np.diff(tab['WAVELENGTH']) > 0)
    assert 3e-9 < tab['FLUX'][np.argmin(np.abs(tab['WAVELENGTH'] - 5556))] < 4e-9   # Vega at 5556 A


def test_snana_roundtrip(tmp_path):
    """write_snana_lcfile keeps the reference's layout (bayesn/bayesn_io.py:13-132): header keys and their order, six
    columns, default file name, returned value, END: terminator - and the reader takes it back."""
    from bayesn_b200 import snana
    name = snana.write_snana_lcfile(str(tmp_path), 'sn1', [1.0, 2.5], ['r_PS1', 'g_PS1'], [17.1, 17.3], [0.02, 0.03],
                                    0.5, 0.03, 0.031, 1e-4, 0.02, ra=10.5, survey='PS1MD', paper='X20')
    assert name == 'sn1_PS1MD_X20.snana.dat'
    text = open(tmp_path / name).read()
    lines = text.split('\n')
    assert lines[0] == '# sn1 ' and lines[1] == '# SNANA-like file generated from user-provided data'
    keys = [l.split(':')[0] for l in lines if l and not l.startswith('#') and not l.startswith('OBS')]
    assert keys == ['SNID', 'SOURCE', 'RA', 'MWEBV', 'REDSHIFT_HELIO', 'REDSHIFT_CMB', 'REDSHIFT_CMB_ERR', 'PEAKMJD',
                    'FILTERS', 'NOBS', 'NVAR', 'VARLIST', 'END']
    assert 'SOURCE: PS1MD' in lines and 'FILTERS: g_PS1,r_PS1' in lines and 'NVAR: 6' in lines
    assert 'VARLIST: MJD FLT FLUXCAL FLUXCALERR MAG MAGERR' in lines
    assert 'OBS: 1.0 r_PS1 14454.3977 266.2598 17.1 0.02' in lines and text.endswith('\nEND:')
    meta, lc = snana.read_snana_ascii(str(tmp_path / name))
    assert meta['REDSHIFT_HELIO'] == 0.03 and meta['MWEBV'] == 0.02 and meta['PEAKMJD'] == 0.5 and meta['SNID'] == 'sn1'
    assert list(lc['FLT']) == ['r_PS1', 'g_PS1'] and np.allclose(lc['MAG'], [17.1, 17.3])
    assert np.allclose(lc['FLUXCAL'], 10 ** ((27.5 - np.array([17.1, 17.3])) / 2.5), atol=1e-4)
    assert snana.write_snana_lcfile(str(tmp_path), 'sn2', [1.0], ['g_PS1'], [17.0], [0.1], 0, 0.03, 0.03, 1e-4, 0.0) \
        == 'sn2.snana.dat'
    with pytest.raises(ValueError, match='does not exist'):
        snana.write_snana_lcfile(str(tmp_path / 'nope'), 'sn1', [1.0], ['g_PS1'], [17.0], [0.1], 0, 0.03, 0.03, 1e-4, 0.0)
    with pytest.raises(ValueError, match='same length'):
        snana.write_snana_lcfile(str(tmp_path), 'sn1', [1.0, 2.0], ['g_PS1'], [17.0], [0.1], 0, 0.03, 0.03, 1e-4, 0.0)


def test_summary_statistics():
    from bayesn_b200 import summary
    rng = np.random.RandomState(0)
    x = rng.randn(4, 2000)
    assert abs(summary.split_rhat(x) - 1) < 0.01
    assert 6000 < summary.effective_sample_size(x) < 10000
    ar = np.zeros((4, 2000))
    for i in range(1, 2000):
        ar[:, i] = 0.9 * ar[:, i - 1] + rng.randn(4)
    ess = summary.effective_sample_size(ar)
    assert 250 < ess < 700          # theory: 8000 (1-0.9)/(1+0.9) = 421
    tab = summary.summary_table({'a': x, 'b': np.stack([x, ar], axis=-1)}, device='cpu')
    assert list(tab.index) == ['a', 'b[0]', 'b[1]']
    assert list(tab.columns) == ['mean', 'sd', 'hdi_3%', 'hdi_97%', 'mcse_mean', 'mcse_sd', 'ess_bulk', 'ess_tail', 'r_hat']


def test_arviz_style_diagnostics_batched_against_per_site():
    """fit_summary.csv / FITRES r-hat: the rank-normalised diagnostics of arviz.summary (Vehtari et al. 2021).  The
    batched torch path against the per-site numpy statement on sites with autocorrelation, ties (stuck transitions),
    a shifted chain, a chain of a different scale; and the properties the statistics are built for."""
    import torch
    from scipy.stats import rankdata
    from bayesn_b200 import summary
    rng = np.random.RandomState(3)
    C, S, B = 4, 250, 120
    x = rng.randn(C, S, B)
    for i in range(1, S):
        x[:, i, :40] = 0.8 * x[:, i - 1, :40] + x[:, i, :40]
    x[:, :, 40:60] = np.round(x[:, :, 40:60], 1)
    x[:, :, 60:70] = np.repeat(x[:, ::5, 60:70], 5, axis=1)
    x[0, :, 70:80] += 1.5
    x[1, :, 80:90] *= 3.0
    a = summary.summary_table({'s': x}, device='cpu')
    b = summary.summary_table_per_site({'s': x})
    assert list(a.index) == list(b.index) and list(a.columns) == list(b.columns)
    for c in a.columns:
        assert np.allclose(a[c].values, b[c].values, rtol=1e-9, atol=1e-12), c
    t = np.round(rng.randn(500), 1)
    assert np.array_equal(summary._rank_average(t), rankdata(t))
    xt = torch.from_numpy(np.ascontiguousarray(np.moveaxis(x, 2, 0)))
    assert np.allclose(summary.rank_rhat_t(xt).numpy(), a['r_hat'].values, rtol=1e-12)
    iid = a['r_hat'].values[90:]
    assert np.all(np.abs(iid - 1) < 0.02)
    assert np.all(a['r_hat'].values[70:80] > 1.1)                       # shifted chain: bulk statistic
    classic = np.array([summary.split_rhat(x[:, :, i]) for i in range(80, 90)])
    assert np.all(a['r_hat'].values[80:90] > 1.05) and np.all(classic < 1.02)       # scale: only the folded statistic sees it
    assert np.all(a['ess_bulk'].values[:40] < 400) and np.all(a['ess_bulk'].values[90:] > 600)
    assert np.allclose(a['mcse_mean'].values[90:], a['sd'].values[90:] / np.sqrt(1000), rtol=0.25)


def test_training_oracle_gradient_finite_difference_and_golden():
    """train_model_globalRV (reference :1115-1182) in the unconstrained space: the autograd gradient
    of the restatement against central differences, and the committed fixture."""
    import torch
    from tests.common import ref_model
    d = np.load(os.path.join(HERE, 'golden', 'train_T21_griz.npz'))
    ref = ref_model('T21_model', [str(b) for b in d['band_names'][1:]])
    qg, ql = d['qg'].astype(np.float64), d['ql'].astype(np.float64)
    lp, gg, gl = ref.train_logp_grad(qg, ql, d['obs'], d['weights'])
    assert abs(lp - float(d['logp'])) < 1e-9 * abs(lp)
    assert np.allclose(gg, d['grad_g'], rtol=1e-9, atol=1e-9) and np.allclose(gl, d['grad_l'], rtol=1e-9, atol=1e-9)

    def f(a, b):
        with torch.no_grad():
            return float(ref.train_logp(torch.as_tensor(a), torch.as_tensor(b), d['obs'], d['weights']))
    dims = ref.train_dims()
    G = dims['G']
    for i in [0, 40, 75, 100, 140, 300, G - 3, G - 2, G - 1]:       # W0, W1, sigma_eps, L_Omega, sigma0, RV, tauA
        h = 1e-6
        a, b = qg.copy(), qg.copy()
        a[i] += h
        b[i] -= h
        fd = (f(a, ql) - f(b, ql)) / (2 * h)
        assert abs(fd - gg[i]) < 1e-5 * max(1.0, abs(gg[i])), (i, fd, gg[i])
    for (s, j) in [(0, 0), (2, dims['n_eps']), (3, dims['n_eps'] + 1), (5, dims['n_eps'] + 2)]:
        h = 1e-6
        a, b = ql.copy(), ql.copy()
        a[s, j] += h
        b[s, j] -= h
        fd = (f(qg, a) - f(qg, b)) / (2 * h)
        assert abs(fd - gl[s, j]) < 1e-5 * max(1.0, abs(gl[s, j])), (s, j, fd, gl[s, j])


def test_corr_cholesky_is_a_correlation_factor_with_numpyro_logdet():
    """CorrCholeskyTransform restatement: unit-norm rows, lower triangular, positive diagonal, and
    its log|det J| equals the log-determinant of the Jacobian d(strict lower triangle)/dy."""
    import torch
    from oracle.ref_model import RefSED
    n = 5
    y = torch.tensor(np.random.RandomState(1).randn(n * (n - 1) // 2) * 0.6, dtype=torch.float64)
    L, ld = RefSED.corr_cholesky(y, n)
    assert torch.allclose((L ** 2).sum(dim=1), torch.ones(n, dtype=torch.float64))
    assert torch.all(torch.diagonal(L) > 0) and torch.allclose(torch.triu(L, 1), torch.zeros(n, n, dtype=torch.float64))
    ii, jj = torch.tril_indices(n, n, -1)
    J = torch.autograd.functional.jacobian(lambda v: RefSED.corr_cholesky(v, n)[0][ii, jj], y)
    assert abs(float(torch.linalg.slogdet(J)[1]) - float(ld)) < 1e-10


def test_reference_module_names_are_importable():
    """``bayesn.spline_utils`` / ``bayesn.bayesn_io`` names of the reference (spline_utils.py:9-180, bayesn_io.py:13)
    exist under ``bayesn_b200`` and agree with the oracle's restatement."""
    from bayesn_b200.spline_utils import invKD_irr, spline_coeffs_irr, cartesian_prod, spline_coeffs_irr_step
    from bayesn_b200.bayesn_io import write_snana_lcfile
    from oracle.ref_model import invKD_irr as o_inv, spline_coeffs_irr as o_sp
    x = np.array([0.0, 1.0, 2.5, 4.0, 7.0])
    xi = np.linspace(-1, 8, 41)
    assert np.allclose(invKD_irr(x), o_inv(x), atol=1e-14)
    J = spline_coeffs_irr(xi, x, invKD_irr(x))
    assert np.allclose(J, o_sp(xi, x, o_inv(x)), atol=1e-13)
    assert np.allclose(spline_coeffs_irr_step(xi[7], x, invKD_irr(x)), J[7])
    assert cartesian_prod([1, 2], [3, 4, 5]).tolist() == [[1, 3], [1, 4], [1, 5], [2, 3], [2, 4], [2, 5]]
    with pytest.raises(ValueError, match='out of bounds'):
        spline_coeffs_irr(xi, x, invKD_irr(x), allow_extrap=False)
    assert callable(write_snana_lcfile)
