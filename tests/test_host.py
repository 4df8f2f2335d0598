"""CPU tests of the boundary: the C-ABI library loads and exports every declared symbol, the
product path fails loudly without a GPU, dataset packing, multi-rank sharding over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_in_header(built_library):
    hdr = open(os.path.join(ROOT, 'include', 'bayesn_b200.h')).read()
    declared = set(re.findall(r'\b(bayesn_[a-z_0-9]+)\s*\(', hdr))
    declared -= {'bayesn_ctx'}
    assert {'bayesn_logp_grad_fit', 'bayesn_nuts_fit', 'bayesn_flux_batch', 'bayesn_set_model'} <= declared
    for name in declared:
        assert hasattr(built_library, name), name
    assert built_library.bayesn_abi_version() == 2
    from bayesn_b200 import native
    assert set(native.EXPORTS) == declared


def test_library_is_sm100a_with_tma_and_no_tensor_fallback(built_library):
    from bayesn_b200 import native
    out = subprocess.run(['cuobjdump', '-lelf', native.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    # the sampler kernel stages its tables with TMA bulk copies completed on an mbarrier (SASS: UBLKCP + SYNCS),
    # and nothing in the library falls back to tensor-core MMA paths (the contractions are K <= 12: FP32 SIMT)
    fns = re.findall(r'Function : (\S+)', subprocess.run(['cuobjdump', '-sass', native.LIB_PATH], capture_output=True,
                                                           text=True).stdout)
    fn = [f for f in fns if 'nuts_kernelILi11ELi6ELi2ELb0ELb1' in f][0]
    sass = subprocess.run(['cuobjdump', '-sass', '-fun', fn, native.LIB_PATH], capture_output=True, text=True).stdout
    assert 'UBLKCP' in sass and 'SYNCS' in sass
    assert 'BAR.SYNC' in sass or 'BAR.SYNC.DEFER_BLOCKING' in sass          # team barrier of the latency mode
    assert not re.search(r'\b(HMMA|IMMA|UTCHMMA|UTCIMMA|WGMMA)\b', sass)


def _sass_loops(path, mangled_substr):
    """(start, end, [instructions]) of every backward-branch loop of the first kernel whose mangled name
    contains ``mangled_substr``."""
    names = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    fn = [m for m in re.findall(r'Function : (\S+)', names) if mangled_substr in m][0]
    txt = subprocess.run(['cuobjdump', '-sass', '-fun', fn, path], capture_output=True, text=True).stdout
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r'/\*([0-9a-f]{4,5})\*/\s+([^;]+);', txt)]
    loops = []
    for a, t in ins:
        m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            loops.append((tgt, a, [x for addr, x in ins if tgt <= addr <= a]))
    return loops


def test_band_pass_loop_of_the_sampler_kernel_is_spill_free_and_vectorised(built_library):
    """Performance guard on the SASS of nuts_kernel<11,6,2,0,1> (W22 fits, the bench kernel): the innermost
    loop that holds the exp2 of the band-pass integral reads the J_l row with LDS.128, does the two knot
    contractions (>= 22 FFMA per round), and neither it nor the pair loop around it touches local memory
    (the kernel's few spills belong to the sampler's bookkeeping)."""
    from bayesn_b200 import native
    loops = _sass_loops(native.LIB_PATH, 'nuts_kernelILi11ELi6ELi2ELb0ELb1')
    ex2 = sorted((len(b), b) for _, _, b in loops if any('MUFU.EX2' in x for x in b))
    assert len(ex2) >= 2
    inner, pair = ex2[0][1], ex2[1][1]
    rounds = sum('MUFU.EX2' in x for x in inner)
    assert rounds >= 1 and len(inner) <= 50 * rounds, (rounds, len(inner))
    assert sum(x.startswith('LDS.128') or ' LDS.128' in x for x in inner) >= 3 * rounds
    # packed FP32 (FFMA2: two multiply-adds per issue slot) counts twice
    assert sum(2 if 'FFMA2' in x else ('FFMA' in x) for x in inner) >= 22 * rounds
    for body in (inner, pair):
        assert not any(('LDL' in x) or ('STL' in x) for x in body)
    # the L_Sigma column walks use immediate-offset loads (compile-time row stride)
    txt = ' '.join(x for _, _, b in loops for x in b)
    assert re.search(r'LDG\.E\.CONSTANT \S+, desc\[\S+\]\[\S+\+0x[0-9a-f]+\]', txt)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_path_fails_loudly_without_gpu():
    from bayesn_b200 import native
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        native.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'bayesn_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in re.sub(r'#.*', '', src).replace('"""', ''), fn


def test_pack_dataset_layout():
    from bayesn_b200 import packing
    rng = np.random.RandomState(0)
    N, N_obs, n_b = 3, 5, 4
    obs = np.zeros((10, N_obs, N))
    obs[0] = rng.uniform(-8, 30, (N_obs, N))
    obs[1] = rng.uniform(1e3, 1e4, (N_obs, N))
    obs[2] = rng.uniform(10, 100, (N_obs, N))
    obs[4] = rng.randint(1, n_b, (N_obs, N))
    obs[7] = np.array([34.0, 35.0, 36.0])[None]
    obs[9] = 1
    obs[9, 1, 0] = 0
    obs[9, :, 2] = 0                    # an SN with no valid observation at all
    w = np.zeros((N, 300, n_b))
    w[:, 40:90, 1] = 0.01
    w[:, 100:101, 2] = 0.5
    w[:, :, 0] = 1 / 300
    zps = np.array([10.0, -20.8, -21.4, -21.8])
    off = np.zeros(n_b)
    p = packing.pack_dataset(obs, w, zps, off, -19.5, 5.0, 0.2, 24, device='cpu')
    assert p['n_valid'].tolist() == [4, 5, 0]
    o4 = p['obs4'].numpy()
    assert np.all(o4[2] == 0)
    # masked observation moved behind the valid ones
    assert np.allclose(np.sort(o4[0, :4, 0]), np.sort(np.delete(obs[0, :, 0], 1)), rtol=1e-6)
    band = o4[..., 3].view(np.int32)
    # valid observations are ordered by decreasing band-support width (pairs of observations share
    # a half-warp round count), each one keeping its own (t, y, 1/sigma, band)
    order = p['order'][0, :4]
    assert sorted(order.tolist()) == [0, 2, 3, 4]
    assert np.all(band[0, :4] == obs[4, order, 0].astype(int))
    assert np.allclose(o4[0, :4, 0], obs[0, order, 0], rtol=1e-6) and np.allclose(o4[0, :4, 2], 1 / obs[2, order, 0], rtol=1e-6)
    width = {1: 50, 2: 1, 3: 0}
    ws = [width[b] for b in band[0, :4]]
    assert ws == sorted(ws, reverse=True)
    sup = p['support'].numpy()
    # SN 0 observes bands drawn from {1, 2, 3}; band 3 has an all-zero weight row, band 0 is unused
    used = set(band[0, :4].tolist())
    exp = {1: [40, 90], 2: [100, 101], 3: [0, 0]}
    off = 0
    for k in range(n_b):
        lo, hi = exp[k] if k in used and k in exp else [0, 0]
        assert sup[0, k, :2].tolist() == [lo, hi]
        if hi > lo:
            assert sup[0, k, 2] == off
            off += hi - lo
    ws = p['wt_stride']
    assert ws % 4 == 0 and p['weights'].shape == (N, ws)
    assert p['_weights_storage'].numel() >= N * ws + 320
    sc = 10 ** (0.4 * 27.5 + 20.8 / 2.5 - 0.4 * (-19.5 + 34.0))
    if 1 in used:
        assert np.isclose(float(p['weights'][0, sup[0, 1, 2] + 10]), 0.01 * sc, rtol=1e-6)
    assert float(p['weights'][2].abs().sum()) == 0          # SN without valid observations: empty tile
    with pytest.raises(ValueError):
        obs[4, 0, 1] = 7
        packing.pack_dataset(obs, w, zps, off, -19.5, 5.0, 0.2, 24, device='cpu')


def test_shard_bounds_cover_everything():
    from bayesn_b200.distributed import shard_bounds
    for n in [0, 1, 7, 1000, 100000]:
        for world in [1, 2, 3, 8]:
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from bayesn_b200.distributed import shard_obs, gather_rows, shard_bounds
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%s' % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
N = 7
obs = np.arange(10 * 3 * N, dtype=np.float64).reshape(10, 3, N)
w = np.arange(N * 300 * 2, dtype=np.float64).reshape(N, 300, 2)
o, ww, (lo, hi) = shard_obs(obs, w, rank, 2)
assert o.shape[-1] == hi - lo == ww.shape[0]
# stand-in for the per-SN fit: a deterministic function of each SN's own data only
local = np.stack([o[1].sum(axis=0), ww.sum(axis=(1, 2))], axis=1)
full = gather_rows(local, N)
want = np.stack([obs[1].sum(axis=0), w.sum(axis=(1, 2))], axis=1)
assert full.shape == (N, 2) and np.array_equal(full, want), (rank, full, want)
dist.barrier(); dist.destroy_process_group()
print('ok', rank)
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and 'ok' in o, o


def test_training_transforms_roundtrip_and_match_oracle():
    """Host side of the training path: CorrCholesky forward/inverse, unconstrain -> constrain round
    trip, and agreement with the oracle's (torch) restatement of numpyro's transform."""
    import torch
    from bayesn_b200 import SEDmodel, train as btrain
    from oracle.ref_model import RefSED
    m = SEDmodel(load_model='T21_model')
    d = btrain.dims(m)
    ne = d['n_eps']
    rng = np.random.RandomState(3)
    y = 0.4 * rng.randn(d['n_corr'])
    Lo = btrain.corr_cholesky(y, ne)
    assert np.allclose((Lo ** 2).sum(axis=1), 1.0) and np.allclose(np.triu(Lo, 1), 0.0)
    Lo_t, _ = RefSED.corr_cholesky(torch.as_tensor(y), ne)
    assert np.allclose(Lo, Lo_t.numpy(), atol=1e-14)
    assert np.allclose(btrain.corr_cholesky_inverse(Lo), y, atol=1e-9)
    obs = np.zeros((10, 4, 5))
    obs[5], obs[7] = 0.03, 35.5
    init = btrain.initial_values(m, obs, 'T21_model', rng)
    qg, ql = btrain.unconstrain(m, init, 5, rng)
    assert qg.shape == (d['G'],) and ql.shape == (5, d['Dl'])
    c = btrain.constrain(m, qg[None, None], ql[:, None, None])
    assert np.allclose(c['W0'][0, 0], init['W0']) and np.allclose(c['sigmaepsilon_tform'][0, 0], init['sigmaepsilon_tform'])
    assert np.allclose(c['L_Omega'][0, 0], np.eye(ne), atol=1e-6) and np.isclose(c['RV'][0, 0], 3.0)
    assert np.isclose(c['tauA_tform'][0, 0], init['tauA_tform']) and np.allclose(c['AV'][0, 0], init['AV'])
    assert c['eps'].shape == (1, 1, ne, 5) and c['theta'].shape == (1, 1, 5)
    # same constrained values as the oracle's unpacking of the same vector
    g, _ = RefSED('T21_model', bands=['g_PS1']).train_unpack(torch.as_tensor(qg))
    assert np.allclose(g['L_Sigma'].numpy(), c['sigmaepsilon'][0, 0][:, None] * c['L_Omega'][0, 0], atol=1e-12)
    assert np.isclose(float(g['sigma0']), c['sigma0'][0, 0]) and np.isclose(float(g['tauA']), c['tauA'][0, 0])


def test_training_postprocess_sign_and_scale(tmp_path):
    """Reference postprocess :2186-2242: chains whose W1 has the opposite sign are flipped together
    with theta, theta is rescaled to unit spread and bayesn.yaml carries the posterior means."""
    import yaml
    from bayesn_b200 import SEDmodel, train as btrain
    m = SEDmodel(load_model='T21_model')
    L, T = len(m.l_knots), len(m.tau_knots)
    ne = (L - 2) * T
    rng = np.random.RandomState(0)
    C, S, N = 2, 6, 5
    W1 = np.repeat(m.W1.flatten(order='F')[None, None], C, 0).repeat(S, 1)
    W1[1] *= -1
    theta = 2.0 * rng.randn(C, S, N)
    samples = {'W0': np.repeat(m.W0.flatten(order='F')[None, None], C, 0).repeat(S, 1), 'W1': W1, 'theta': theta.copy(),
               'sigmaepsilon': np.full((C, S, ne), 0.1), 'L_Omega': np.tile(np.eye(ne), (C, S, 1, 1)),
               'sigma0': np.full((C, S), 0.1), 'tauA': np.full((C, S), 0.3), 'RV': np.full((C, S), 2.6)}
    ydat, out = btrain.postprocess_training(m, samples, str(tmp_path))
    assert np.allclose(np.std(out['theta'], axis=2), 1.0)
    assert np.all(np.sign(out['W1'][1, 0]) == np.sign(out['W1'][0, 0]))
    assert 'RV' not in ydat and np.isclose(ydat['TAUA'], 0.3)
    back = yaml.safe_load(open(tmp_path / 'bayesn.yaml'))
    assert np.allclose(back['W0'], m.W0) and np.allclose(back['L_SIGMA_EPSILON'], 0.1 * np.eye(ne))
    assert (tmp_path / 'chains.pkl').exists() and (tmp_path / 'initial_chains.pkl').exists()
    import pandas as pd
    summ = pd.read_csv(tmp_path / 'fit_summary.csv', index_col=0)          # arviz.summary layout, one row per scalar
    assert list(summ.columns) == ['mean', 'sd', 'hdi_3%', 'hdi_97%', 'mcse_mean', 'mcse_sd', 'ess_bulk', 'ess_tail', 'r_hat']
    assert 'theta[4]' in summ.index and f'L_Omega[{ne - 1},{ne - 1}]' in summ.index and 'tauA' in summ.index
    assert np.isfinite(summ.loc['theta[0]', 'r_hat']) and np.isnan(summ.loc['tauA', 'r_hat'])      # constants: NaN, as arviz


def test_process_dataset_reads_snana_text_directory(tmp_path):
    """Reference process_dataset, SNANA text branch (:2577-2730): phases, filter map, band drops,
    padding, training magnitudes and masks."""
    from bayesn_b200 import SEDmodel, snana
    m = SEDmodel(load_model='T21_model')
    d = tmp_path / 'SIMX'
    d.mkdir()
    rng = np.random.RandomState(0)
    names = []
    for i, (z, n) in enumerate([(0.03, 6), (0.05, 9)]):
        mjd = 57000 + np.linspace(-5, 30, n) * (1 + z)
        flt = np.array(['g', 'r', 'i'])[np.arange(n) % 3]
        flux = rng.uniform(1e3, 1e4, n)
        if i == 1:
            flux[2] = -5.0                               # non-positive flux: masked in training
            mjd[0] = 57000 - 30 * (1 + z)                # outside the tau-knot range: dropped
        snana.write_snana_fluxfile(str(d), f'sn{i}', mjd, flt, flux, 0.05 * np.abs(flux), 57000.0, z, z + 0.001, ebv_mw=0.02 * i)
        names.append(f'sn{i}.DAT')
    (d / 'SIMX.LIST').write_text('\n'.join(names) + '\n')
    args = dict(mode='fitting', version_photometry=str(d), map={'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1'})
    ds = m.process_dataset(args)
    obs = ds['obs']
    assert obs.shape == (10, 8, 2) and ds['sn_names'] == ['sn0', 'sn1']
    assert list(ds['band_inds']) == [0, m.band_dict['g_PS1'], m.band_dict['r_PS1'], m.band_dict['i_PS1']]
    assert obs[9, :, 0].sum() == 6 and obs[9, :, 1].sum() == 8
    assert np.allclose(obs[0, :6, 0], np.linspace(-5, 30, 6)) and np.all(obs[4, :6, 0] == [1, 2, 3, 1, 2, 3])
    assert np.isclose(obs[7, 0, 1], m.cosmo.distmod(np.array([0.051]))[0]) and obs[8, 0, 1] == 0.02
    assert np.allclose(obs[2, 6:, 0], 1 / np.sqrt(2 * np.pi)) and np.all(obs[1, 6:, 0] == 0)
    tr = m.process_dataset(dict(args, mode='training_globalRV'))['obs']
    assert tr[9, :, 1].sum() == 7                        # the negative flux is masked
    k = int(np.argmin(tr[9, :8, 1]))
    assert tr[1, k, 1] == 0 and tr[4, k, 1] == 0
    good = tr[9, :, 0] != 0
    assert np.allclose(tr[1, good, 0], 27.5 - 2.5 * np.log10(obs[1, good, 0]))


def test_parse_yaml_input_defaults_and_errors(tmp_path):
    """Reference parse_yaml_input :1656-1709: defaults, command-line overrides, fit_method validation."""
    import argparse
    from bayesn_b200 import SEDmodel, driver
    m = SEDmodel(load_model='T21_model')
    a = driver.parse_yaml_input(m, {'mode': 'fitting', 'outputdir': str(tmp_path / 'o')})
    assert (a['num_chains'], a['num_warmup'], a['num_samples'], a['fit_method'], a['initialisation']) == (4, 500, 500, 'mcmc', 'median')
    assert a['outfile_prefix'] == 'output' and a['l_knots'] == m.l_knots.tolist() and (tmp_path / 'o').is_dir()
    cmd = argparse.Namespace(input='x.yaml', filters=None, num_warmup=100, num_samples=None)
    a = driver.parse_yaml_input(m, {'mode': 'fitting', 'num_warmup': 250, 'outputdir': str(tmp_path)}, cmd)
    assert a['num_warmup'] == 100 and a['num_samples'] == 500 and 'input' not in a
    with pytest.raises(ValueError):
        driver.parse_yaml_input(m, {'mode': 'fitting', 'fit_method': 'laplace', 'outputdir': str(tmp_path)})
    with pytest.raises(NotImplementedError):
        driver.parse_yaml_input(m, {'mode': 'fitting', 'fit_method': 'VI', 'outputdir': str(tmp_path)})
    with pytest.raises(ValueError):
        driver.process_dataset(m, {'mode': 'fitting'})


def test_mu_draws_do_not_depend_on_the_sharding():
    """postprocess draws mu ~ N(.., ..) per posterior draw (reference :2250-2254); the noise is keyed by
    (seed, global block of SNe), so a rank's block of a sharded job gets exactly the draws of the unsharded run."""
    from bayesn_b200 import outputs

    class M:
        sigma0 = 0.1
    g = torch.Generator().manual_seed(0)
    N, C, S = 2500, 2, 5
    Ds = 35 + torch.randn((N, C, S), generator=g)
    muhat = 35 + np.arange(N) * 1e-3
    full = outputs.draw_mu_device(M, {'Ds': Ds}, muhat, seed=3)
    for lo, hi in ((0, 700), (700, 1100), (1100, 2500)):
        part = outputs.draw_mu_device(M, {'Ds': Ds[lo:hi]}, muhat[lo:hi], seed=3, sn_offset=lo)
        assert torch.equal(part['mu'], full['mu'][lo:hi]) and torch.equal(part['delM'], full['delM'][lo:hi])
    # the conditional of mu given Ds: mean (25 Ds + s0^2 muhat) / (25 + s0^2), sd s0 5 / sqrt(25 + s0^2)
    resid = (full['mu'] - (25 * Ds.double() + 0.01 * torch.as_tensor(muhat)[:, None, None]) / 25.01)
    assert abs(float(resid.std()) - 0.1 * 5 / np.sqrt(25.01)) < 2e-3 and abs(float(resid.mean())) < 2e-3


def test_training_prior_initialisation_modes():
    """initialisation: median | sample for training (reference :1756-1759): finite unconstrained positions, a
    different one per chain, L_Omega at the identity for the median."""
    from bayesn_b200 import train as btrain

    class M:
        l_knots = np.linspace(3000., 18000., 6)
        tau_knots = np.linspace(-10., 40., 6)
    d = btrain.dims(M)
    rng = np.random.RandomState(0)
    obs = np.zeros((10, 4, 7))
    obs[-5] = rng.uniform(0.02, 0.08, 7)[None]
    obs[-4] = 1e-4
    obs[-3] = 35.0
    for how in ('median', 'sample'):
        qg, ql = btrain.prior_position(M, obs, np.random.RandomState(1), how)
        assert qg.shape == (d['G'],) and ql.shape == (7, d['Dl']) and np.all(np.isfinite(qg)) and np.all(np.isfinite(ql))
        qg2, _ = btrain.prior_position(M, obs, np.random.RandomState(2), how)
        assert not np.allclose(qg, qg2)
    qg, _ = btrain.prior_position(M, obs, np.random.RandomState(1), 'median')
    ne = d['n_eps']
    y = qg[2 * 36 + ne:d['G'] - 3]
    assert np.max(np.abs(y)) < 0.6 and np.max(np.abs(qg[:72])) < 1.0          # medians of 15 draws sit near the prior medians


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The boundary is a plain C ABI: the ctypes mirrors in bayesn_b200/native.py must have the size and field offsets
    a C compiler gives the structs of include/bayesn_b200.h (a reference-side binding would be written against that
    header)."""
    import ctypes as C
    from bayesn_b200 import native
    pairs = {'bayesn_model_desc': native.ModelDesc, 'bayesn_dataset_desc': native.DatasetDesc,
             'bayesn_nuts_cfg': native.NutsCfg, 'bayesn_tiles_desc': native.TilesDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bayesn_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / 'abi.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'abi'
    subprocess.check_call(['gcc', '-std=c11', '-I' + os.path.join(ROOT, 'include'), '-o', str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        cname, what, val = line.split()
        cls = pairs[cname]
        if what == 'size':
            assert C.sizeof(cls) == int(val), (cname, C.sizeof(cls), val)
        else:
            assert getattr(cls, what).offset == int(val), (cname, what, getattr(cls, what).offset, val)


def test_parse_yaml_input_snana_batch_keys(tmp_path):
    """jobsplit switches the SNANA batch mode on (reference :1703-1709); save_chains / warps_per_chain defaults."""
    from bayesn_b200 import driver

    class M:
        l_knots = np.array([3000., 5000.])
        tau_knots = np.array([-10., 40.])
    a = driver.parse_yaml_input(M, dict(mode='fitting', outputdir=str(tmp_path / 'o')))
    assert a['snana'] is False and a['jobsplit'] == [1, 1] and a['save_chains'] is True and a['warps_per_chain'] == 'auto'
    assert (tmp_path / 'o').is_dir()
    b = driver.parse_yaml_input(M, dict(mode='fitting', outputdir=str(tmp_path / 'p'), jobsplit=[3, 8], sim_prescale=2))
    assert b['snana'] is True and b['jobid'] == 3 and b['njobtot'] == 16
    assert not (tmp_path / 'p').exists()             # batch fitting writes next to outfile_prefix, not into outputdir
    with pytest.raises(ValueError):
        driver.parse_yaml_input(M, dict(outputdir=str(tmp_path)))
    with pytest.raises(NotImplementedError):
        driver.parse_yaml_input(M, dict(mode='fitting', fit_method='vi', outputdir=str(tmp_path)))


def test_survey_id_from_sndata_root(tmp_path, monkeypatch):
    from bayesn_b200 import driver
    monkeypatch.delenv('SNDATA_ROOT', raising=False)
    assert driver.survey_id('LSST') == 0
    (tmp_path / 'SURVEY.DEF').write_text('# comment\nSURVEY: SDSS 1\nSURVEY: LSST 12\n')
    monkeypatch.setenv('SNDATA_ROOT', str(tmp_path))
    assert driver.survey_id('LSST') == 12 and driver.survey_id('SDSS') == 1 and driver.survey_id('NOPE') == 0


def test_output_writers_reference_layout(tmp_path):
    """host_samples turns device draws (N, C, S[, dim]) into the reference's (chains, draws[, dim], N); write_chains
    writes chains.pkl + fit_summary.csv; write_fitres drops NaN distances; write_lcplot keeps the reference's columns."""
    import pickle
    import pandas as pd
    from bayesn_b200 import driver, outputs
    g = torch.Generator().manual_seed(1)
    N, C, S, ne = 3, 2, 6, 4
    cons = {'theta': torch.randn((N, C, S), generator=g), 'eps': torch.randn((N, C, S, ne), generator=g)}
    hs = outputs.host_samples(cons)
    assert hs['theta'].shape == (C, S, N) and hs['eps'].shape == (C, S, ne, N)
    assert hs['eps'][1, 2, 3, 0] == pytest.approx(float(cons['eps'][0, 1, 2, 3]))
    outputs.write_chains(hs, str(tmp_path))
    back = pickle.load(open(tmp_path / 'chains.pkl', 'rb'))
    assert np.array_equal(back['eps'], hs['eps'])
    summ = pd.read_csv(tmp_path / 'fit_summary.csv', index_col=0)
    assert 'theta[0]' in summ.index and 'eps[3,2]' in summ.index and {'mean', 'sd', 'r_hat'} <= set(summ.columns)
    cols = {'zHEL': np.array([0.1, 0.2, 0.3]), 'MU_LCFIT': np.array([35.0, np.nan, 36.0])}
    assert outputs.write_fitres(str(tmp_path / 'x.FITRES.TEXT'), ['a', 'b', 'c'], cols) == 1
    rows = [l for l in open(tmp_path / 'x.FITRES.TEXT') if l.startswith('SN:')]
    assert len(rows) == 2 and rows[0].split()[:2] == ['SN:', 'a']

    class M:
        tau_knots = np.array([-10., 0., 10.])
    t = np.arange(-10., 10., 2)
    ds = dict(obs=np.zeros((10, 2, 2)), peak_mjds=np.array([58000., 58100.]),
              lcplot=[dict(CID='s1', MJD=np.array([57999., 58001.]), FLUXCAL=np.array([1., 2.]), FLUXCALERR=np.array([.1, .1]),
                           FLT=np.array(['g_PS1', 'r_PS1'])),
                      dict(CID='s2', MJD=np.array([58099.]), FLUXCAL=np.array([3.]), FLUXCALERR=np.array([.2]), FLT=np.array(['g_PS1']))])
    ds['obs'][5, 0, :] = [0.01, 0.02]
    fm = np.arange(2 * 2 * len(t), dtype=float).reshape(2, 2, len(t))
    out = driver.write_lcplot(M, str(tmp_path / 'x.LCPLOT'), ds, t, fm, np.zeros_like(fm))
    assert list(out.columns) == ['CID', 'MJD', 'FLUXCAL', 'FLUXCALERR', 'FLT', 'DATA_FLAG']
    assert (out.DATA_FLAG == 1).sum() == 3 and (out.DATA_FLAG == 0).sum() == 2 * len(t) + 1 * len(t)
    s2 = out[(out.CID == 's2') & (out.DATA_FLAG == 0)]
    assert np.allclose(np.sort(s2.MJD.values), 58100. + t * 1.02) and set(s2.FLT) == {'g_PS1'}


def _write_objects(d, model, n=5, seed=4):
    """A few SNANA text files with hand-made griz photometry (no forward model needed on the CPU)."""
    from bayesn_b200 import snana
    rng = np.random.RandomState(seed)
    tgrid = np.arange(-12, 44, 4.0)                  # two epochs fall outside the tau-knot range (-10, 40)
    bands = np.repeat(['g', 'r', 'i', 'z'], len(tgrid))
    tt = np.tile(tgrid, 4)
    files, zs = [], rng.uniform(0.02, 0.06, n)
    for i in range(n):
        flux = rng.uniform(500, 5000, len(tt))
        flux[3] = -5.0                                # a negative flux: masked for training, kept for fitting
        err = 0.03 * np.abs(flux) + 1.0
        snana.write_snana_fluxfile(str(d), f'{100 + i}', 58000.0 + tt * (1 + zs[i]), bands, flux, err, 58000.0, zs[i], zs[i] + 1e-4,
                                   ebv_mw=0.01 * i, extra={'SURVEY': 'TESTSURVEY', 'TYPE': 1, 'FIELD': 'DEEP', 'VPEC': 100.0,
                                                           'HOSTGAL_LOGMASS': 10.5})
        files.append(f'{100 + i}.DAT')
    name = os.path.split(str(d))[-1]
    (d / f'{name}.LIST').write_text('\n'.join(files) + '\n')
    return files, zs, tt


def test_process_dataset_text_directory_jobsplit_and_data_table(tmp_path):
    """process_dataset (reference :2577-2925) on the CPU: padding, phase cut, band order, FITRES metadata in the
    reference's column order, SNANA batch subsetting, and the data_table branch reading the same files."""
    from bayesn_b200 import SEDmodel, driver
    m = SEDmodel(load_model='T21_model')
    d = tmp_path / 'PHOT'
    d.mkdir()
    files, zs, tt = _write_objects(d, m)
    fmap = {'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'}
    args = driver.parse_yaml_input(m, dict(mode='fitting', version_photometry=str(d), outputdir=str(tmp_path / 'o'), map=fmap))
    ds = driver.process_dataset(m, args)
    N, n_in = 5, int(((tt > -10 * 1.0) & (tt < 40)).sum())
    assert ds['obs'].shape[0] == 10 and ds['obs'].shape[2] == N and ds['sn_names'] == [str(100 + i) for i in range(N)]
    # (rest-frame phases recomputed from 5-decimal MJDs land on either side of the tau-knot edges: ragged light curves)
    cnt = ds['obs'][9].sum(axis=0)
    assert np.all(cnt <= n_in + 4) and np.all(cnt >= n_in - 8) and ds['obs'].shape[1] == int(cnt.max())
    assert np.all(ds['obs'][0][ds['obs'][9] != 0] > -10) and np.all(ds['obs'][0][ds['obs'][9] != 0] < 40)
    assert list(ds['band_inds']) == [0] + [m.band_dict[fmap[b]] for b in 'griz']         # NULL + order of first use
    assert list(ds['fitres'].keys()) == list(driver.FITRES_COLUMNS)
    assert np.allclose(ds['fitres']['zHEL'], zs) and np.all(ds['fitres']['FIELD'] == 'DEEP') and np.all(ds['fitres']['TYPE'] == 1)
    assert np.all(ds['fitres']['SNRMAX1'] >= ds['fitres']['SNRMAX2']) and np.all(ds['fitres']['SNRMAX2'] >= ds['fitres']['SNRMAX3'])
    assert ds['survey'] == 'TESTSURVEY' and ds['sim'] is False
    # peculiar velocity enters the Hubble-diagram redshift of the distance prior (:2598-2600)
    zpec = np.sqrt((1 + 100.0 / driver.C_KMS) / (1 - 100.0 / driver.C_KMS)) - 1
    assert np.allclose(ds['obs'][7, 0, :], m.cosmo.distmod((1 + zs + 1e-4) / (1 + zpec) - 1))
    # training: magnitudes, non-positive fluxes masked
    dt = driver.process_dataset(m, dict(args, mode='training_globalRV'))
    assert dt['obs'][9].sum() == ds['obs'][9].sum() - N and np.all(dt['obs'][1][dt['obs'][9] != 0] > 15)
    # SNANA batch job 2 of 3 takes objects 2 and 5 (1-based)
    a2 = driver.parse_yaml_input(m, dict(mode='fitting', version_photometry=str(d), outputdir=str(tmp_path / 'o'), map=fmap,
                                         jobsplit=[2, 3]))
    assert driver.process_dataset(m, a2)['sn_names'] == ['101', '104']
    # the data_table branch on the same files
    tab = tmp_path / 'table.txt'
    tab.write_text('# comment\nSNID SEARCH_PEAKMJD REDSHIFT_CMB REDSHIFT_CMB_ERR files\n' +
                   '\n'.join(f'{100 + i} 58000.0 {zs[i] + 1e-4} 0.0005 PHOT/{files[i]}' for i in range(N)) + '\n')
    at = driver.parse_yaml_input(m, dict(mode='fitting', data_table='table.txt', data_root=str(tmp_path), outputdir=str(tmp_path / 'o'),
                                         map=fmap))
    d2 = driver.process_dataset(m, at)
    assert d2['sn_names'] == ds['sn_names'] and np.array_equal(d2['obs'][:5], ds['obs'][:5]) and np.array_equal(d2['obs'][9], ds['obs'][9])
    assert np.allclose(d2['obs'][7, 0, :], m.cosmo.distmod(zs + 1e-4)) and np.all(d2['obs'][6] == 0.0005)
    with pytest.raises(ValueError):
        driver.process_dataset(m, dict(at, data_root=None) if False else {k: v for k, v in at.items() if k != 'data_root'})


def test_snana_batch_mode_finds_photometry_version_by_name(tmp_path, monkeypatch):
    """With ``jobsplit`` (SNANA batch mode) ``version_photometry`` is the NAME of a photometry version searched under
    $SNDATA_ROOT/lcmerge, $SNDATA_ROOT/SIM, the directories of PATH_SNDATA_SIM.LIST and ``private_data_path``
    (reference :2401-2426); IDSURVEY comes from $SNDATA_ROOT/SURVEY.DEF (:2427-2433)."""
    from bayesn_b200 import SEDmodel, driver
    m = SEDmodel(load_model='T21_model')
    root = tmp_path / 'sndata'
    (root / 'SIM').mkdir(parents=True)
    (root / 'lcmerge').mkdir()
    extra = tmp_path / 'moresims'
    extra.mkdir()
    (root / 'SIM' / 'PATH_SNDATA_SIM.LIST').write_text('$MORESIMS\n')
    (root / 'SURVEY.DEF').write_text('SURVEY: OTHER 5\nSURVEY: TESTSURVEY 42\n')
    monkeypatch.setenv('SNDATA_ROOT', str(root))
    monkeypatch.setenv('MORESIMS', str(extra))
    monkeypatch.setenv('MYPRIV', str(tmp_path))
    fmap = {'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'}
    for where, name in ((root / 'SIM', 'VER_A'), (extra, 'VER_B'), (tmp_path / 'priv' / 'deep', 'VER_C')):
        d = where / name
        d.mkdir(parents=True)
        _write_objects(d, m)
    base = dict(mode='fitting', outputdir=str(tmp_path / 'o'), map=fmap, jobsplit=[1, 1])
    for name, extra_args in (('VER_A', {}), ('VER_B', {}), ('VER_C', dict(private_data_path='$MYPRIV/priv/deep'))):
        args = driver.parse_yaml_input(m, dict(base, version_photometry=name, **extra_args))
        assert args['snana'] and isinstance(args['private_data_path'], list)
        ds = driver.process_dataset(m, args)
        assert len(ds['sn_names']) == 5 and ds['survey'] == 'TESTSURVEY' and np.all(np.asarray(ds['fitres']['IDSURVEY']) == 42)
    with pytest.raises(ValueError, match='not found in any of the usual public'):
        driver.process_dataset(m, driver.parse_yaml_input(m, dict(base, version_photometry='VER_C')))
    (root / 'lcmerge' / 'VER_A').mkdir()
    with pytest.raises(ValueError, match='found in multiple locations'):
        driver.process_dataset(m, driver.parse_yaml_input(m, dict(base, version_photometry='VER_A')))
    with pytest.raises(ValueError, match='Please pass either data_dir'):
        driver.process_dataset(m, driver.parse_yaml_input(m, dict(mode='fitting', outputdir=str(tmp_path / 'o'))))


def test_simulate_spectrum_argument_errors_match_reference():
    """simulate_spectrum (reference :2962-3056) validates its per-object arguments before any device work."""
    from bayesn_b200 import SEDmodel
    m = SEDmodel(load_model='T21_model')
    t = np.array([0.0, 10.0])
    for kw, msg in [(dict(del_M=np.zeros(2)), 'del_M'), (dict(AV=np.zeros(4)), 'AV'), (dict(theta=np.zeros(2)), 'theta'),
                    (dict(eps=0.5), 'only scalar value accepted is 0'), (dict(eps=np.zeros((2, 2))), 'shape'),
                    (dict(z=np.zeros(2)), 'For z'), (dict(mu=np.zeros(2)), 'For mu'), (dict(RV=np.ones(2)), 'For RV'),
                    (dict(ebv_mw=np.zeros(2)), 'For ebv_mw')]:
        with pytest.raises(ValueError, match=msg):
            m.simulate_spectrum(t, 3, **kw)


def test_f99_ultraviolet_branch_lives_in_the_dust_table(tmp_path):
    """Host-dust law below 2700 A (reference get_spectra :627-638): the FM90 form k = c1 + c2 x + c3 D (+ c4 F) is
    P(x) + Q(x) / R_V, which the two ultraviolet anchors of the F99 spline span - so the rows of ``M_fitz`` carry it
    exactly for every R_V, and the oracle's explicit override agrees."""
    import torch
    from bayesn_b200 import SEDmodel
    from bayesn_b200.static import fitz_matrix, f99_yk, f99_alav
    from oracle.ref_model import RefSED
    from tests.common import write_uv_model
    wave = np.concatenate([np.linspace(1150, 2699.99, 300), np.linspace(2700, 18000, 60)])
    M = fitz_matrix(wave)
    assert np.all(M[:300, :7] == 0) and np.any(M[300:, :7] != 0)
    for RV in (1.2, 2.0, 3.1, 4.5, 6.0):
        yk, dyk = f99_yk(np.float64(RV))
        assert np.allclose(1 + M @ yk / RV, f99_alav(wave, RV), rtol=0, atol=1e-12)
        h = 1e-5                                            # and the R_V derivative the sampled-R_V kernels use
        fd = (f99_alav(wave, RV + h) - f99_alav(wave, RV - h)) / (2 * h)
        assert np.allclose(M @ dyk / RV - M @ yk / RV ** 2, fd, atol=1e-7)
    path = write_uv_model(tmp_path)
    m = SEDmodel(load_model=path)                           # accepted now (it used to be rejected)
    ref = RefSED(path, bands=['g_PS1'])
    nuv = int((m.grid.model_wave < 2700).sum())
    assert nuv > 20 and ref.uv_ind1.sum() == nuv
    N = 4
    RV = np.array([1.5, 2.5, 3.1, 5.0])
    AV = np.array([0.1, 0.5, 1.0, 0.3])
    t = torch.zeros(1, N, dtype=torch.float64)
    J_t, hsi = ref.J_t_and_hsiao(t)
    tn = lambda a: torch.as_tensor(np.asarray(a, dtype=float))
    z = np.zeros
    s1 = ref.get_spectra(tn(z(N)), tn(AV), tn(ref.W0), tn(ref.W1), tn(z((N,) + ref.W0.shape)), tn(RV), J_t, hsi).numpy()
    s0 = ref.get_spectra(tn(z(N)), tn(0 * AV), tn(ref.W0), tn(ref.W1), tn(z((N,) + ref.W0.shape)), tn(RV), J_t, hsi).numpy()
    A_oracle = -2.5 * np.log10(s1[:, :, 0] / s0[:, :, 0])
    a_prod, _ = m.grid.dust_basis(RV)
    assert np.allclose(A_oracle, AV[:, None] * a_prod, rtol=1e-10, atol=1e-12)
    assert np.allclose(a_prod, np.stack([f99_alav(m.grid.model_wave, r) for r in RV]), atol=1e-12)
    # the override is not cosmetic: the optical spline extrapolated below 2700 A is off by several per cent
    from bayesn_b200.static import spline_matrix, inv_kd, F99_XK
    a_spline = 1 + (spline_matrix(1e4 / m.grid.model_wave, F99_XK, inv_kd(F99_XK)) @ f99_yk(np.float64(3.1))[0]) / 3.1
    assert np.max(np.abs(a_spline[:nuv] / a_prod[2, :nuv] - 1)) > 0.01


def test_training_knots_from_input_and_mirror_methods(tmp_path):
    """Training runs may set ``l_knots`` / ``tau_knots`` (reference :1721-1727): grids and tables are rebuilt, the
    starting values are the reference model interpolated onto them (``initial_guess`` :1568-1654, same keys and the
    same order of random draws); ``parse_yaml_input`` / ``spline_coeffs_irr_step`` exist on the model as in the reference."""
    from bayesn_b200 import SEDmodel, driver
    m = SEDmodel(load_model='T21_model')
    d = tmp_path / 'PHOT'
    d.mkdir()
    _write_objects(d, m)
    fmap = {'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'}
    lk, tk = [3500.0, 4500.0, 5500.0, 6500.0, 7700.0, 8700.0, 9500.0], [-10.0, 0.0, 10.0, 20.0, 30.0, 40.0]
    J_before = m.J_l_T.copy()
    args = m.parse_yaml_input(dict(mode='training_globalRV', version_photometry=str(d), outputdir=str(tmp_path / 'o'), map=fmap,
                                   l_knots=lk, tau_knots=tk))
    assert list(m.l_knots) == lk and m.J_l_T.shape == (300, 7) and J_before.shape == (300, 6)
    assert m.W0.shape == (7, 6) and m.L_Sigma.shape == (30, 30) and m.n_eps == 30
    assert m.data.shape[0] == 10 and m.hsiao_interp.shape == (3,) + m.data.shape[1:]
    np.random.seed(11)
    init = m.initial_guess(args, reference_model='T21_model')
    assert list(init.keys()) == ['W0', 'W1', 'RV', 'tauA_tform', 'sigma0_tform', 'sigma0', 'theta', 'AV', 'epsilon_tform',
                                 'epsilon', 'sigmaepsilon_tform', 'sigmaepsilon', 'L_Omega', 'Ds']
    N = m.data.shape[-1]
    assert init['W0'].shape == (42,) and init['epsilon_tform'].shape == (30, N) and init['epsilon'].shape == (N, 30)
    # the same seed replayed by hand in the reference's order reproduces the dictionary
    from bayesn_b200 import datafiles
    p = datafiles.load_model_yaml('T21_model')
    np.random.seed(11)
    tauA_ = float(p['TAUA']) + np.random.normal(0, 0.01)
    sigma0_ = 0.1 + np.random.normal(0, 0.01)
    j0, j1 = np.random.normal(0, 0.01, 42), np.random.normal(0, 0.01, 42)
    theta = np.random.normal(0, 1, N)
    AV = np.random.exponential(tauA_, N)
    assert np.isclose(init['sigma0'], sigma0_) and np.array_equal(init['theta'], theta) and np.array_equal(init['AV'], AV)
    W0 = np.array(p['W0'])
    # interpolation onto the new wavelength knots keeps the values at knots shared with the reference model
    assert np.allclose((init['W0'] - j0).reshape((7, 6), order='F')[[0, 6]], W0[[0, 5]])
    with pytest.raises(ValueError, match='Invalid initialisation method'):
        m.initial_guess(args, reference_model='no_such_model')
    # unchanged knots: nothing is rebuilt
    m2 = SEDmodel(load_model='T21_model')
    a2 = driver.parse_yaml_input(m2, dict(mode='training_globalRV', version_photometry=str(d), outputdir=str(tmp_path / 'o')))
    assert driver.apply_training_knots(m2, a2) is False
    row = SEDmodel.spline_coeffs_irr_step(3.0, m2.tau_knots, m2.KD_t)
    assert row.shape == (6,) and np.isclose(row.sum(), 1.0)


def test_module_entry_point_mirrors_run_bayesn(tmp_path, monkeypatch):
    """``python -m bayesn_b200 input.yaml --overrides``: the reference's ``run_bayesn`` script - default model T21,
    command-line values override the YAML through ``parse_yaml_input`` (:1670-1679), a missing input file or data set
    raises the reference's errors."""
    import yaml
    from bayesn_b200 import __main__ as cli, SEDmodel, driver
    with pytest.raises(FileNotFoundError, match='Specified input file'):
        cli.main([str(tmp_path / 'none.yaml')])
    inp = tmp_path / 'input.yaml'
    inp.write_text(yaml.safe_dump(dict(mode='fitting', num_chains=4, outputdir=str(tmp_path / 'out'))))
    with pytest.raises(ValueError, match='Please pass either data_dir'):
        cli.main([str(inp)])
    seen = {}

    def fake_run(self, args, cmd_args=None):
        seen['model'] = self
        seen['args'] = driver.parse_yaml_input(self, args, cmd_args)
        return 'ran'
    monkeypatch.setattr(SEDmodel, 'run', fake_run)
    mapf = tmp_path / 'map.txt'
    mapf.write_text('g g_PS1\nr r_PS1\n')
    out = cli.main([str(inp), '--num_chains', '2', '--num_warmup', '10', '--jobsplit', '2', '5', '--map', str(mapf),
                    '--drop_bands', 'z', 'y', '--load_model', 'M20_model', '--version_photometry', 'X'])
    a = seen['args']
    assert out == 'ran' and len(seen['model'].l_knots) == 9                  # M20 from the command line
    assert a['num_chains'] == 2 and a['num_warmup'] == 10 and a['num_samples'] == 500 and a['jobsplit'] == [2, 5]
    assert a['snana'] is True and a['jobid'] == 2 and a['njobtot'] == 5 and a['map'] == {'g': 'g_PS1', 'r': 'r_PS1'}
    assert a['drop_bands'] == ['z', 'y'] and a['version_photometry'] == 'X' and a['outfile_prefix'] == 'output'
