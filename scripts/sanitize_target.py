"""Smallest end-to-end case for compute-sanitizer: logp+grad (1 and 4 chains, fixed and sampled R_V), a short
NUTS run, the tile builder, training log-density and a few training sampler passes."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel, train as btrain
m = SEDmodel(load_model='T21_model')
np.random.seed(1)
z = np.random.uniform(0.02, 0.06, 5)
flux, err, par = m.simulate_light_curve(np.arange(-8, 36, 6.0), 5, ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'], yerr=0.03, z=z, mu='z', mag=False)
obs, w, bi = m.simulated_obs_array(flux, err, par)
obs[9, ::5, 1] = 0
for C in (1, 4):
    ctx = m.prepare(obs, None, bi)
    q = ctx.init_to_median(C, seed=1)
    lp, g = ctx.logp_grad(q)
    assert torch.isfinite(lp).all()
ctx_u = m.prepare(obs, w, bi, rv_mode=2)
qu = ctx_u.init_to_median(2, seed=1)
ctx_u.logp_grad(qu)
out = ctx.nuts(q, 12, 6, 5, seed=1)
assert torch.isfinite(out['samples']).all()
# round 2: latency mode (K warps per (SN, chain)) in the one-shot kernel and the sampler, both schedules
for K in (2, 4, 8):
    lpk, gk = ctx.logp_grad(q, warps_per_chain=K)
    assert torch.allclose(lpk, lp, rtol=2e-6) and torch.isfinite(gk).all()
    ok = ctx.nuts(q, 12, 6, 5, seed=1, warps_per_chain=K)
    assert torch.isfinite(ok['samples']).all()
obs_t, w_t, bi_t = btrain.simulate_training_set(m, 6, seed=2, bands=('g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'), t=np.arange(-8, 36, 6.0))
ct = m.prepare_training(obs_t, None, bi_t)
qg0, ql0 = btrain.initial_position(m, obs_t, 2, 'T21_model', seed=1)
ct.train_logp_grad(qg0, ql0)
o = ct.train_nuts(qg0, ql0, num_warmup=3, num_samples=2, max_tree_depth=4, seed=1, use_graph=False, passes_per_graph=4)
torch.cuda.synchronize()
print('sanitize target done', o['passes'], o['finished'])
