"""Reproduce one (SN, chain) of the C5 share in isolation (random streams are keyed by the global unit, so a
one-SN data set with sn_offset reproduces the chain exactly) and print its per-transition statistics.
usage: python scripts/stuck_chain_diag.py SN [N_total]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel
sn = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
bands = ['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST']
m = SEDmodel(load_model='W22_model')
np.random.seed(20241019)
n = min(4000, N)
z = np.random.uniform(0.02, 0.09, n)
data, yerr, params = m.simulate_light_curve(np.arange(-8, 40, 5.0), n, bands, yerr=0.05, err_type='mag', z=z, mu='z', ebv_mw=0, mag=False)
obs, w, bi = m.simulated_obs_array(data, yerr, params)
ctx = m.prepare(obs[:, :, sn:sn + 1], None, bi)
q0 = ctx.init_to_median(4, seed=0, sn_offset=sn)
out = ctx.nuts(q0, 250, 250, 10, seed=0, sn_offset=sn)
torch.cuda.synchronize()
st = out['stats'].cpu().numpy()[0]          # (C, W+S, 5): accept, n_leapfrog, diverging, potential energy, step
ci = out['chain_info'].cpu().numpy()[0]
ne = m.n_eps
names = ['theta', 'logAV', 'logit_tmax', 'Ds']
print('truth: theta %.3f AV %.3f z %.4f' % (params['theta'][sn], params['AV'][sn], z[sn]))
print('q0 scalars per chain:', q0[0, :, ne:].cpu().numpy())
lp, g = ctx.logp_grad(q0)
print('logp at q0:', lp.cpu().numpy(), ' |grad| max:', g.abs().amax(dim=-1).cpu().numpy(), 'argmax', g.abs().argmax(dim=-1).cpu().numpy())
print('grad scalars at q0:', g[0, :, ne:].cpu().numpy())
for c in range(4):
    print(f'chain {c}: leapfrogs {ci[c, 2]:.0f} step {ci[c, 0]:.3g} div {ci[c, 3]:.0f}')
    for it in list(range(0, 30)) + list(range(30, 500, 47)):
        a = st[c, it]
        print(f'   it {it:3d} accept {a[0]:.3f} n {a[1]:5.0f} div {a[2]:.0f} pe {a[3]:14.3f} step {a[4]:.3g}')
    if c == 0:
        zs = out['samples'][0, 0].cpu().numpy()
        print('   last sample scalars', zs[-1, ne:], ' eps_tform absmax', np.abs(zs[-1, :ne]).max())
