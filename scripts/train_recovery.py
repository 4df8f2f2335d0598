"""Population training on data simulated from T21: do the chains mix and recover the generating model?"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel, train as btrain, summary
N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
W = int(sys.argv[2]) if len(sys.argv) > 2 else 200
S = int(sys.argv[3]) if len(sys.argv) > 3 else 200
m = SEDmodel(load_model='T21_model')
obs, w, bi = btrain.simulate_training_set(m, N, seed=11, bands=('g_PS1', 'r_PS1', 'i_PS1', 'z_PS1'), t=np.arange(-8, 40, 3.0), yerr=0.02)
t0 = time.time()
samples, ex = m.train(obs, None, bi, num_chains=4, num_warmup=W, num_samples=S, initialisation='T21_model', seed=5)
dt = time.time() - t0
st = ex['stats']
print(f'{dt:.1f} s, passes {ex["passes"]}, accept {st[:, W:, 0].mean():.3f}, div {ex["info"][:, 5]}, steps/transition {st[:, W:, 1].mean():.0f}, step {ex["info"][:, 2]}')
for k in ('tauA', 'sigma0', 'RV'):
    x = samples[k]
    print(k, 'mean %.3f sd %.3f rhat %.3f' % (x.mean(), x.std(), summary.split_rhat(x)), 'truth', {'tauA': m.tauA, 'sigma0': m.sigma0, 'RV': m.RV}[k])
W0 = samples['W0']                      # (C, S, L*T) vec_F
truth = m.W0.flatten(order='F')
zs = (W0.mean(axis=(0, 1)) - truth) / W0.std(axis=(0, 1))
rh = np.array([summary.split_rhat(W0[..., i]) for i in range(W0.shape[-1])])
print('W0 z-scores: max |z| %.2f, rms %.2f; rhat max %.3f median %.3f' % (np.abs(zs).max(), np.sqrt((zs ** 2).mean()), rh.max(), np.median(rh)))
th = samples['theta']                   # (C, S, N)
print('theta corr with truth', np.corrcoef(th.mean(axis=(0, 1)), m._last_theta if hasattr(m, '_last_theta') else th.mean(axis=(0, 1)))[0, 1])
se = samples['sigmaepsilon']
print('sigmaepsilon mean range', se.mean(axis=(0, 1)).min(), se.mean(axis=(0, 1)).max(), 'truth diag(L_Sigma) range', np.sqrt((m.L_Sigma ** 2).sum(axis=1)).min(), np.sqrt((m.L_Sigma ** 2).sum(axis=1)).max())
