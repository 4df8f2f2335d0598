import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_gpu_train import _small_training_problem
from oracle import ref_fit
case, ref, m, ctx, qg32, ql32 = _small_training_problem()
n_tr = 5
t0 = time.time()
out = ctx.train_nuts(qg32, ql32, num_warmup=0, num_samples=n_tr, max_tree_depth=5, init_step_size=0.004,
                     adapt_step_size=False, adapt_mass_matrix=False, seed=11, passes_per_graph=8, max_passes=4000,
                     progress=lambda p, s: print('passes', p, s[:, :3].tolist()) if p % 64 < 8 else None)
print('finished', out['finished'], 'passes', out['passes'], time.time() - t0)
st = out['stats'].cpu().numpy()
print('gpu stats', st)
qg, ql = qg32.cpu().numpy().astype(np.float64), ql32.cpu().numpy().astype(np.float64)
G = qg.shape[1]
sg = out['samples_g'].cpu().numpy()
for c in range(2):
    rec = ref_fit.train_cpu(ref, case['obs'], case['weights'], qg[c], ql[:, c], chain=c, n_chains=2, num_samples=n_tr, max_depth=5, init_step=0.004, seed=11)
    print('cpu chain', c, 'n_steps', rec['n_steps'], 'accept', np.round(rec['accept'], 4), 'pe', rec['pe'])
    z = np.array(rec['z'])
    print('  max |dz globals| per transition', np.abs(sg[c] - z[:, :G]).max(axis=1))
