"""Config C1: T21, 10 simulated SNe (griz, 48 obs), per-object fits 4 x (250+250)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel
d = np.load('tests/golden/c1_oracle_posterior.npz')
m = SEDmodel(load_model='T21_model')
bi = np.array([m.band_dict[str(n)] for n in d['band_names']])
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cons, out = m.fit_batch(d['obs'], None, bi, seed=rep, return_device=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    lf = float(out['chain_info'][..., 2].sum())
    print(f'C1: 10 SNe fitted in {dt:.3f} s (incl. packing on the host + device tile build), {lf:.0f} leapfrogs; oracle sampler: {float(d["seconds"]):.0f} s on 8 cores, {int(d["evals"])} evaluations')
