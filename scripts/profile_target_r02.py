"""Short, deterministic workloads for ncu (round 2): the persistent sampler kernel on
  (a) the C3 shape, 1184 SNe x 4 chains, one warp per chain (work-queue schedule)      -> launch 1
  (b) the C5 shape (LSST ugrizy, 60 obs), 592 SNe x 4 chains, static schedule            -> launch 2
  (c) the C3 shape, 125 SNe x 4 chains in latency mode, 4 warps per chain                -> launch 3
plus the one-launch get_flux_from_chains kernel.  20 + 10 transitions, tree depth <= 6."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from bayesn_b200 import SEDmodel

m = SEDmodel(load_model='W22_model')
obs, w, bi = bench.build_workload(m, 1184, 777)
ctx = m.prepare(obs, None, bi)
q4 = ctx.init_to_median(4, seed=6)
out = ctx.nuts(q4, num_warmup=20, num_samples=10, max_tree_depth=6, seed=1)
torch.cuda.synchronize()
print('C3 K=1 leapfrogs', float(out['chain_info'][..., 2].sum()))
lsst = ['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST']
np.random.seed(3)
z = np.random.uniform(0.02, 0.09, 592)
data, yerr, params = m.simulate_light_curve(np.arange(-8, 40, 5.0), 592, lsst, yerr=0.05, err_type='mag', z=z, mu='z', ebv_mw=0, mag=False)
o5, w5, b5 = m.simulated_obs_array(data, yerr, params)
c5 = m.prepare(o5, None, b5)
q5 = c5.init_to_median(4, seed=6)
out5 = c5.nuts(q5, num_warmup=20, num_samples=10, max_tree_depth=6, seed=1)
torch.cuda.synchronize()
print('C5 K=1 leapfrogs', float(out5['chain_info'][..., 2].sum()))
ck = m.prepare(obs[:, :, :125], None, bi)
qk = ck.init_to_median(4, seed=6)
outk = ck.nuts(qk, num_warmup=20, num_samples=10, max_tree_depth=6, seed=1, warps_per_chain=4)
torch.cuda.synchronize()
print('C3 K=4 leapfrogs', float(outk['chain_info'][..., 2].sum()))
cons = m.constrain(out['samples'][:64])
from bayesn_b200 import outputs
outputs.draw_mu_device(m, cons, obs[7, 0, :64])
fl = lambda v: v.transpose(1, 2).reshape(64, -1)
f = m.chain_flux_device(np.arange(-10, 40, 2.0), bench.BANDS, fl(cons['theta']), fl(cons['AV']), fl(cons['tmax']), fl(cons['mu']) + fl(cons['delM']),
                        cons['eps'].transpose(1, 2).reshape(64, -1, cons['eps'].shape[-1]).contiguous(), obs[5, 0, :64], obs[8, 0, :64])
torch.cuda.synchronize()
print('chain flux', tuple(f.shape), float(f.mean()))
