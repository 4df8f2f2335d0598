"""Short, deterministic workload for ncu: the fused log p + gradient kernel at C2 scale, the same
kernel with 4 chains per SN (TMA-staged weight tiles), and a short NUTS run (work-queue schedule)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from bayesn_b200 import SEDmodel

m = SEDmodel(load_model='W22_model')
obs, w, bi = bench.build_workload(m, 16384, 777)
ctx = m.prepare(obs, w, bi)
q1 = ctx.init_to_median(1, seed=5)
for _ in range(3):
    ctx.logp_grad(q1)
torch.cuda.synchronize()
# 1184 SNe x 4 chains = two waves of (SN, chain) units: the sampler runs its work-queue schedule
obs2, w2, bi2 = obs[:, :, :1184], w[:1184], bi
ctx2 = m.prepare(obs2, w2, bi2)
q4 = ctx2.init_to_median(4, seed=6)
for _ in range(3):
    ctx2.logp_grad(q4)
out = ctx2.nuts(q4, num_warmup=20, num_samples=10, max_tree_depth=6, seed=1)
torch.cuda.synchronize()
print('profile target done; leapfrogs', float(out['chain_info'][..., 2].sum()))
