import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bayesn_b200 import SEDmodel
m = SEDmodel(load_model='W22_model')
obs, w, bi = bench.build_workload(m, 1000, 20241019)
ctx = m.prepare(obs, w, bi)
q0 = ctx.init_to_median(4, seed=1)
out = ctx.nuts(q0, 250, 250, 10, seed=1)
torch.cuda.synchronize()
st = out['stats'].cpu().numpy()
sn, ch = 522, 2
np.set_printoptions(linewidth=200, precision=4, suppress=False)
print('accept', st[sn, ch, :40, 0]); print('nsteps', st[sn, ch, :40, 1]); print('div', st[sn, ch, :40, 2]); print('pe', st[sn, ch, :40, 3]); print('step', st[sn, ch, :40, 4])
print('normal chain: accept', st[sn, 0, :20, 0]); print('nsteps', st[sn, 0, :20, 1]); print('pe', st[sn, 0, :20, 3]); print('step', st[sn, 0, :20, 4])
np.savez('gpurun_out/slow_chain.npz', obs=obs[:, :, sn:sn+1], w=w[sn:sn+1], q0=q0[sn].cpu().numpy(), st=st[sn])
