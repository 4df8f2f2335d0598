import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel, native, train as btrain
m = SEDmodel(load_model='M20_model')
obs, weights, band_inds = btrain.simulate_training_set(m, 500, seed=20241019)
ctx = m.prepare_training(obs, None, band_inds)
qg0, ql0 = btrain.initial_position(m, obs, num_chains=4, reference_model='M20_model', seed=1, device='cuda')
acc = np.zeros(9)
n = 0
def prog(p, st):
    global acc, n
    torch.cuda.synchronize()
    t = (C.c_longlong * 16)()
    native.load_library().bayesn_debug_ctl_times(t)
    a = np.array(list(t)[:9], dtype=np.float64)
    acc += np.r_[np.diff(a), a[-1] - a[0]]
    n += 1
out = ctx.train_nuts(qg0, ql0, num_warmup=12, num_samples=4, max_tree_depth=10, seed=2, use_graph=False, passes_per_graph=8, progress=prog)
names = ['load ctl', 'finish_dev(chain rule)', 'step2 loop', 'block_sums', 'decisions(t0)', 'actions', 'step5 pre', 'globals_dev', 'TOTAL']
for k, v in zip(names, acc / n / 1.965e3):
    print(f'{k:26s} {v:8.2f} us')
