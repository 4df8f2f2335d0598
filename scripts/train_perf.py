"""C4-shaped training run (M20, 500 SNe x 84 obs, 4 chains): time per pass (= one leapfrog of every chain)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel
from bayesn_b200 import train as btrain
sys.argv = [a for a in sys.argv if a != '--eager'] + (['--eager'] if '--eager' in sys.argv else [])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
W = int(sys.argv[2]) if len(sys.argv) > 2 else 30
S = int(sys.argv[3]) if len(sys.argv) > 3 else 20
m = SEDmodel(load_model='M20_model')
obs, weights, band_inds = btrain.simulate_training_set(m, N, seed=20241019)
ctx = m.prepare_training(obs, weights, band_inds)
qg0, ql0 = btrain.initial_position(m, obs, num_chains=4, reference_model='M20_model', seed=1, device='cuda')
graph = '--eager' not in sys.argv
for ppg in (32,):
    torch.cuda.synchronize(); t0 = time.time()
    out = ctx.train_nuts(qg0, ql0, num_warmup=W, num_samples=S, max_tree_depth=10, seed=2, passes_per_graph=ppg, use_graph=graph)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = out['stats'].cpu().numpy()
    print(f'N={N} {W}+{S}: {dt:.2f} s, passes {out["passes"]}, {dt / out["passes"] * 1e6:.1f} us/pass, leapfrogs/chain {out["info"][:, 4]}, '
          f'accept(sampling) {st[:, W:, 0].mean():.3f}, div {out["info"][:, 5]}, step {out["info"][:, 2]}')
    print('n_steps per transition chain0', st[0, :, 1].astype(int).tolist())
