"""Config C4 at its nominal length (BASELINE.json configs[3]): M20 population training (train_model_globalRV: W0, W1,
L_Sigma, sigma0, R_V, tau_A) on 500 SNe simulated from M20 itself, 4 chains x (500 warm-up + 500 draws), max_tree_depth
10, NUTS step_size 0.1 - run once per round; prints the recovered globals with split R-hat next to the truth.
Single GPU: python scripts/c4_nominal.py   |   sharded: torchrun ... scripts/c4_nominal.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesn_b200 import SEDmodel, train as btrain, summary
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
group = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    group = dist.group.WORLD
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
W = S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
m = SEDmodel(load_model='M20_model', device=f'cuda:{local}')
obs, w, bi = btrain.simulate_training_set(m, N, seed=20241019)
torch.cuda.synchronize(); t0 = time.time()
samples, ex = m.train(obs, None, bi, num_chains=4, num_warmup=W, num_samples=S, initialisation='M20_model', seed=2, group=group)
torch.cuda.synchronize(); dt = time.time() - t0
if rank == 0:
    st = ex['stats']
    res = dict(config=f'C4 nominal: M20 train_model_globalRV, {N} simulated SNe x {obs.shape[1]} obs, 4 x ({W}+{S}), {world} GPU(s)',
               seconds=dt, passes=int(ex['passes']), us_per_leapfrog_pass=dt / ex['passes'] * 1e6,
               mean_accept_sampling=float(st[:, W:, 0].mean()), leapfrogs_per_transition_sampling=float(st[:, W:, 1].mean()),
               divergences=float(ex['info'][:, 5].sum()), final_step_sizes=[float(x) for x in ex['info'][:, 2]])
    for k, truth in (('tauA', m.tauA), ('sigma0', m.sigma0), ('RV', m.RV)):
        x = samples[k]
        res[k] = dict(mean=float(x.mean()), sd=float(x.std()), rhat=float(summary.split_rhat(x)), truth=float(truth),
                      z=float((x.mean() - truth) / x.std()))
    W0 = samples['W0']
    zs = (W0.mean(axis=(0, 1)) - m.W0.flatten(order='F')) / W0.std(axis=(0, 1))
    rh = np.array([summary.split_rhat(W0[..., i]) for i in range(W0.shape[-1])])
    res['W0'] = dict(z_rms=float(np.sqrt((zs ** 2).mean())), z_max=float(np.abs(zs).max()), rhat_max=float(rh.max()), rhat_median=float(np.median(rh)))
    print(json.dumps(res))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
