"""Config C5 for real (BASELINE.json configs[4]): LSST-scale W22 fit of 100 000 simulated SNe (ugrizy, 10 epochs -> 60
observations each) across the GPUs of one box, from simulate_light_curve inputs to FITRES.TEXT, through the product's
sharded entry (driver.fit_dataset: NUTS fits, device-side summaries, only FITRES rows gathered to rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/c5_full.py 100000 gpurun_out/c5_full
Every rank simulates ITS OWN block of SNe (rank-seeded) - the data are synthetic either way."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from bayesn_b200 import SEDmodel, driver
from bayesn_b200.distributed import shard_bounds

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
outdir = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/c5_full'
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
group = None
if world > 1:
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    group = dist.group.WORLD
bands = ['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST']
m = SEDmodel(load_model='W22_model', device=f'cuda:{local}')
lo, hi = shard_bounds(n_total, rank, world)
t_all = time.perf_counter()
np.random.seed(20241019 + rank)
obs_l, truth = [], []
for s0 in range(lo, hi, 4000):
    n = min(4000, hi - s0)
    z = np.random.uniform(0.02, 0.09, n)
    data, yerr, params = m.simulate_light_curve(np.arange(-8, 40, 5.0), n, bands, yerr=0.05, err_type='mag', z=z, mu='z', ebv_mw=0, mag=False)
    o, w, bi = m.simulated_obs_array(data, yerr, params)
    obs_l.append(o)
    truth.append(params['mu'] + params['del_M'])
obs = np.concatenate(obs_l, axis=2)
truth = np.concatenate(truth)
t_sim = time.perf_counter() - t_all
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
t0 = time.perf_counter()
ds = dict(obs=obs, band_inds=bi, shard=(lo, hi, n_total), sn_names=None, peak_mjds=None)
args = dict(num_chains=4, num_warmup=250, num_samples=250, outputdir=outdir, outfile_prefix='c5', save_chains=False,
            return_samples=False, lcplot=False, warps_per_chain=1, seed=0)
os.makedirs(outdir, exist_ok=True)
samples, dropped = driver.fit_dataset(m, ds, args, group)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
tm = ds['timing']
mine = torch.tensor([tm['fit_batch_s'], tm['leapfrogs'], tm['divergences'], torch.cuda.max_memory_allocated() / 2 ** 30, t_sim],
                    dtype=torch.float64, device=f'cuda:{local}')
allr = [torch.empty_like(mine) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, mine)
else:
    allr = [mine]
# check the recovered distances of this rank's block against the simulated truth
if rank == 0:
    cols = ds['fitres_columns']
    per = [dict(rank=r, fit_batch_s=float(a[0]), leapfrogs=float(a[1]), divergences=float(a[2]), peak_mem_gib=float(a[3]), simulate_s=float(a[4]),
                sne=shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]) for r, a in enumerate(allr)]
    for p in per:
        p['sne_per_s'] = p['sne'] / p['fit_batch_s']
    mu_fit, mu_err = cols['MU_LCFIT'][lo:hi], cols['MUERR_LCFIT'][lo:hi]
    zsc = (mu_fit - truth) / np.maximum(mu_err, 1e-3)
    F = bench.flops_per_eval(11, 6, 60)
    lf = sum(p['leapfrogs'] for p in per)
    res = dict(config=f'C5: W22 LSST ugrizy, {n_total} SNe x 60 obs, 4 x (250+250), {world} GPU(s), simulate_light_curve -> FITRES.TEXT',
               wall_s_fit_to_fitres=dt, sne_per_s_job=n_total / dt, per_rank=per, fitres_rows=int(len(cols['MU_LCFIT']) - dropped),
               fitres_bytes=os.path.getsize(os.path.join(outdir, 'c5.FITRES.TEXT')), leapfrogs=lf,
               algorithmic_tflops_per_gpu=F * lf / world / max(p['fit_batch_s'] for p in per) / 1e12,
               maxrhat_p50=float(np.median(cols['MAXRHAT'])), maxrhat_p99=float(np.percentile(cols['MAXRHAT'], 99)),
               rank0_mu_zscore_std=float(zsc.std()), rank0_mu_abs_err_p50=float(np.median(np.abs(mu_fit - truth))),
               total_divergences=sum(p['divergences'] for p in per))
    print(json.dumps(res))
    with open(os.path.join(outdir, 'c5_full.json'), 'w') as fh:
        json.dump(res, fh, indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
