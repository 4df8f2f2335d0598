"""Config C5: LSST-scale W22 fit, ugrizy, 10 epochs -> 60 obs, z ~ U(0.02, 0.09);
one GPU's share of the 100,000-SN job.  Data simulated with simulate_light_curve in chunks; band weights
built on the device."""
import sys, time, json, numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
bands = ['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST']
m = SEDmodel(load_model='W22_model')
np.random.seed(20241019)
obs_l, t0 = [], time.perf_counter()
for s0 in range(0, N, 4000):
    n = min(4000, N - s0)
    z = np.random.uniform(0.02, 0.09, n)
    data, yerr, params = m.simulate_light_curve(np.arange(-8, 40, 5.0), n, bands, yerr=0.05, err_type='mag', z=z, mu='z', ebv_mw=0, mag=False)
    o, w, bi = m.simulated_obs_array(data, yerr, params)
    obs_l.append(o)
obs = np.concatenate(obs_l, axis=2)
t_sim = time.perf_counter() - t0
t0 = time.perf_counter()
ctx = m.prepare(obs, None, bi)              # device tile builder
torch.cuda.synchronize(); t_prep = time.perf_counter() - t0
q0 = ctx.init_to_median(4, seed=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = ctx.nuts(q0, 250, 250, 10, seed=0)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
lf = float(out['chain_info'][..., 2].sum())
F = bench.flops_per_eval(11, 6, 60)
res = dict(config=f'C5 share: W22 ugrizy, {N} SNe x 60 obs, 4 x (250+250)', sne_per_s=N / dt, seconds=dt, evals_per_s=lf / dt,
           leapfrogs_per_sn=lf / N, algorithmic_tflops=F * lf / dt / 1e12, prepare_s=t_prep, simulate_s=t_sim,
           divergences=float(out['chain_info'][..., 3].sum()), tile_floats_per_sn=int(ctx._keep['data']['wt_stride']))
ci = out['chain_info'].cpu().numpy()
lfc = ci[..., 2]
res['leapfrogs_per_chain'] = dict(mean=float(lfc.mean()), p50=float(np.median(lfc)), p99=float(np.percentile(lfc, 99)),
                                  max=float(lfc.max()), argmax=[int(i) for i in np.unravel_index(lfc.argmax(), lfc.shape)],
                                  top5=[float(x) for x in np.sort(lfc.ravel())[-5:]],
                                  step_of_max=float(ci[..., 0].ravel()[lfc.argmax()]), step_median=float(np.median(ci[..., 0])))
print(json.dumps(res))
