"""Time of the device-side posterior summaries (FITRES columns incl. the rank-normalised R-hat of every site) for one
GPU's share of C5: N SNe x 4 chains x 250 draws, synthetic autocorrelated draws.  usage: python scripts/summary_perf.py [N]"""
import sys, time, json, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel, outputs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
m = SEDmodel(load_model='W22_model')
g = torch.Generator(device='cuda').manual_seed(1)
q = torch.randn(N, 4, 250, m.n_eps + 4, device='cuda', generator=g)
for i in range(1, 250):
    q[:, :, i] = 0.5 * q[:, :, i - 1] + q[:, :, i]
cons = m.constrain(q)
muhat = 35.0 + torch.rand(N, device='cuda', generator=g).double().cpu().numpy()
torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats(); base = torch.cuda.memory_allocated()
t0 = time.perf_counter()
summ = outputs.device_fit_summary(m, cons, muhat)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps(dict(N=N, rows_per_sn=int(8 + m.n_eps), seconds=dt, peak_extra_gb=(torch.cuda.max_memory_allocated() - base) / 2 ** 30,
                      meanrhat_p50=float(sorted(summ['MEANRHAT'])[N // 2]), maxrhat_max=float(summ['MAXRHAT'].max()))))
