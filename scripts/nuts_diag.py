import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bayesn_b200 import SEDmodel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m = SEDmodel(load_model='W22_model')
obs, w, bi = bench.build_workload(m, n, 20241019)
ctx = m.prepare(obs, w, bi)
q0 = ctx.init_to_median(4, seed=1)
torch.cuda.synchronize(); t0 = time.time()
out = ctx.nuts(q0, 250, 250, 10, seed=1)
torch.cuda.synchronize(); dt = time.time() - t0
ci = out['chain_info'].cpu().numpy(); st = out['stats'].cpu().numpy()
tot = ci[..., 2]
print('time', dt, 'total evals', tot.sum(), 'evals/s', tot.sum() / dt)
print('per-chain total steps: mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f' % (tot.mean(), np.median(tot), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
samp = st[..., 1]
print('sampling-phase steps/iter: mean %.1f median %.1f p99 %.1f max %.0f' % (samp.mean(), np.median(samp), np.percentile(samp, 99), samp.max()))
print('warmup steps per chain mean', (tot - samp.sum(-1)).mean(), ' sampling per chain mean', samp.sum(-1).mean())
print('step size: median %.3f min %.4f' % (np.median(ci[..., 0]), ci[..., 0].min()), 'mean accept', ci[..., 1].mean(), 'div', ci[..., 3].sum())
cta = tot.max(axis=1)
print('per-SN max-chain steps: mean %.0f max %.0f ; sum of CTA-max / sum %.2f' % (cta.mean(), cta.max(), cta.sum() * 4 / tot.sum()))
np.savez('gpurun_out/nuts_diag.npz', ci=ci, steps=samp)
i = np.unravel_index(np.argmax(tot), tot.shape)
print('slowest chain', i, ci[i][:4], 'its SN other chains', tot[i[0]])
lp0, g0 = ctx.logp_grad(q0)
print('init logp of that SN', lp0[i[0]].cpu().numpy(), 'finite grads', torch.isfinite(g0[i[0]]).all().item())
print('non-finite init logp count', (~torch.isfinite(lp0)).sum().item(), 'non-finite grad rows', (~torch.isfinite(g0).all(-1)).sum().item())
s_ = out['samples'][i[0], i[1]].cpu().numpy()
print('samples of slow chain: first', s_[0][-4:], 'last', s_[-1][-4:], 'accept', st[i[0], i[1], :5, 0], 'pe', st[i[0], i[1], :3, 3])
print('init q scalars', q0[i[0], i[1], -4:].cpu().numpy(), 'obs t', obs[0, :5, i[0]], 'z', obs[5, 0, i[0]])
print('inv mass of slow chain', ci[i][4:][-6:], 'normal chain', ci[i[0], (i[1] + 1) % 4][4:][-6:])
