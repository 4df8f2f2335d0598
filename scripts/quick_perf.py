"""Quick perf probe (GPU box): NUTS time vs N, chain imbalance, logp+grad throughput."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bayesn_b200 import SEDmodel
m = SEDmodel(load_model='W22_model')
Ns = [int(a) for a in sys.argv[1:]] or [1000]
obs, w, bi = bench.build_workload(m, max(Ns), 20241019)
for N in Ns:
    ctx = m.prepare(obs[:, :, :N], w[:N], bi)
    bufs = {}
    for rep in range(2):
        q0 = ctx.init_to_median(4, seed=rep)
        torch.cuda.synchronize(); t0 = time.time()
        out = ctx.nuts(q0, 250, 250, 10, seed=rep, out=bufs)
        torch.cuda.synchronize(); dt = time.time() - t0
        lf = out['chain_info'][..., 2].cpu().numpy()
        tot = float(lf.sum())
        print(f'NUTS N={N}: {dt:.3f} s  {N/dt:.1f} SNe/s  {tot/dt/1e6:.2f} M evals/s  div {float(out["chain_info"][...,3].sum()):.0f}'
              f'  leapfrogs/chain mean {lf.mean():.0f} min {lf.min():.0f} max {lf.max():.0f}  mean(max over chains of SN) {lf.max(axis=1).mean():.0f}', flush=True)
obs2, w2, bi2 = bench.build_workload(m, 16384, 777)
ctx2 = m.prepare(obs2, w2, bi2)
for C in (1, 4):
    q = ctx2.init_to_median(C, seed=5)
    for _ in range(3): ctx2.logp_grad(q)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx2.logp_grad(q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'logp_grad N=16384 C={C}: {ms:.3f} ms  {16384*C/ms/1e3:.2f} M evals/s')
