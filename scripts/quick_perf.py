import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bayesn_b200 import SEDmodel
m = SEDmodel(load_model='W22_model')
obs, w, bi = bench.build_workload(m, 1000, 20241019)
ctx = m.prepare(obs, w, bi)
bufs = {}
for rep in range(2):
    q0 = ctx.init_to_median(4, seed=rep)
    torch.cuda.synchronize(); t0 = time.time()
    out = ctx.nuts(q0, 250, 250, 10, seed=rep, out=bufs)
    torch.cuda.synchronize(); dt = time.time() - t0
    tot = float(out['chain_info'][..., 2].sum())
    print(f'NUTS C3: {dt:.3f} s  {1000/dt:.1f} SNe/s  {tot/dt/1e6:.2f} M evals/s  div {float(out["chain_info"][...,3].sum()):.0f}')
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
