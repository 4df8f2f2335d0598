"""C3 as BASELINE.json states it - 1000 simulated W22 SNe in total, sharded over 1/2/4/8 GPUs - emulated on ONE GPU:
the fit mode has no data-path collective, so the N-GPU time is the time of the largest shard alone on a GPU.  For every
shard size the sampler is timed in the throughput layout and in latency mode (warps_per_chain 2, 4, 8, auto).
usage: python scripts/strong_scaling.py [n_total] [reps]"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bayesn_b200 import SEDmodel
from bayesn_b200.distributed import shard_bounds
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = SEDmodel(load_model=bench.MODEL)
obs, weights, band_inds = bench.build_workload(m, n_total, 20241019)
rows = []
t1 = None
for world in (1, 2, 4, 8):
    lo, hi = shard_bounds(n_total, 0, world)
    ctx = m.prepare(obs[:, :, lo:hi], None, band_inds)
    q0 = ctx.init_to_median(4, seed=0, sn_offset=lo)
    for K in (1, 2, 4, 8):
        if (hi - lo) * 4 * K > 2 * ctx.resident_warps():
            continue
        bufs, ms = {}, []
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ctx.nuts(q0, 250, 250, 10, seed=0, sn_offset=lo, warps_per_chain=K, out=bufs)
            e1.record(); torch.cuda.synchronize()
            if r:
                ms.append(e0.elapsed_time(e1))
        lf = out['chain_info'][..., 2]
        row = dict(world=world, shard_sne=hi - lo, warps_per_chain=K, auto=ctx.auto_warps_per_chain((hi - lo) * 4), seconds=float(np.mean(ms)) / 1e3,
                   leapfrogs=float(lf.sum()), longest_chain=float(lf.max()), mean_chain=float(lf.mean()),
                   us_per_leapfrog_longest=float(np.mean(ms)) * 1e3 / float(lf.max()))
        if world == 1 and K == 1:
            t1 = row['seconds']
        row['sne_per_s_job'] = n_total / row['seconds']
        row['strong_efficiency'] = t1 / (world * row['seconds'])
        rows.append(row)
        print(json.dumps(row), flush=True)
print(json.dumps({'config': f'C3 strong scaling emulated on one GPU: {n_total} SNe total, shard 0 of N', 'rows': rows}))
