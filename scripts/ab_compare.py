"""Same-box A/B of library builds on the headline shapes (box-to-box spread is +-2 %, more than most changes):
usage: python scripts/ab_compare.py [--c5] name=path/to/lib.so name2=path2.so ...   (runs A B A B, subprocess per run)"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == '--one':
    sys.path.insert(0, root)
    import numpy as np, torch
    import bench
    from bayesn_b200 import SEDmodel
    c5 = '--c5' in sys.argv
    m = SEDmodel(load_model='W22_model')
    if c5:
        np.random.seed(20241019)
        n = 1184
        z = np.random.uniform(0.02, 0.09, n)
        data, yerr, params = m.simulate_light_curve(np.arange(-8, 40, 5.0), n, ['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST'],
                                                    yerr=0.05, err_type='mag', z=z, mu='z', ebv_mw=0, mag=False)
        obs, w, bi = m.simulated_obs_array(data, yerr, params)
    else:
        n = 1000
        obs, w, bi = bench.build_workload(m, n, 20241019)
    ctx = m.prepare(obs, w, bi)
    q0 = ctx.init_to_median(4, seed=0)
    bufs, ms = {}, []
    for r in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ctx.nuts(q0, 250, 250, 10, seed=0, out=bufs)
        e1.record(); torch.cuda.synchronize()
        if r:
            ms.append(e0.elapsed_time(e1))
    lf = float(out['chain_info'][..., 2].sum())
    res = dict(nuts_s=float(np.mean(ms)) / 1e3, sne_per_s=n / (float(np.mean(ms)) / 1e3), evals_per_s=lf / (float(np.mean(ms)) / 1e3))
    if not c5:
        obs2, w2, bi2 = bench.build_workload(m, 16384, 777)
        c2 = m.prepare(obs2, w2, bi2)
        for C in (1, 4):
            q = c2.init_to_median(C, seed=5)
            for _ in range(3):
                c2.logp_grad(q)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                c2.logp_grad(q)
            e1.record(); torch.cuda.synchronize()
            res[f'logp_grad_C{C}_Mevals'] = 16384 * C / (e0.elapsed_time(e1) / 10) / 1e3
    print('AB_RESULT ' + json.dumps(res))
    sys.exit(0)
flags = [a for a in sys.argv[1:] if a.startswith('--')]
libs = [a.split('=', 1) for a in sys.argv[1:] if '=' in a]
rows = {k: [] for k, _ in libs}
for rep in range(2):
    for name, path in libs:
        env = dict(os.environ, BAYESN_B200_LIB=os.path.abspath(path))
        p = subprocess.run([sys.executable, __file__, '--one'] + flags, env=env, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith('AB_RESULT ')]
        if not line:
            print(name, 'FAILED', p.stderr[-800:])
            continue
        rows[name].append(json.loads(line[0][10:]))
        print(name, rows[name][-1], flush=True)
print(json.dumps(rows))
