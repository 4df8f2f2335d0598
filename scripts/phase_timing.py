"""Where a single warp's evaluation time goes (latency view).  Builds the -DBAYESN_PHASE_TIMING variant of the
library, runs a short NUTS on few chains (so warp 0 of CTA 0 runs alone on its SM) and prints clock64() cycles per
evaluation and phase, for warps_per_chain = 1, 2, 4, 8.
usage: python scripts/phase_timing.py [c3|c1|c5] [n_sne]"""
import ctypes, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
lib = os.path.join(root, 'bayesn_b200', 'csrc', 'libbayesn_b200_phase.so')
src = os.path.join(root, 'bayesn_b200', 'csrc')
if not os.path.exists(lib) or any(os.path.getmtime(os.path.join(src, f)) > os.path.getmtime(lib) for f in os.listdir(src) if f.endswith(('.cu', '.cuh'))):
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++20', '-Xcompiler', '-fPIC',
                           '-shared', '-DBAYESN_PHASE_TIMING', '-I' + os.path.join(root, 'include'), '-o', lib, os.path.join(src, 'bayesn_capi.cu')])
if '--build-only' in sys.argv:
    sys.exit(0)
os.environ['BAYESN_B200_LIB'] = lib
import numpy as np, torch
import bench
from bayesn_b200 import SEDmodel, native
which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if which == 'c1':
    m = SEDmodel(load_model='T21_model'); bands = ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1']; t = np.arange(-8, 40, 4.0); zr = (0.015, 0.08)
elif which == 'c5':
    m = SEDmodel(load_model='W22_model'); bands = ['u_LSST', 'g_LSST', 'r_LSST', 'i_LSST', 'z_LSST', 'y_LSST']; t = np.arange(-8, 40, 5.0); zr = (0.02, 0.09)
else:
    m = SEDmodel(load_model='W22_model'); bands = bench.BANDS; t = bench.EPOCHS; zr = (0.015, 0.08)
np.random.seed(5)
z = np.random.uniform(zr[0], zr[1], n)
data, yerr, params = m.simulate_light_curve(t, n, bands, yerr=0.05, err_type='mag', z=z, mu='z', ebv_mw=0, mag=False)
obs, w, bi = m.simulated_obs_array(data, yerr, params)
ctx = m.prepare(obs, None, bi)
l = native.load_library()
l.bayesn_debug_phase.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
names = ['scalars', 'W assembly', 'obs setup', 'pair loop', 'combine/exchange', 'priors', 'sampler bookkeeping']
sub = {8: 'eps=L e', 9: 'dtheta', 10: 'L^T dW'}
res = {}
for K in (1, 2, 4, 8):
    q0 = ctx.init_to_median(4, seed=1)
    l.bayesn_debug_phase(None, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ctx.nuts(q0, 60, 40, 8, seed=1, warps_per_chain=K)
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    l.bayesn_debug_phase(buf, 0)
    v = np.array(buf[:16], dtype=np.float64)
    ne = max(v[7], 1)
    lf = float(out['chain_info'][..., 2].sum())
    res[K] = dict(cycles_per_eval={k: round(v[i] / ne, 1) for i, k in enumerate(names)}, total_cycles_per_eval=round((v[:7].sum() + v[8:11].sum()) / ne, 1), sub_phases={k: round(v[i] / ne, 1) for i, k in sub.items()},
                  evals_warp0=int(ne), kernel_ms=e0.elapsed_time(e1), leapfrogs_all=lf)
    print(f'{which} n={n} K={K}: {res[K]["total_cycles_per_eval"]:.0f} cycles/eval  ' +
          '  '.join(f'{k} {v[i] / ne:.0f}' for i, k in enumerate(names)) + '  ' + '  '.join(f'{k} {v[i] / ne:.0f}' for i, k in sub.items()) + f'   kernel {res[K]["kernel_ms"]:.2f} ms, {lf:.0f} leapfrogs')
print(json.dumps({'config': which, 'n_sne': n, 'phases': res}))
