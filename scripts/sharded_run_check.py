"""Sharded product entry under torchrun (one process per GPU): SEDmodel.run(mode='fitting') on a small SNANA text
directory with the SNe sharded over the ranks against the same run on one GPU - FITRES rows, LCPLOT and chains.pkl
must agree (draws bit for bit: the streams are keyed by the global SN index).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
        scripts/sharded_run_check.py /tmp/sharded_check"""
import os, pickle, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bayesn_b200 import SEDmodel, snana

root = sys.argv[1] if len(sys.argv) > 1 else '/tmp/sharded_check'
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
PS1 = ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1']
m = SEDmodel(load_model='T21_model', device=f'cuda:{local}')
N = 11
d = os.path.join(root, 'SIMSHARD')
if rank == 0:
    os.makedirs(d, exist_ok=True)
    np.random.seed(3)
    z = np.random.uniform(0.02, 0.06, N)
    tgrid = np.arange(-8, 36, 4.0)
    flux, err, par = m.simulate_light_curve(tgrid, N, PS1, yerr=0.03, err_type='mag', z=z, mu='z', ebv_mw=0.01, mag=False)
    bands = np.repeat(['g', 'r', 'i', 'z'], len(tgrid))
    tt = np.tile(tgrid, 4)
    files = []
    for i in range(N):
        snana.write_snana_fluxfile(d, f'{3000 + i}', 58000.0 + tt * (1 + z[i]), bands, flux[:, i], err[:, i], 58000.0, z[i], z[i], ebv_mw=0.01)
        files.append(f'{3000 + i}.DAT')
    open(os.path.join(d, 'SIMSHARD.LIST'), 'w').write('\n'.join(files) + '\n')
dist.barrier()
base = dict(mode='fitting', version_photometry=d, outfile_prefix='fit', num_warmup=80, num_samples=50,
            map={'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'}, warps_per_chain=2)
s_sh = m.run(dict(base, outputdir=os.path.join(root, 'sharded')))
dist.barrier()
ok = True
if rank == 0:
    s_1 = m.run(dict(base, outputdir=os.path.join(root, 'single'), distributed=False))
    for k in ('theta', 'AV', 'Ds', 'tmax', 'eps', 'mu', 'delM', 'peak_MJD'):
        same = np.array_equal(s_sh[k], s_1[k])
        ok &= same
        print(f'{k:9s} sharded == single GPU: {same}  shape {s_sh[k].shape}')
    rows = lambda p: [l.split() for l in open(p).read().splitlines() if l.startswith('SN:')]
    a, b = rows(os.path.join(root, 'sharded', 'fit.FITRES.TEXT')), rows(os.path.join(root, 'single', 'fit.FITRES.TEXT'))
    def num(rows_):            # numeric cells only (FIELD / IDSURVEY may be strings)
        out = []
        for r in rows_:
            vals = []
            for x in r[2:]:
                try:
                    vals.append(float(x))
                except ValueError:
                    vals.append(float(sum(map(ord, x))))
            out.append(vals)
        return np.array(out)
    fa, fb = num(a), num(b)
    print('FITRES rows', len(a), len(b), 'max |diff|', np.abs(fa - fb).max())
    ok &= len(a) == N and np.abs(fa - fb).max() < 2e-4
    ca, cb = pickle.load(open(os.path.join(root, 'sharded', 'chains.pkl'), 'rb')), pickle.load(open(os.path.join(root, 'single', 'chains.pkl'), 'rb'))
    ok &= all(np.array_equal(ca[k], cb[k]) for k in cb)
    la, lb = open(os.path.join(root, 'sharded', 'fit.LCPLOT')).read(), open(os.path.join(root, 'single', 'fit.LCPLOT')).read()
    ok &= len(la.splitlines()) == len(lb.splitlines())
    print('SHARDED_RUN_CHECK', 'OK' if ok else 'FAILED', f'world={world}')
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
