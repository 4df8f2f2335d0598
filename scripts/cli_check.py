"""End-to-end check of ``python -m bayesn_b200``: simulate a handful of T21 SNe into SNANA text files, write an input
YAML, run the module entry (also under torchrun), list the output files.  usage: python scripts/cli_check.py outdir"""
import os, subprocess, sys, numpy as np, yaml
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel, snana
out = sys.argv[1]
d = os.path.join(out, 'SIMFIT')
os.makedirs(d, exist_ok=True)
m = SEDmodel(load_model='T21_model')
np.random.seed(3)
N, bands = 6, ['g_PS1', 'r_PS1', 'i_PS1', 'z_PS1']
z = np.random.uniform(0.02, 0.06, N)
t = np.arange(-8, 36, 4.0)
flux, err, par = m.simulate_light_curve(t, N, bands, yerr=0.03, err_type='mag', z=z, mu='z', ebv_mw=0.01, mag=False)
files = []
for i in range(N):
    flt = np.repeat(bands, len(t))
    snana.write_snana_fluxfile(d, f'sn{i}', 58000.0 + np.tile(t, len(bands)) * (1 + z[i]), flt, flux[:, i], err[:, i], 58000.0, z[i], z[i],
                               ebv_mw=0.01)
    files.append(f'sn{i}.DAT')
open(os.path.join(d, 'SIMFIT.LIST'), 'w').write('\n'.join(files) + '\n')
inp = os.path.join(out, 'input.yaml')
yaml.safe_dump(dict(mode='fitting', version_photometry=d, outputdir=os.path.join(out, 'res'), num_warmup=60, num_samples=40), open(inp, 'w'))
r = subprocess.run([sys.executable, '-m', 'bayesn_b200', inp, '--num_chains', '2', '--outfile_prefix', 'cli'], capture_output=True, text=True)
print('rc', r.returncode, r.stderr[-400:] if r.returncode else '')
print(sorted(os.listdir(os.path.join(out, 'res'))))
print(open(os.path.join(out, 'res', 'cli.FITRES.TEXT')).read()[:600])
