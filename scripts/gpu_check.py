"""Ad-hoc GPU check: kernel log p / grad vs the fp64 oracle on simulated data."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel
from bayesn_b200 import native
from oracle.ref_model import RefSED

def to_kernel_layout(q_or, n_sc):
    return np.concatenate([q_or[..., n_sc:], q_or[..., :n_sc]], axis=-1)

def run(model_name, bands, t, N=8, C=3, rv=None, seed=1):
    np.random.seed(seed)
    ref = RefSED(model_name, bands=bands)
    z = np.random.uniform(0.015, 0.08, N)
    data, yerr, pd, (tt, bidx, bw) = ref.simulate_light_curve(t, N, bands, yerr=0.05, z=z, mu='z', mag=False, RV=(None if ref.model_type=='fixed_RV' else ref.mu_R))
    N_obs = data.shape[0]
    obs = np.zeros((10, N_obs, N)); obs[0]=tt; obs[1]=data; obs[2]=yerr; obs[4]=bidx; obs[5]=z[None]; obs[6]=1e-4
    obs[7]=ref.distmod(z)[None]; obs[9]=1
    # mask a few observations to exercise the compaction
    obs[9, ::7, ::2] = 0
    m = SEDmodel(load_model=model_name)
    band_inds = np.array([0] + [m.band_dict[ref.inv_band_dict[k]] for k in range(1, len(bands) + 1)])
    w_prod = m._calculate_band_weights(z, np.zeros(N), band_inds)
    print(model_name, 'band weights prod vs oracle max rel', np.max(np.abs(w_prod - bw) / (np.abs(bw).max())))
    rv_mode = {None: native.RV_FIXED if ref.model_type == 'fixed_RV' else native.RV_POP, 'uniform': native.RV_UNIFORM}[rv]
    ctx = m.prepare(obs, w_prod, band_inds, rv_mode)
    D = ctx.D; ne = m.n_eps; n_sc = D - ne
    q0 = m.init_to_median(obs[7,0,:], C, rv_mode, seed=3).cpu().numpy().astype(np.float64)
    q0 += 0.3*np.random.randn(*q0.shape)
    q32 = torch.tensor(q0, dtype=torch.float32, device='cuda').contiguous()
    lp, g = ctx.logp_grad(q32); torch.cuda.synchronize()
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    qk = q32.cpu().numpy().astype(np.float64)
    q_or = np.concatenate([qk[..., ne:], qk[..., :ne]], axis=-1)
    mode = {native.RV_FIXED:'fixed', native.RV_POP:'pop', native.RV_UNIFORM:'uniform'}[rv_mode]
    worst_lp = worst_g = 0
    for c in range(C):
        lpo, go = ref.fit_logp_grad(q_or[:, c], obs, bw, rv_mode=mode)
        go = np.concatenate([go[:, n_sc:], go[:, :n_sc]], axis=-1)
        rl = np.abs(lp[:, c] - lpo) / np.abs(lpo)
        rg = np.max(np.abs(g[:, c] - go), axis=1) / np.max(np.abs(go), axis=1)
        worst_lp = max(worst_lp, rl.max()); worst_g = max(worst_g, rg.max())
    print(f'{model_name} rv={mode} N_obs={N_obs} D={D}: logp rel {worst_lp:.2e}  grad rel(inf-norm) {worst_g:.2e}')
    return ctx, m

t12 = np.arange(-8, 40, 4.0)
run('T21_model', ['g_PS1','r_PS1','i_PS1','z_PS1'], t12)
run('W22_model', ['u_LSST','g_PS1','r_PS1','i_PS1','z_PS1','Y_WFCAM','J_WFCAM','H_WFCAM'], np.array([-8.,2,12,22,32]))
run('M20_model', ['B_CSP','V_CSP','r_CSP','i_CSP','Y_RC','J_RC1','H_RC'], t12, C=1)
run('pop_RV_model', ['g_PS1','r_PS1','i_PS1','z_PS1'], t12, C=4)
run('T21_model', ['g_PS1','r_PS1','i_PS1','z_PS1'], t12, rv='uniform', C=2)
run('G25_85day_model', ['g_PS1','r_PS1','i_PS1','z_PS1'], t12, C=4)

# ---- short NUTS run on SN 2016W -------------------------------------------------------------
import time
m = SEDmodel(load_model='T21_model')
t0 = time.time()
samples, sn_props = m.fit_from_file(m.example_lc, filt_map={'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'})
print('fit wall', time.time() - t0)
for k in ['AV', 'Ds', 'theta', 'tmax', 'mu']:
    print(k, samples[k].mean(), samples[k].std())
