"""``SEDmodel``: host-side mirror of the reference's public class (``bayesn/bayesn_model.py:59``)
for the hot path.  Same constructor, attributes, method names, argument meaning and array
layouts; the arithmetic runs in the sm_100a kernels behind ``libbayesn_b200.so``.

What is kept from the reference API (file:line = ``bayesn/bayesn_model.py``):
``SEDmodel(num_devices, load_model, filter_yaml, fiducial_cosmology)`` :199, ``get_spectra`` :556
(the fit kernels never materialise spectra; this is a separate forward kernel), ``get_flux_batch`` :645, ``get_mag_batch``
:711, ``fit`` :1994, ``fit_from_file`` :1926, ``simulate_light_curve`` :3103, ``sample_*``
:3347-3424, ``get_flux_from_chains`` :3426, plus ``fit_batch`` (the vmapped per-object fit of
``run`` :1809-1849 as a plain method).  numpyro models are replaced by ``logp_grad`` (the
unconstrained log-density + gradient of ``fit_model_globalRV`` / ``popRV`` / ``uniformRV``).
"""
import math
import os
import pickle
import time

import numpy as np

from . import datafiles, native, packing
from .static import FilterSet, FlatLambdaCDM, ModelGrid, spline_row, HSIAO_PHASE0_INDEX
from . import snana, summary


class SEDmodel(object):
    def __init__(self, num_devices=4, load_model='T21_model', filter_yaml=None,
                 fiducial_cosmology={"H0": 73.24, "Om0": 0.28}, device=None, verbose=False):
        self.start_time = time.time()
        self.end_time = None
        self.num_devices = num_devices          # kept for signature compatibility; sharding is per process
        self.__root_dir__ = datafiles.DATA_ROOT
        self.cosmo = FlatLambdaCDM(**fiducial_cosmology)
        self.data = None
        self.hsiao_interp = None
        self.RV_MW = 3.1
        self.sigma_pec = 150 / 3e5
        self.sn_list = None
        self.filter_yaml = filter_yaml
        self.verbose = verbose
        if verbose:
            print(f'Loading model {load_model}')
        params = datafiles.load_model_yaml(load_model)
        self.example_lc = datafiles.EXAMPLE_LC
        self.l_knots = np.array(params['L_KNOTS'], dtype=np.float64)
        self.tau_knots = np.array(params['TAU_KNOTS'], dtype=np.float64)
        self.W0 = np.array(params['W0'], dtype=np.float64)
        self.W1 = np.array(params['W1'], dtype=np.float64)
        self.L_Sigma = np.array(params['L_SIGMA_EPSILON'], dtype=np.float64)
        self.M0 = float(params['M0'])
        self.sigma0 = float(params['SIGMA0'])
        self.tauA = float(params['TAUA'])
        if 'RV' in params:
            self.model_type = 'fixed_RV'
            self.RV = float(params['RV'])
        elif 'MUR' in params:
            self.model_type = 'pop_RV'
            self.mu_R = float(params['MUR'])
            self.sigma_R = float(params['SIGMAR'])
        self.trunc_val = 1.2
        self.used_band_inds = None
        self.band_weights = None
        self._device = device
        self._ctxs = {}
        self._setup_band_weights()

    # ------------------------------------------------------------------ static tables
    def _setup_band_weights(self):
        """Reference :287-497."""
        self.spectrum_bins = 300
        self.band_oversampling = 51
        self.max_redshift = 4
        self.grid = ModelGrid(self.l_knots, self.tau_knots)
        self.filters = FilterSet(self.grid, self.filter_yaml)
        f = self.filters
        self.band_dict, self.zp_dict, self.band_lim_dict = f.band_dict, f.zp_dict, f.band_lim_dict
        self.inv_band_dict = f.inv_band_dict
        self.zps, self.offsets = f.zps, f.offsets
        self.used_band_inds = np.array(list(self.band_dict.values()))
        self.used_band_dict = {v: v for v in self.band_dict.values()}
        self.band_interpolate_weights = f.master
        self.band_interpolate_spacing = self.grid.band_spacing
        self.model_wave = self.grid.model_wave
        self.J_l_T = self.grid.J_l
        self.KD_t = self.grid.KD_t
        self.hsiao_flux = self.grid.hsiao_flux
        self.hsiao_t = self.grid.hsiao_phase
        self.xk = np.array([0.0, 1e4 / 26500., 1e4 / 12200., 1e4 / 6000., 1e4 / 5470., 1e4 / 4670., 1e4 / 4110.,
                            1e4 / 2700., 1e4 / 2600.])
        self.M_fitz_block = self.grid.M_fitz
        self.ZPT = 27.5
        self.sim = False
        self._ctxs = {}

    def _calculate_band_weights(self, redshifts, ebv, band_inds=None):
        """Reference :499-554.  Unlike the reference this does not mutate
        ``band_interpolate_weights``; ``band_inds`` defaults to
        ``self.used_band_inds``."""
        if band_inds is None:
            band_inds = self.used_band_inds
        return self.filters.band_weights(redshifts, ebv, band_inds, RV_MW=self.RV_MW)

    def J_t_map(self, t, tau_knots=None, KD_t=None):
        """Rows of the time-spline matrix (vmapped ``spline_coeffs_irr_step``, reference :258)."""
        return spline_row(t, self.tau_knots if tau_knots is None else tau_knots,
                          self.KD_t if KD_t is None else KD_t)

    # ------------------------------------------------------------------ native context
    def _device_index(self):
        import torch
        if self._device is not None:
            return torch.device(self._device).index or 0
        return int(os.environ.get('LOCAL_RANK', torch.cuda.current_device() if torch.cuda.is_available() else 0))

    def _rv_mode(self, RV=None):
        if RV == 'uniform':
            return native.RV_UNIFORM
        return native.RV_FIXED if self.model_type == 'fixed_RV' else native.RV_POP

    def _context(self, rv_mode, fix_tmax=False, fix_theta=False, theta_val=0.0, fix_AV=False, AV_val=0.0):
        """A native context whose model tables match the current attributes and fit flags."""
        import torch
        dev = self._device_index()
        torch.cuda.set_device(dev)
        ctx = native.Context(dev)
        RV = float(getattr(self, 'RV', 3.0))
        a_dust, _ = self.grid.dust_basis(np.array(RV))
        ctx.set_model(tau_knots=self.tau_knots, KD_t=self.KD_t, J_l=self.J_l_T, hsiao=self.hsiao_flux,
                      M_fitz=self.M_fitz_block, a_dust=a_dust, W0=self.W0, W1=self.W1, L_Sigma=self.L_Sigma,
                      tauA=self.tauA, Ds_prior_sd=math.sqrt(25.0 + self.sigma0 ** 2), rv_mode=rv_mode, RV=RV,
                      mu_R=float(getattr(self, 'mu_R', 0.0)), sigma_R=float(getattr(self, 'sigma_R', 1.0)),
                      trunc_R=self.trunc_val, fix_tmax=fix_tmax, fix_theta=fix_theta, fix_AV=fix_AV,
                      theta_val=theta_val, AV_val=AV_val)
        return ctx

    @property
    def n_eps(self):
        return (len(self.l_knots) - 2) * len(self.tau_knots)

    # ------------------------------------------------------------------ parameter packing
    def param_dim(self, rv_mode=native.RV_FIXED):
        return self.n_eps + (4 if rv_mode == native.RV_FIXED else 5)

    def run(self, args, cmd_args=None):
        """Reference :1735-1924 for the SNANA text branch, modes ``fitting`` and ``training_globalRV``
        (``bayesn_b200/driver.py``)."""
        from . import driver
        return driver.run(self, args, cmd_args)

    def process_dataset(self, args):
        """Reference :2370-2730, SNANA text directory branch (``bayesn_b200/driver.py``)."""
        from . import driver
        return driver.process_dataset(self, args)

    # thin mirrors of reference methods that live in other modules here
    def parse_yaml_input(self, args, cmd_args=None):
        """Reference :1656-1733: defaults + command-line overrides, the knots of a training run, then the data set."""
        from . import driver
        args = driver.parse_yaml_input(self, args, cmd_args)
        if 'training' in args['mode'].lower():
            driver.apply_training_knots(self, args)
        ds = driver.process_dataset(self, args)
        self.data, self.sn_list, self.peak_mjds = ds['obs'], ds['sn_names'], ds['peak_mjds']
        self.used_band_inds = ds['band_inds']
        t = self.data[0, ...]
        self.hsiao_interp = np.array([HSIAO_PHASE0_INDEX + np.floor(t), HSIAO_PHASE0_INDEX + np.ceil(t), np.remainder(t, 1)])
        self._dataset = ds
        return args

    def initial_guess(self, args, reference_model='M20_model'):
        """Reference :1568-1654: the dictionary of starting values of a training run (global numpy RNG, the
        reference's draw order), for the data set loaded by ``parse_yaml_input`` / ``run``."""
        from . import train as btrain
        return btrain.initial_values(self, self.data, reference_model, mode=args.get('mode', 'training_globalRV'))

    def postprocess(self, samples, args):
        """Reference :2169-2367 for samples in the reference layout: the training branch (sign convention, rescaling,
        ``bayesn.yaml``, chains) or, for fits, ``chains.pkl`` + ``fit_summary.csv`` (the FITRES / LCPLOT tables are
        written by ``run`` from the device-side summaries)."""
        from . import train as btrain, outputs
        if 'W1' in samples:
            return btrain.postprocess_training(self, samples, args['outputdir'], args['mode'])
        return outputs.write_chains(samples, args['outputdir'])

    @staticmethod
    def spline_coeffs_irr_step(x_now, x, invkd):
        """Reference :761-816: the row of the cubic-spline matrix at one point (vectorised here over any shape)."""
        return spline_row(x_now, np.asarray(x, dtype=np.float64), np.asarray(invkd, dtype=np.float64))

    def _load_hsiao_template(self):
        """Reference :260-285: the Hsiao template resampled onto the model's wavelength grid."""
        self.hsiao_flux = self.grid.hsiao_flux
        self.hsiao_t = self.grid.hsiao_phase

    def prepare_training(self, obs, weights, band_inds=None):
        """Upload a training set (``obs[1]``/``obs[2]`` in magnitudes, reference :2711-2730) and
        return the native context for ``train_logp_grad`` (``train_model_globalRV``, :1115-1182)."""
        if band_inds is None:
            band_inds = np.arange(weights.shape[2])
        ctx = self._context(native.RV_FIXED)
        builder = packing.DeviceTileBuilder(ctx, self.filters, band_inds, self.RV_MW) if weights is None else None
        packed = packing.pack_dataset(obs, weights, self.zps[band_inds], self.offsets[band_inds], self.M0,
                                      math.sqrt(25.0 + self.sigma0 ** 2), self.tauA, self.n_eps,
                                      device=f'cuda:{ctx.device}', train=True, sigma_pec=self.sigma_pec,
                                      builder=builder)
        ctx.bind_dataset(packed)
        return ctx

    def train(self, obs, weights, band_inds=None, num_chains=4, num_warmup=500, num_samples=500, max_tree_depth=10,
              initialisation='M20_model', seed=0, group=None, outputdir=None, mode='training_globalRV',
              progress=None):
        """Population training with ``train_model_globalRV`` (reference :1115-1182) sampled by the
        device-resident NUTS - what ``run`` does for ``mode: training_globalRV`` (:1766-1770,
        :1913-1924): ``NUTS(step_size=0.1, target_accept_prob=0.8, dense_mass=False,
        regularize_mass_matrix=False)``, every chain started from ``initial_guess`` (:1568-1654).

        ``obs`` (10, N_obs, N) / ``weights`` are the FULL training set on every rank; with ``group``
        (torch.distributed) each rank keeps a contiguous block of SNe, the per-leapfrog sums are
        all-reduced over the group and the latent draws are gathered at the end.  Returns (samples
        dict with the reference's site names, ``(chains, draws, ...)`` shapes; extras dict)."""
        import torch
        from . import train as btrain
        from .distributed import shard_bounds
        obs = np.asarray(obs)
        N = obs.shape[-1]
        rank, world = 0, 1
        if group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = shard_bounds(N, rank, world)
        ctx = self.prepare_training(obs[:, :, lo:hi], None if weights is None else weights[lo:hi], band_inds)
        dev = f'cuda:{ctx.device}'
        qg0, ql0 = btrain.initial_position(self, obs, num_chains, initialisation, seed=seed, device=dev)
        out = ctx.train_nuts(qg0, ql0[lo:hi].contiguous(), num_warmup=num_warmup, num_samples=num_samples,
                             max_tree_depth=max_tree_depth, seed=seed, group=group, sn_offset=lo, progress=progress)
        sl = out['samples_l']
        if group is not None and world > 1:
            from .distributed import gather_rows
            sl = gather_rows(sl, N, group)           # device tensors: NCCL on GPUs
        samples = btrain.constrain(self, out['samples_g'].cpu().numpy(), sl.cpu().numpy())
        stats = out['stats'].cpu().numpy()
        extras = {'potential_energy': stats[:, num_warmup:, 3], 'stats': stats, 'info': out['info'], 'passes': out['passes'], 'launches': ctx.launches()}
        if outputdir is not None and rank == 0:
            extras['yaml'], _ = btrain.postprocess_training(self, samples, outputdir, mode)
        return samples, extras

    def prepare(self, obs, weights, band_inds=None, rv_mode=None, **fix):
        """Upload model + data set; returns the native context ready for ``logp_grad`` / ``nuts``.
        ``obs`` (10, N_obs, N) and ``weights`` (N, 300, n_b) use the reference layouts; the band
        indices in ``obs[-6]`` address the columns of ``weights`` (``band_inds`` gives the global
        filter index of each column, default: all filters)."""
        if rv_mode is None:
            rv_mode = self._rv_mode()
        if band_inds is None:
            band_inds = np.arange(weights.shape[2])
        ctx = self._context(rv_mode, **fix)
        # weights=None: the per-SN band weights (reference _calculate_band_weights :499-554) are built
        # on the device straight into the packed tiles, from obs[-5] (z) and obs[-2] (E(B-V)_MW)
        builder = packing.DeviceTileBuilder(ctx, self.filters, band_inds, self.RV_MW) if weights is None else None
        packed = packing.pack_dataset(obs, weights, self.zps[band_inds], self.offsets[band_inds], self.M0,
                                      math.sqrt(25.0 + self.sigma0 ** 2), self.tauA, self.n_eps,
                                      device=f'cuda:{ctx.device}', builder=builder)
        ctx.bind_dataset(packed)
        return ctx

    def constrain(self, q, rv_mode=native.RV_FIXED):
        """Unconstrained draws (..., D) -> dict of constrained arrays (torch), reference site names."""
        import torch
        ne = self.n_eps
        out = {'eps_tform': q[..., :ne], 'theta': q[..., ne], 'AV': torch.exp(q[..., ne + 1]),
               'tmax': -10 + 20 * torch.sigmoid(q[..., ne + 2]), 'Ds': q[..., ne + 3]}
        Ls = torch.as_tensor(self.L_Sigma, dtype=q.dtype, device=q.device)
        out['eps'] = out['eps_tform'] @ Ls.T
        if rv_mode == native.RV_UNIFORM:
            out['RV'] = 1 + 5 * torch.sigmoid(q[..., ne + 4])
        elif rv_mode == native.RV_POP:
            rt = torch.sigmoid(q[..., ne + 4])
            out['RV_tform'] = rt
            nd = torch.distributions.Normal(0.0, 1.0)
            pa = float(nd.cdf(torch.tensor((self.trunc_val - self.mu_R) / self.sigma_R)))
            out['RV'] = self.mu_R + self.sigma_R * nd.icdf((pa + rt.double() * (1 - pa))).to(q.dtype)
        return out

    # ------------------------------------------------------------------ forward model
    def _forward(self, mag, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights,
                 band_inds=None):
        import torch
        dev = f'cuda:{self._device_index()}'
        key = ('fwd',)
        ctx = self._ctxs.get(key)
        if ctx is None:
            ctx = self._ctxs[key] = self._context(native.RV_FIXED)
        to = lambda a, dt=torch.float32: torch.as_tensor(np.array(a), device=dev).to(dt)
        theta = to(theta)
        N = theta.shape[0]
        RV = np.broadcast_to(np.asarray(RV, dtype=np.float64), (N,))
        if band_inds is None:
            band_inds = np.arange(np.asarray(weights).shape[2]) if np.asarray(weights).shape[2] == len(self.zps) \
                else self.used_band_inds
        bls = packing.band_log10_scale(self.zps[band_inds], self.offsets[band_inds])
        out = ctx.flux_batch(M0, theta, to(AV), to(W0), to(W1), to(eps), to(Ds), to(RV),
                             to(band_indices, torch.int32), to(np.asarray(mask) != 0, torch.uint8), to(J_t),
                             to(hsiao_interp), to(weights), bls, mag=mag)
        return out.cpu().numpy().astype(np.float64)

    def get_spectra(self, theta, AV, W0, W1, eps, RV, J_t, hsiao_interp):
        """Reference :556-643: rest-frame, host-dust-extinguished SEDs (N, 300, N_obs) on the model grid."""
        import torch
        dev = f'cuda:{self._device_index()}'
        ctx = self._ctxs.get(('fwd',))
        if ctx is None:
            ctx = self._ctxs[('fwd',)] = self._context(native.RV_FIXED)
        to = lambda a: torch.as_tensor(np.array(a), device=dev).to(torch.float32)
        theta = to(theta)
        RV = np.broadcast_to(np.asarray(RV, dtype=np.float64), (theta.shape[0],))
        out = ctx.spectra_batch(theta, to(AV), to(W0), to(W1), to(eps), to(RV), to(J_t), to(hsiao_interp))
        return out.cpu().numpy().astype(np.float64)

    def get_flux_batch(self, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights):
        """Reference :645-709: FLUXCAL (N_obs, N)."""
        return self._forward(False, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights)

    def get_mag_batch(self, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights):
        """Reference :711-759: magnitudes (N_obs, N)."""
        return self._forward(True, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights)

    # ------------------------------------------------------------------ fitting
    def build_fit_data(self, t, flux, flux_err, filters, z, ebv_mw=0, peak_mjd=None, filt_map={}, drop_bands=[],
                       mag=False):
        """Data preparation of ``fit`` (reference :2061-2106): returns ``obs`` (10, n, 1) with band
        indices into the compact used-band list, ``weights`` (1, 300, n_used) and ``band_inds``."""
        if type(drop_bands) == str:
            drop_bands = [drop_bands]
        drop_bands = list(drop_bands)
        t, flux, flux_err, filters = np.array(t, dtype=float), np.array(flux, dtype=float), \
            np.array(flux_err, dtype=float), np.array(filters)
        if mag:
            flux = np.power(10, (27.5 - flux) / 2.5)
            flux_err = (np.log(10) / 2.5) * flux * flux_err
        if peak_mjd is not None:
            t = (t - peak_mjd) / (1 + z)
        keep = (t > self.tau_knots.min()) & (t < self.tau_knots.max())
        flux, flux_err, filters, t = flux[keep], flux_err[keep], filters[keep], t[keep]
        filters = np.array([filt_map.get(f, f) for f in filters])
        for f in np.unique(filters):
            if f not in self.band_dict.keys():
                raise ValueError(f'Filter "{f}" not defined in BayeSN, either add a mapping to filt_map to ensure '
                                 f'that your filter names match up with ones built-in or add some custom filters if '
                                 f'you want to use your own')
            if z > (self.band_lim_dict[f][0] / self.l_knots[0] - 1) or z < (
                    self.band_lim_dict[f][1] / self.l_knots[-1] - 1):
                drop_bands.append(f)
        for f in drop_bands:
            inds = filters != f
            flux, flux_err, filters, t = flux[inds], flux_err[inds], filters[inds], t[inds]
        n_data = len(t)
        if n_data == 0:
            raise ValueError('No data in rest-frame phase range covered by model, maybe you gave the wrong peak MJD?')
        glob = np.array([self.band_dict[f] for f in filters])
        band_inds = np.r_[0, np.unique(glob)]                 # NULL band + the bands in play
        local = np.searchsorted(band_inds, glob)
        data = np.zeros((10, n_data, 1))
        data[0, :, 0] = t
        data[1, :, 0] = flux
        data[2, :, 0] = flux_err
        data[4, :, 0] = local
        data[5, :, 0] = z
        data[7, :, 0] = self.cosmo.distmod(np.array([z]))[0]
        data[8, :, 0] = ebv_mw
        data[9, :, 0] = 1
        weights = self._calculate_band_weights(data[-5, 0, :], data[-2, 0, :], band_inds)
        return data, weights, band_inds

    def _fit_devices(self, N):
        """Devices one process spreads a batch fit over.  The reference's ``num_devices`` (:202) asks numpyro
        for that many local devices and falls back with a warning when there are fewer (:1831,
        ``chain_method='parallel'``); here SNe - not chains - are the parallel axis.  Under ``torchrun`` (one
        process per GPU) or with an explicit ``device=`` the process owns exactly one GPU."""
        import torch
        if self._device is not None or 'LOCAL_RANK' in os.environ or not torch.cuda.is_available():
            return [self._device_index()]
        n = max(1, min(int(self.num_devices or 1), torch.cuda.device_count(), int(N)))
        return list(range(n))

    def _to_host(self, t):
        """Device tensor -> fresh numpy array through a page-locked staging buffer kept on the model (a pageable
        cudaMemcpy of the ~1 GB of draws of a 1000-SN fit runs at ~3 GB/s; pinned DMA + a host memcpy is 3x faster)."""
        import torch
        n = t.numel() * t.element_size()
        if n < (1 << 20):
            return t.cpu().numpy()
        buf = getattr(self, '_pinned', None)
        if buf is None or buf.numel() < n:
            buf = self._pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
        view = buf[:n].view(t.dtype).view(t.shape)
        view.copy_(t)
        # torch's host copy is multi-threaded; torchrun pins OMP_NUM_THREADS=1, so give the copy this rank's share of
        # the host cores for its duration
        nt = torch.get_num_threads()
        cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 8)
        share = max(1, cores // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', 1))))
        try:
            torch.set_num_threads(max(nt, min(8, share)))
            return torch.empty(t.shape, dtype=t.dtype).copy_(view).numpy()
        finally:
            torch.set_num_threads(nt)

    def fit_batch(self, obs, weights, band_inds=None, num_chains=4, num_warmup=250, num_samples=250,
                  max_tree_depth=10, seed=0, RV=None, fix_tmax=False, fix_theta=False, theta_val=0.0, fix_AV=False,
                  AV_val=0.0, init='median', return_device=False, q_init=None, sn_offset=0, group=None,
                  warps_per_chain=1, gather=True):
        """Independent NUTS fits of N supernovae (the vmapped ``fit_vmap_mcmc`` of reference
        :1809-1849).  Returns (samples dict with arrays shaped (chains, draws, [dim,] N) like
        the reference after :1839-1848, extras dict).

        Sharding (SNe are independent posteriors, so there is no data-path collective): with ``group`` (a
        ``torch.distributed`` process group, one process per GPU) every rank fits the contiguous block
        ``shard_bounds(N, rank, world)`` of ``obs`` and - with ``gather`` - the per-SN rows are all-gathered so
        every rank returns the full arrays; without a group one process spreads the SNe over
        ``min(num_devices, visible GPUs)`` devices.  ``sn_offset``: index of the first SN of ``obs`` in the
        whole job - the random streams are keyed by the global SN index, so any sharding reproduces the
        single-GPU draws bit for bit (for the same ``warps_per_chain``).

        ``warps_per_chain`` (1, 2, 4, 8 or 'auto'): latency mode - K warps share the observations of one
        (SN, chain); 'auto' picks the largest K whose warps still fit the GPU in one resident wave."""
        import torch
        from .distributed import shard_bounds, gather_rows
        obs = np.asarray(obs)
        N = obs.shape[-1]
        rv_mode = self._rv_mode(RV)
        fix = dict(fix_tmax=bool(fix_tmax), fix_theta=bool(fix_theta), theta_val=float(theta_val), fix_AV=bool(fix_AV),
                   AV_val=float(AV_val))
        kw = dict(num_warmup=num_warmup, num_samples=num_samples, max_tree_depth=max_tree_depth, seed=seed,
                  warps_per_chain=warps_per_chain)
        rank, world = 0, 1
        if group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = shard_bounds(N, rank, world)
        sl = lambda a, i0, i1: None if a is None else a[i0:i1]
        devices = self._fit_devices(hi - lo) if (group is None and not return_device and q_init is None) else [self._device_index()]
        runs = []
        for d, dv in enumerate(devices):
            l0, l1 = shard_bounds(hi - lo, d, len(devices))
            l0, l1 = lo + l0, lo + l1
            with torch.cuda.device(dv):
                if len(devices) > 1:
                    self._device = f'cuda:{dv}'
                try:
                    ctx = self.prepare(obs[:, :, l0:l1], sl(weights, l0, l1), band_inds, rv_mode, **fix)
                finally:
                    if len(devices) > 1:
                        self._device = None
                if q_init is None:
                    qi = ctx.init_to_median(num_chains, seed=seed, sn_offset=sn_offset + l0)
                else:
                    # initial positions for the whole data set are sliced like the data; a shard-sized array is used as is
                    qi = torch.as_tensor(q_init, dtype=torch.float32, device=f'cuda:{ctx.device}')
                    qi = (qi[l0:l1] if qi.shape[0] == N and N != l1 - l0 else qi).contiguous()
                out = ctx.nuts(qi, sn_offset=sn_offset + l0, **kw)           # asynchronous: the devices run concurrently
            runs.append((ctx, out))
        for ctx, out in runs:
            torch.cuda.synchronize(ctx.device)
        if return_device:
            ctx, out = runs[0]
            return self.constrain(out['samples'], rv_mode), out
        parts = []
        for ctx, out in runs:
            cons = self.constrain(out['samples'], rv_mode)                   # (n, C, S[, dim]) on the device
            cons['_stats'] = out['stats']
            cons['_info'] = out['chain_info']
            parts.append({k: v for k, v in cons.items()})
        merged = {k: (torch.cat([p[k].to(parts[0][k].device) for p in parts], dim=0) if len(parts) > 1 else parts[0][k])
                  for k in parts[0]}
        if group is not None and world > 1 and gather:
            merged = {k: gather_rows(v.contiguous(), N, group) for k, v in merged.items()}
        stats_all = merged.pop('_stats').cpu().numpy()
        info = merged.pop('_info').cpu().numpy()
        samples = {}
        d2h = stats_all.nbytes + info.nbytes
        for k, v in merged.items():
            # reference :1846-1849: 4-D sites (N, C, S, dim) -> (C, S, dim, N), 3-D sites -> (C, S, N); the
            # transposition and the float64 widening run on the device, the host receives the final layout
            v = v.permute(1, 2, 3, 0) if v.ndim == 4 else v.permute(1, 2, 0)
            samples[k] = self._to_host(v.contiguous().double())
            d2h += samples[k].nbytes
        stats = stats_all[:, :, num_warmup:]
        samples['diverging'] = stats[..., 2].transpose(1, 2, 0) > 0
        extras = {'accept_prob': stats[..., 0].transpose(1, 2, 0), 'num_steps': stats[..., 1].transpose(1, 2, 0),
                  'potential_energy': stats[..., 3].transpose(1, 2, 0),
                  'chain_info': info, 'launches': sum(c.launches() for c, _ in runs),
                  'all_stats': stats_all, 'devices': [c.device for c, _ in runs], 'shard': (lo, hi),
                  'warps_per_chain': runs[0][1].get('warps_per_chain', 1), 'd2h_bytes': int(d2h),
                  'h2d_bytes': int(sum(c._keep['data']['h2d_bytes'][0] for c, _ in runs))}
        return samples, extras

    # ------------------------------------------------------------------ log densities (numpyro models of the reference)
    def fit_model_globalRV(self, obs, weights, fix_tmax=False, fix_theta=False, theta_val=0, fix_AV=False, AV_val=0,
                           band_inds=None):
        """Reference :883-945.  The reference's method is a numpyro model (a callable the sampler traces); its
        counterpart here is the bound log density: an object with ``logp_grad(q)`` ((N, C, D) unconstrained
        positions -> log joint density incl. Jacobians (N, C) and its gradient), ``init_to_median(n_chains)``,
        ``constrain(q)`` and ``nuts(q_init, ...)`` - what ``fit`` / ``fit_batch`` hand to the device sampler."""
        return FitLogDensity(self, obs, weights, band_inds, native.RV_FIXED, fix_tmax, fix_theta, theta_val, fix_AV, AV_val)

    def fit_model_popRV(self, obs, weights, fix_tmax=False, fix_theta=False, theta_val=0, fix_AV=False, AV_val=0,
                        band_inds=None):
        """Reference :996-1060 (R_V ~ truncated N(mu_R, sigma_R) per SN)."""
        return FitLogDensity(self, obs, weights, band_inds, native.RV_POP, fix_tmax, fix_theta, theta_val, fix_AV, AV_val)

    def fit_model_uniformRV(self, obs, weights, fix_tmax=False, fix_theta=False, theta_val=0, fix_AV=False, AV_val=0,
                            band_inds=None):
        """Reference :818-881 (R_V ~ U(1, 6) per SN)."""
        return FitLogDensity(self, obs, weights, band_inds, native.RV_UNIFORM, fix_tmax, fix_theta, theta_val, fix_AV, AV_val)

    def train_model_globalRV(self, obs, weights, band_inds=None):
        """Reference :1115-1182: the joint training density; ``logp_grad(qg, ql[, group])`` evaluates it (and
        its gradient) for unconstrained globals (C, G) and per-SN latents (N, C, Dl)."""
        return TrainLogDensity(self, obs, weights, band_inds)

    def fit(self, t, flux, flux_err, filters, z, ebv_mw=0, peak_mjd=None, filt_map={}, print_summary=True,
            file_prefix=None, drop_bands=[], fix_tmax=False, fix_theta=False, fix_AV=False, RV=False, mu_R=False,
            sigma_R=False, mag=False, seed=0, warps_per_chain='auto'):
        """Reference :1994-2168 (4 chains x (250 + 250), max_tree_depth 10).  A single object is four chains on
        a whole GPU: the sampler runs in latency mode (``warps_per_chain='auto'``)."""
        data, band_weights, band_inds = self.build_fit_data(t, flux, flux_err, filters, z, ebv_mw, peak_mjd,
                                                            filt_map, drop_bands, mag)
        rv_arg = None
        if RV == 'uniform':
            rv_arg = 'uniform'
        else:
            if RV:
                self.RV = float(RV)
                self.model_type = 'fixed_RV'
            elif mu_R:
                if not sigma_R:
                    raise ValueError('You have set a custom mu_R, please also set a custom sigma_R')
                self.mu_R, self.sigma_R = float(mu_R), float(sigma_R)
                self.model_type = 'pop_RV'
        theta_val = 0
        if fix_theta:
            theta_val, fix_theta = fix_theta, True
        AV_val = 0
        if fix_AV:
            AV_val, fix_AV = fix_AV, True
        samples, extras = self.fit_batch(data, band_weights, band_inds, num_chains=4, num_warmup=250,
                                         num_samples=250, max_tree_depth=10, seed=seed, RV=rv_arg,
                                         fix_tmax=fix_tmax, fix_theta=fix_theta, theta_val=theta_val, fix_AV=fix_AV,
                                         AV_val=AV_val, warps_per_chain=warps_per_chain)
        samples.pop('diverging')
        # the plate of a single-object fit keeps eps_tform as (chains, draws, 1, n_eps) (example_fits.ipynb:166-197);
        # the vmapped batch layout is (chains, draws, n_eps, N)
        samples['eps_tform'] = samples['eps_tform'].transpose(0, 1, 3, 2)
        # numpyro keeps extra fields out of get_samples(): like mcmc.get_extra_fields() they stay on the side
        self.extra_fields = {'potential_energy': extras['potential_energy'][..., 0]}
        if print_summary:
            # mcmc.print_summary() leaves deterministic sites out: eps, and R_V where it is a transform of RV_tform
            det = {'eps'} | ({'RV'} if 'RV_tform' in samples else set())
            summary.print_summary({k: v for k, v in samples.items() if k not in det},
                                  n_div=int((extras['chain_info'][..., 3]).sum()))
        if peak_mjd is not None:
            samples['peak_MJD'] = peak_mjd + samples['tmax'] * (1 + z)
        muhat = self.cosmo.distmod(np.array([z]))[0]
        muhat_err = 5
        Ds_err = np.sqrt(muhat_err * muhat_err + self.sigma0 * self.sigma0)
        samples['mu'] = np.random.normal(
            (samples['Ds'] * np.power(muhat_err, 2) + muhat * np.power(self.sigma0, 2)) / np.power(Ds_err, 2),
            np.sqrt((np.power(self.sigma0, 2) * np.power(muhat_err, 2)) / np.power(Ds_err, 2)))
        samples['delM'] = samples['Ds'] - samples['mu']
        if fix_tmax:
            samples['tmax'] = np.zeros_like(samples['tmax'])
        if fix_theta:
            samples['theta'] = np.full_like(samples['theta'], theta_val)
        if fix_AV:
            samples['AV'] = np.full_like(samples['AV'], AV_val)
        if file_prefix is not None:
            summary.summary_table(samples).to_csv(f'{file_prefix}_fit_summary.csv')
            with open(f'{file_prefix}_chains.pkl', 'wb') as file:
                pickle.dump(samples, file)
        return samples, (z, ebv_mw)

    def fit_from_file(self, path, filt_map={}, peak_mjd_key='SEARCH_PEAKMJD', print_summary=True, file_prefix=None,
                      drop_bands=[], fix_tmax=False, fix_theta=False, fix_AV=False, RV=False, mu_R=False,
                      sigma_R=False, mag=False, seed=0, warps_per_chain='auto'):
        """Reference :1926-1992 (SNANA text file)."""
        meta, lc = snana.read_snana_ascii(path)
        return self.fit(lc['MJD'], lc['FLUXCAL'], lc['FLUXCALERR'], lc['FLT'], meta['REDSHIFT_HELIO'],
                        ebv_mw=meta['MWEBV'], peak_mjd=meta[peak_mjd_key], filt_map=filt_map,
                        print_summary=print_summary, file_prefix=file_prefix, drop_bands=drop_bands,
                        fix_tmax=fix_tmax, fix_theta=fix_theta, fix_AV=fix_AV, RV=RV, mu_R=mu_R, sigma_R=sigma_R,
                        mag=mag, seed=seed, warps_per_chain=warps_per_chain)

    # ------------------------------------------------------------------ simulation (reference :3103-3424)
    def sample_del_M(self, N):
        return np.random.normal(0, self.sigma0, N)

    def sample_AV(self, N):
        return np.random.exponential(self.tauA, N)

    def sample_theta(self, N):
        return np.random.normal(0, 1, N)

    def sample_epsilon(self, N):
        L, T = len(self.l_knots), len(self.tau_knots)
        n_sig = (L - 2) * T
        eps_tform = np.random.multivariate_normal(np.zeros(n_sig), np.eye(n_sig), N).T
        eps = np.matmul(self.L_Sigma, eps_tform).T
        eps = np.reshape(eps, (N, L - 2, T), order='F')
        eps_full = np.zeros((N, L, T))
        eps_full[:, 1:-1, :] = eps
        return eps_full

    def simulate_light_curve(self, t, N, bands, yerr=0, err_type='mag', z=0, zerr=1e-4, mu=0, ebv_mw=0, RV=None,
                             logM=None, tmax=0, del_M=None, AV=None, theta=None, eps=None, mag=True,
                             write_to_files=False, output_dir=None):
        """Reference :3103-3345: same validation, same draw order from the global numpy RNG."""
        def per_object(val, name):
            val = np.array(val)
            if len(val.shape) == 0:
                return val.repeat(N)
            if val.shape[0] != N:
                raise ValueError(f'If not providing a scalar {name} value, array must be of same length as the '
                                 f'number of objects to simulate, N')
            return val
        del_M = self.sample_del_M(N) if del_M is None else per_object(del_M, 'del_M')
        AV = self.sample_AV(N) if AV is None else per_object(AV, 'AV')
        theta = self.sample_theta(N) if theta is None else per_object(theta, 'theta')
        L, T = self.l_knots.shape[0], self.tau_knots.shape[0]
        if eps is None:
            eps = self.sample_epsilon(N)
        elif len(np.array(eps).shape) == 0:
            if np.array(eps) == 0:
                eps = np.zeros((N, L, T))
            else:
                raise ValueError('For epsilon, please pass an array-like object of shape (N, l_knots, tau_knots). '
                                 'The only scalar value accepted is 0, which will effectively remove the effect of '
                                 'epsilon')
        else:
            eps = np.array(eps)
            if len(eps.shape) == 2 and eps.shape == (L, T):
                eps = np.repeat(eps[None], N, axis=0)
            if len(eps.shape) != 3 or eps.shape[0] != N or eps.shape[1] != L or eps.shape[2] != T:
                raise ValueError('For epsilon, please pass an array-like object of shape (N, l_knots, tau_knots)')
        ebv_mw = per_object(ebv_mw, 'ebv_mw').astype(float)
        tmax = per_object(tmax, 'tmax').astype(float)
        RV = per_object(self.RV if RV is None else RV, 'RV').astype(float)
        z = per_object(z, 'z').astype(float)
        if type(mu) == str and mu == 'z':
            mu = self.cosmo.distmod(z)
        else:
            mu = per_object(mu, 'mu').astype(float)
        param_dict = {'del_M': del_M, 'AV': AV, 'theta': theta, 'eps': eps, 'z': z, 'mu': mu, 'ebv_mw': ebv_mw,
                      'RV': RV}
        t = np.asarray(t, dtype=float)
        for band in bands:
            if band not in self.band_dict.keys():
                raise ValueError(f'{band} is not present in filters yaml file')
        if t.shape[0] == np.array(bands).shape[0]:
            glob = np.array([self.band_dict[band] for band in bands])
        else:
            num_per_band = t.shape[0]
            glob = np.repeat(np.array([self.band_dict[band] for band in bands]), num_per_band)
            t = np.repeat(t[:, None], len(bands), axis=1).flatten(order='F')
        band_inds = np.r_[0, np.unique(glob)]
        band_indices = np.searchsorted(band_inds, glob)[:, None].repeat(N, axis=1).astype(int)
        mask = np.ones_like(band_indices)
        if N > 1 and np.all(z == z[0]) and np.all(ebv_mw == ebv_mw[0]):
            # one object drawn N times (get_flux_from_chains): its band weights are computed once
            band_weights = np.broadcast_to(self._calculate_band_weights(z[:1], ebv_mw[:1], band_inds),
                                           (N, 300, len(band_inds)))
        else:
            band_weights = self._calculate_band_weights(z, ebv_mw, band_inds)
        tt = np.repeat(t[..., None], N, axis=1) - tmax[None, :]
        hsiao_interp = np.array([HSIAO_PHASE0_INDEX + np.floor(tt), HSIAO_PHASE0_INDEX + np.ceil(tt),
                                 np.remainder(tt, 1)])
        J_t = self.J_t_map(tt).transpose(1, 2, 0)                      # (N, T, N_obs)
        data = self._forward(bool(mag), self.M0, theta, AV, self.W0, self.W1, eps, mu + del_M, RV, band_indices,
                             mask, J_t, hsiao_interp, band_weights, band_inds=band_inds)
        yerr = np.array(yerr, dtype=float)
        if err_type == 'mag' and not mag:
            yerr = yerr * (np.log(10) / 2.5) * data
        if len(yerr.shape) == 0:
            yerr = np.ones_like(data) * yerr
        elif len(yerr.shape) == 1:
            assert data.shape[0] == yerr.shape[0], f'If passing a 1d array, shape of yerr must match number of ' \
                                                   f'simulated data points per objects, {data.shape[0]}'
            yerr = np.repeat(yerr[..., None], N, axis=1)
        else:
            assert data.shape == yerr.shape, f'If passing a 2d array, shape of yerr must match generated data ' \
                                             f'shape of {data.shape}'
        data = np.random.normal(data, yerr)
        if write_to_files and mag:
            if output_dir is None:
                raise ValueError('If writing to SNANA files, please provide an output directory')
            os.makedirs(output_dir, exist_ok=True)
            for i in range(N):
                flt = [self.inv_band_dict[band_inds[b]] for b in band_indices[:, i]]
                snana.write_snana_lcfile(output_dir, f'{i}', tt[:, i] * (1 + z[i]), flt, data[:, i], yerr[:, i], 0,
                                         z[i], z[i], zerr, ebv_mw[i])
        elif write_to_files:
            raise ValueError('If writing to SNANA files, please generate mags')
        self._last_sim = dict(t=np.repeat(t[:, None], N, axis=1), band_indices=band_indices, band_inds=band_inds,
                              band_weights=band_weights)
        return data, yerr, param_dict

    def simulate_spectrum(self, t, N, dl=10, z=0, mu=0, ebv_mw=0, RV=None, logM=None, del_M=None, AV=None, theta=None,
                          eps=None):
        """Reference :2923-3101: rest-frame, host-dust-extinguished spectra of N simulated objects on a linear
        wavelength grid of spacing ``dl`` between the first and last wavelength knot, at phases ``t``.  Returns
        (``l_o`` (N, n_wave) observer-frame wavelengths, ``spectra`` (N, n_wave, len(t)), ``param_dict``); priors are
        sampled in the reference's order (del_M, AV, theta, eps) from the global numpy RNG.

        The reference rebuilds ``model_wave`` / ``J_l_T`` / the Hsiao table of the INSTANCE for that grid (a side effect,
        SURVEY Appendix E) and computes host / Milky Way extinction arrays it never applies; here the instance is left
        untouched: the grid is evaluated in chunks of 300 bins by the ``get_spectra`` kernel on temporary tables."""
        import torch
        from .static import spline_matrix, inv_kd, fitz_matrix

        def per_object(val, name, msg=None):
            val = np.array(val)
            if len(val.shape) == 0:
                return val.repeat(N)
            if val.shape[0] != N:
                raise ValueError(msg or f'If not providing a scalar {name} value, array must be of same length as the '
                                        f'number of objects to simulate, N')
            return val
        each = lambda name: (f'For {name}, either pass a single scalar value or an array of values for each of the N '
                             f'simulated objects')
        del_M = self.sample_del_M(N) if del_M is None else per_object(del_M, 'del_M')
        AV = self.sample_AV(N) if AV is None else per_object(AV, 'AV')
        theta = self.sample_theta(N) if theta is None else per_object(theta, 'theta')
        L, T = self.l_knots.shape[0], self.tau_knots.shape[0]
        if eps is None:
            eps = self.sample_epsilon(N)
        else:
            eps = np.array(eps)
            if len(eps.shape) == 0:
                if eps == 0:
                    eps = np.zeros((N, L, T))
                else:
                    raise ValueError('For epsilon, please pass an array-like object of shape (N, l_knots, tau_knots). The '
                                     'only scalar value accepted is 0, which will effectively remove the effect of epsilon')
            elif eps.shape == (L, T):
                eps = eps[None, ...].repeat(N, axis=0)
            elif eps.shape != (N, L, T):
                raise ValueError('For epsilon, please pass an array-like object of shape (N, l_knots, tau_knots)')
        ebv_mw = per_object(ebv_mw, 'ebv_mw', each('ebv_mw')).astype(float)
        RV = per_object(self.RV if RV is None else RV, 'RV', each('RV')).astype(float)
        z = per_object(z, 'z', each('z')).astype(float)
        mu = per_object(mu, 'mu', each('mu')).astype(float)
        param_dict = {'del_M': del_M, 'AV': AV, 'theta': theta, 'eps': eps, 'z': z, 'mu': mu, 'ebv_mw': ebv_mw, 'RV': RV}
        l_r = np.arange(min(self.l_knots), max(self.l_knots) + dl, dl, dtype=float)
        l_r = l_r[l_r <= max(self.l_knots)]
        l_o = l_r[None, ...].repeat(N, axis=0) * (1 + z[:, None])
        t = np.asarray(t, dtype=float)
        tt = np.repeat(t[..., None], N, axis=1)
        hsiao_interp = np.array([HSIAO_PHASE0_INDEX + np.floor(tt), HSIAO_PHASE0_INDEX + np.ceil(tt), np.remainder(tt, 1)])
        J_t = self.J_t_map(tt).transpose(1, 2, 0)                       # (N, T, n_t)
        dev = f'cuda:{self._device_index()}'
        to = lambda a: torch.as_tensor(np.array(a, dtype=np.float32), device=dev)
        args = (to(theta), to(AV), to(self.W0), to(self.W1), to(eps), to(RV), to(J_t), to(hsiao_interp))
        KD_l = inv_kd(self.l_knots)
        ph, hw, hf = datafiles.load_hsiao()
        KD_h = inv_kd(hw)
        nb = native.NBINS
        pad = lambda a: np.concatenate([a, np.zeros((nb - a.shape[0],) + a.shape[1:])], axis=0)
        ctx = self._context(native.RV_FIXED)
        spectra = np.zeros((N, l_r.shape[0], t.shape[0]))
        for c0 in range(0, l_r.shape[0], nb):
            lam = l_r[c0:c0 + nb]
            ctx.set_model(tau_knots=self.tau_knots, KD_t=self.KD_t, J_l=pad(spline_matrix(lam, self.l_knots, KD_l)),
                          hsiao=pad(spline_matrix(lam, hw, KD_h) @ hf.T), M_fitz=pad(fitz_matrix(lam)),
                          a_dust=np.ones(nb), W0=self.W0, W1=self.W1, L_Sigma=self.L_Sigma, tauA=self.tauA,
                          Ds_prior_sd=math.sqrt(25.0 + self.sigma0 ** 2))
            out = ctx.spectra_batch(*args)                               # (N, 300, n_t)
            spectra[:, c0:c0 + lam.shape[0]] = out[:, :lam.shape[0]].cpu().numpy()
        return l_o, spectra, param_dict

    def simulated_obs_array(self, data, yerr, param_dict):
        """Pack the output of the last ``simulate_light_curve(mag=False)`` call into the
        reference's (10, N_obs, N) ``obs`` layout."""
        sim = self._last_sim
        N_obs, N = data.shape
        obs = np.zeros((10, N_obs, N))
        obs[0] = sim['t']
        obs[1] = data
        obs[2] = yerr
        obs[4] = sim['band_indices']
        obs[5] = param_dict['z'][None, :]
        obs[6] = 1e-4
        obs[7] = self.cosmo.distmod(param_dict['z'])[None, :]
        obs[8] = param_dict['ebv_mw'][None, :]
        obs[9] = 1
        return obs, sim['band_weights'], sim['band_inds']

    def chain_flux_device(self, t, band_lists, theta, AV, tmax, Ds, eps, zs, ebv_mws, RV=None, mag=False):
        """Device core of ``get_flux_from_chains``: ONE kernel launch over (SN x draw x phase x band).
        ``theta, AV, tmax, Ds`` (N, S) and ``eps`` (N, S, n_eps) are torch tensors on the model's device
        (``Ds = mu + delM``, float64 welcome), ``band_lists`` one list of filter names per SN (or one list for
        all).  Returns a (N, S, max_bands, len(t)) float32 device tensor."""
        import torch
        ctx = self._ctxs.get(('fwd',))
        if ctx is None:
            ctx = self._ctxs[('fwd',)] = self._context(native.RV_FIXED)
        dev = f'cuda:{ctx.device}'
        N = theta.shape[0]
        if not isinstance(band_lists[0], (list, tuple, np.ndarray)):
            band_lists = [list(band_lists)] * N
        for bl in band_lists:
            for b in bl:
                if b not in self.band_dict:
                    raise ValueError(f'{b} is not present in filters yaml file')
        glob = [np.array([self.band_dict[b] for b in bl], dtype=np.int64) for bl in band_lists]
        band_inds = np.unique(np.concatenate(glob))
        max_bands = max(len(g) for g in glob)
        local = np.full((N, max_bands), -1, dtype=np.int32)
        used = np.zeros((N, len(band_inds)), dtype=bool)
        for i, g in enumerate(glob):
            loc = np.searchsorted(band_inds, g)
            local[i, :len(g)] = loc
            used[i, loc] = True
        # distance: fold the SN's mean distance into its tile (float64), the kernel multiplies by 10^(-0.4 dD)
        Ds = torch.as_tensor(Ds, device=dev).double()
        fold = Ds.mean(dim=1)
        dD = (Ds - fold[:, None]).float()
        builder = packing.DeviceTileBuilder(ctx, self.filters, band_inds, self.RV_MW)
        bscale = packing.band_log10_scale(self.zps[band_inds], self.offsets[band_inds])
        zs = np.broadcast_to(np.atleast_1d(np.asarray(zs, dtype=np.float64)), (N,))
        ebv_mws = np.broadcast_to(np.atleast_1d(np.asarray(ebv_mws, dtype=np.float64)), (N,))
        sup, wt, store, wt_stride = packing.device_tiles(builder, zs, ebv_mws, used, fold.cpu().numpy() + self.M0,
                                                         bscale, dev)
        to = lambda a: torch.as_tensor(a, device=dev)
        out = ctx.chain_flux(to(np.asarray(t, dtype=np.float32)), to(theta), to(AV), to(tmax), dD, to(eps),
                             to(local), wt, sup, wt_stride, RV=None if RV is None else to(RV), mag=mag)
        torch.cuda.synchronize()
        del store
        return out

    def get_flux_from_chains(self, t, bands, chains, zs, ebv_mws, mag=True, num_samples=None, mean=False):
        """Reference :3426-3524: model photometry for posterior samples,
        (N_sne, num_samples, n_bands, len(t)).  The reference loops over SNe through
        ``simulate_light_curve`` (4.91 s per SN in its notebook); here all (SN, draw, phase, band) cells are
        one kernel launch (``bayesn_chain_flux``).  Draw order and slicing follow the reference:
        ``chains[k][..., i].flatten(order='F')[:num_samples]``; ``mean=True`` sets ``num_samples = 1`` BEFORE the
        slice (:3474-3475, :3507-3511), so - like the reference - it evaluates the first draw."""
        if type(chains) == str:
            with open(chains, 'rb') as file:
                chains = pickle.load(file)
        th = np.asarray(chains['theta'])
        C_, S_, N_sne = th.shape
        tot = C_ * S_
        if num_samples is None:
            num_samples = tot
        if mean:
            num_samples = 1
        fl = lambda k: np.asarray(chains[k], dtype=np.float64).transpose(1, 0, 2).reshape(tot, N_sne)[:num_samples].T
        eps = np.asarray(chains['eps'], dtype=np.float32)
        eps = eps.transpose(1, 0, 2, 3).reshape(tot, eps.shape[2], N_sne)[:num_samples].transpose(2, 0, 1)
        RV = fl('RV') if 'RV' in chains.keys() else None
        t = np.asarray(t, dtype=float)
        out = self.chain_flux_device(t, bands, fl('theta'), fl('AV'), fl('tmax'), fl('mu') + fl('delM'),
                                     np.ascontiguousarray(eps), zs, ebv_mws, RV=RV, mag=mag)
        return out.cpu().numpy().astype(np.float64)


class FitLogDensity:
    """Bound per-object log density (see ``SEDmodel.fit_model_globalRV``)."""

    def __init__(self, model, obs, weights, band_inds, rv_mode, fix_tmax, fix_theta, theta_val, fix_AV, AV_val):
        self.model, self.rv_mode = model, rv_mode
        self.ctx = model.prepare(obs, weights, band_inds, rv_mode, fix_tmax=bool(fix_tmax), fix_theta=bool(fix_theta),
                                 theta_val=float(theta_val), fix_AV=bool(fix_AV), AV_val=float(AV_val))
        self.N, self.D = self.ctx.N, self.ctx.D

    def init_to_median(self, n_chains=4, seed=0, sn_offset=0):
        return self.ctx.init_to_median(n_chains, seed=seed, sn_offset=sn_offset)

    def logp_grad(self, q):
        import torch
        q = torch.as_tensor(q, dtype=torch.float32, device=f'cuda:{self.ctx.device}').contiguous()
        return self.ctx.logp_grad(q)

    def __call__(self, q):
        return self.logp_grad(q)[0]

    def constrain(self, q):
        return self.model.constrain(q, self.rv_mode)

    def nuts(self, q_init, **kw):
        return self.ctx.nuts(q_init, **kw)


class TrainLogDensity:
    """Bound training log density (see ``SEDmodel.train_model_globalRV``)."""

    def __init__(self, model, obs, weights, band_inds):
        self.model = model
        self.ctx = model.prepare_training(obs, weights, band_inds)
        self.G, self.Dl, self.R = self.ctx.train_dims()

    def logp_grad(self, qg, ql, group=None):
        import torch
        dev = f'cuda:{self.ctx.device}'
        qg = torch.as_tensor(qg, dtype=torch.float32, device=dev).contiguous()
        ql = torch.as_tensor(ql, dtype=torch.float32, device=dev).contiguous()
        return self.ctx.train_logp_grad(qg, ql, group=group)

    def __call__(self, qg, ql, group=None):
        return self.logp_grad(qg, ql, group)[0]
