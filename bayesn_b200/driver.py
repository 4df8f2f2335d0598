"""``run`` / ``parse_yaml_input`` / ``process_dataset`` of the reference's ``SEDmodel``
(``bayesn/bayesn_model.py:1735-1924``, ``:1656-1733``, ``:2370-2730``) for the SNANA **text**
branch: a directory ``<version_photometry>/`` holding ``<name>.LIST`` and one SNANA text file per
object.  Modes ``fitting`` (``fit_method: mcmc``) and ``training_globalRV``; same YAML keys and
defaults, same output files (``chains.pkl``, ``input.yaml``, ``<prefix>.FITRES.TEXT``; training:
``bayesn.yaml``, ``initial_chains.pkl``; fitting also ``<prefix>.LCPLOT``).  Out of scope here: SNANA
FITS input, ``data_table`` input, SNANA cluster conventions (``jobsplit`` directory search), the VI
and dust modes.
"""
import os
import time

import numpy as np

from . import snana

C_KMS = 299792.458


def parse_yaml_input(model, args, cmd_args=None):
    """Defaults of reference :1684-1709 (the keys this driver understands)."""
    args = dict(args)
    if cmd_args is not None:
        for k, v in (vars(cmd_args) if not isinstance(cmd_args, dict) else cmd_args).items():
            if k in ('input', 'filters') or v is None:
                continue
            if k == 'map' and isinstance(v, str):
                fm = np.loadtxt(v, dtype=str)
                v = {row[0]: row[1] for row in np.atleast_2d(fm)}
            args[k] = v
    args.pop('CONFIG', None)
    args.pop('config', None)
    for k, d in (('num_chains', 4), ('num_warmup', 500), ('num_samples', 500), ('fit_method', 'mcmc'),
                 ('chain_method', 'parallel'), ('initialisation', 'median'), ('map', {}), ('drop_bands', []),
                 ('outputdir', os.getcwd()), ('outfile_prefix', 'output'), ('sim_prescale', 1),
                 ('save_fit_errors', False), ('error_floor', 0.0)):
        args[k] = args.get(k, d)
    pdp = args.get('private_data_path', [])
    args['private_data_path'] = [pdp] if isinstance(pdp, str) else list(pdp)
    args['l_knots'] = args.get('l_knots', model.l_knots.tolist())
    args['tau_knots'] = args.get('tau_knots', model.tau_knots.tolist())
    args['fit_method'] = str(args['fit_method']).lower()
    if args['fit_method'] not in ('vi', 'mcmc'):
        raise ValueError(f'Requested fitting method {args["fit_method"]}, must be one of "mcmc" or "vi"')
    if args['fit_method'] == 'vi':
        raise NotImplementedError('the VI path is outside the scope of this implementation; use fit_method: mcmc')
    if 'mode' not in args:
        raise ValueError("input needs a 'mode'")
    # SNANA batch mode (reference :1703-1709): jobsplit = [job id, number of jobs]
    args['jobsplit'] = args.get('jobsplit')
    if args['jobsplit'] is not None:
        args['snana'] = True
    else:
        args['jobsplit'] = [1, 1]
        args['snana'] = False
    args['jobid'] = args['jobsplit'][0]
    args['njobtot'] = args['jobsplit'][1] * args['sim_prescale']
    args['save_chains'] = args.get('save_chains', True)
    args['warps_per_chain'] = args.get('warps_per_chain', 'auto')
    if not (args['mode'].lower() == 'fitting' and args['snana']):
        os.makedirs(args['outputdir'], exist_ok=True)
    return args


def apply_training_knots(model, args):
    """Training runs may define the model they infer on their own knots (``l_knots`` / ``tau_knots`` of the input,
    reference :1721-1727): the wavelength / time grids, filter curves and spline tables are rebuilt for them.  W0, W1
    and L_Sigma of the loaded model then no longer have the shape of the model being trained; they are not used by
    the training density (the sampler carries its own), so they are replaced by placeholders of the new shape."""
    lk, tk = np.array(args['l_knots'], dtype=float), np.array(args['tau_knots'], dtype=float)
    if lk.shape == model.l_knots.shape and tk.shape == model.tau_knots.shape \
            and np.array_equal(lk, model.l_knots) and np.array_equal(tk, model.tau_knots):
        return False
    model.l_knots, model.tau_knots = lk, tk
    model._setup_band_weights()
    L, T = len(lk), len(tk)
    if np.shape(model.W0) != (L, T):
        ne = (L - 2) * T
        model.W0, model.W1, model.L_Sigma = np.zeros((L, T)), np.zeros((L, T)), 0.1 * np.eye(ne)
    return True


def resolve_photometry_dir(args):
    """Where ``version_photometry`` lives (reference :2401-2426).  Outside SNANA batch mode it is a path.  With
    ``jobsplit`` (``args['snana']``) it is the NAME of a photometry version, looked up under ``$SNDATA_ROOT/lcmerge``,
    ``$SNDATA_ROOT/SIM``, the directories listed in ``$SNDATA_ROOT/SIM/PATH_SNDATA_SIM.LIST`` and ``private_data_path``
    (entries may start with ``$ENVVAR``); it has to exist in exactly one of them.  (The reference glues the path
    components after the first together without separators, so only one-level sub-directories of an environment
    variable work there; here the rest of the path is kept as it is, and absolute paths are taken literally.)"""
    data_dir = args['version_photometry']
    if not args.get('snana'):
        return data_dir
    root = os.environ.get('SNDATA_ROOT')
    dir_list = ['SNDATA_ROOT/lcmerge', 'SNDATA_ROOT/SIM']
    sim_list = os.path.join(root, 'SIM', 'PATH_SNDATA_SIM.LIST') if root else None
    if sim_list and os.path.exists(sim_list):
        dir_list += [str(d)[1:] if str(d).startswith('$') else str(d) for d in np.atleast_1d(np.loadtxt(sim_list, dtype=str))]
    dir_list += [p[1:] if p.startswith('$') else p for p in args.get('private_data_path', [])]
    found = []
    for d in dir_list:
        if os.path.isabs(d):
            cand = os.path.join(d, data_dir)
        else:
            head, _, rest = d.partition('/')
            cand = os.path.join(os.environ.get(head, 'NULL'), rest, data_dir)
        if os.path.exists(cand) and os.path.realpath(cand) not in [os.path.realpath(f) for f in found]:
            found.append(cand)
    if len(found) == 0:
        raise ValueError(f'Requested photometry {data_dir} was not found in any of the usual public '
                         f'locations, maybe you need to specify an additional private data location')
    if len(found) > 1:
        raise ValueError(f'Requested photometry {data_dir} was found in multiple locations, please remove '
                         f'duplicates and ensure the one you want to use remains')
    return found[0]


def _snrmax(flux, ferr, bands):
    """SNRMAX1/2/3 of the FITRES table: the highest S/N, the highest in any other band, and in a third band
    (reference :2676-2690, :2849-2863); -99 when there is no such band."""
    out = []
    snr = flux / ferr
    keep = np.ones(len(snr), dtype=bool)
    for _ in range(3):
        if not keep.any():
            out.append(-99.0)
            continue
        i = int(np.argmax(np.where(keep, snr, -np.inf)))
        out.append(float(snr[i]))
        keep &= bands != bands[i]
    return out


def _one_light_curve(model, args, meta, lc, peak, zhel, fmap, used):
    """Per-file steps shared by the two input branches (reference :2600-2645, :2776-2830): rest-frame phases, filter
    mapping, bands outside the model's wavelength coverage or in ``drop_bands`` removed, error floor, non-finite
    photometry dropped, phase cut to the tau-knot range; new bands are appended to ``used`` in order of first use."""
    band_col = lc['BAND'] if 'BAND' in lc else lc['FLT']
    flt = np.array([fmap.get(str(b), str(b)) for b in band_col])
    if 'FLUXCAL' in lc:
        flux, ferr = lc['FLUXCAL'].astype(float), lc['FLUXCALERR'].astype(float)
    else:
        flux = 10 ** ((27.5 - lc['MAG']) / 2.5)
        ferr = (np.log(10) / 2.5) * flux * lc['MAGERR']
    ferr = np.maximum(ferr, args.get('error_floor', 0.0) * (np.log(10) / 2.5) * flux)
    t = (lc['MJD'] - peak) / (1 + zhel)
    keep = np.isfinite(flux) & np.isfinite(ferr) & (t > model.tau_knots.min()) & (t < model.tau_knots.max())
    for f in np.unique(flt):
        if f in args.get('drop_bands', []):
            keep &= flt != f
            continue
        if f not in model.band_dict:
            raise KeyError(f'Filter {f} not present in BayeSN, check your filter mapping')
        lo, hi = model.band_lim_dict[f]
        if zhel > lo / model.l_knots[0] - 1 or zhel < hi / model.l_knots[-1] - 1:
            keep &= flt != f
    gl = np.array([model.band_dict.get(f, 0) for f in flt])
    glk = gl[keep]
    _, first = np.unique(glk, return_index=True)
    for g in glk[np.sort(first)]:                          # order of first appearance, like pandas .unique()
        if g not in used:
            used.append(int(g))
    return dict(t=t[keep], flux=flux[keep], ferr=ferr[keep], gl=gl[keep], mjd=lc['MJD'][keep], flt=flt[keep])


FITRES_COLUMNS = ('IDSURVEY', 'TYPE', 'FIELD', 'zHEL', 'zHELERR', 'zHD', 'zHDERR', 'VPEC', 'VPECERR', 'MWEBV', 'HOST_LOGMASS',
                  'HOST_LOGMASS_ERR', 'SNRMAX1', 'SNRMAX2', 'SNRMAX3')
SIM_COLUMNS = (('SIM_GENTYPE', 'SIM_GENTYPE'), ('SIM_TEMPLATE_INDEX', 'SIM_TEMPLATE_INDEX'), ('SIM_LIBID', 'SIM_LIBID'),
               ('SIM_ZCMB', 'SIM_REDSHIFT_CMB'), ('SIM_VPEC', 'SIM_VPEC'), ('SIM_DLMAG', 'SIM_DLMU'), ('SIM_PEAKMJD', 'SIM_PEAKMJD'),
               ('SIM_THETA', 'SIM_THETA'), ('SIM_AV', 'SIM_AV'), ('SIM_RV', 'SIM_RV'))


def process_dataset(model, args):
    """SNANA text input -> reference-layout arrays.  Two branches of the reference's ``process_dataset``:

    * ``version_photometry``: a directory holding ``<name>.LIST`` and one SNANA text file per object (:2577-2753),
      incl. the SNANA batch subsetting (``jobsplit``: job j of n takes every n-th object, :2586-2587) and the
      ``SIM_*`` columns of simulated data;
    * ``data_table`` + ``data_root``: a whitespace table with columns ``SNID``, ``files`` (comma-separated SNANA text
      files relative to ``data_root``), ``REDSHIFT_CMB``, ``REDSHIFT_CMB_ERR`` and optionally ``SEARCH_PEAKMJD``
      (:2754-2925); the light curves of an object's files are concatenated.

    Returns dict(obs (10, N_obs, N) with FLUXCAL (fitting) or MAG (training) in rows 1-2, band_inds, sn_names, peak_mjds,
    lcplot frames, fitres columns in the reference's order).  Zero padding with mask 0, non-positive fluxes masked for
    training (:2714-2718, :2892-2897).  SNANA FITS input is out of scope."""
    training = 'training' in args['mode'].lower()
    fmap = dict(args.get('map', {}))
    used = [0]
    lcs, names, peaks = [], [], []
    cols = {k: [] for k in FITRES_COLUMNS}
    sim_cols = {k: [] for k, _ in SIM_COLUMNS}
    survey, sim = 'NULL', False
    sigma_pec = getattr(model, 'sigma_pec', 150 / 3e5)

    def add_fitres(meta, obj, zhel, zhel_err, zhd, zhd_err, idsurvey):
        snr = _snrmax(obj['flux'], obj['ferr'], obj['gl']) if len(obj['flux']) else [-99.0, -99.0, -99.0]
        vals = dict(IDSURVEY=idsurvey, TYPE=meta.get('TYPE', 0), FIELD=str(meta.get('FIELD', 'VOID')), zHEL=zhel,
                    zHELERR=float(meta.get('REDSHIFT_HELIO_ERR', zhel_err)), zHD=zhd, zHDERR=zhd_err,
                    VPEC=float(meta.get('VPEC', 0.0)), VPECERR=float(meta.get('VPEC_ERR', sigma_pec * 3e5)),
                    MWEBV=float(meta.get('MWEBV', 0.0)), HOST_LOGMASS=float(meta.get('HOSTGAL_LOGMASS', -9.0)),
                    HOST_LOGMASS_ERR=float(meta.get('HOSTGAL_LOGMASS_ERR', -9.0)), SNRMAX1=round(snr[0], 2),
                    SNRMAX2=round(snr[1], 2), SNRMAX3=round(snr[2], 2))
        for k in FITRES_COLUMNS:
            cols[k].append(vals[k])

    if 'version_photometry' not in args and 'data_table' not in args:
        raise ValueError('Please pass either data_dir (for a directory containing all SNANA files such as a '
                         'simulation output) or a combination of data_table and data_root')
    if 'version_photometry' in args:
        data_dir = resolve_photometry_dir(args)
        name = os.path.split(os.path.normpath(data_dir))[-1]
        sn_files = np.atleast_1d(np.loadtxt(os.path.join(data_dir, f'{name}.LIST'), dtype=str))
        meta0, _ = snana.read_snana_ascii(os.path.join(data_dir, str(sn_files[0])))
        sim = 'SIM_REDSHIFT_HELIO' in meta0                                       # :2581
        jobid, njob = int(args.get('jobid', 1)), int(args.get('njobtot', 1))
        if not sim:
            njob = int(args.get('jobsplit', [1, 1])[1])                           # real data ignore sim_prescale
        for ind, fn in enumerate(sn_files):
            if (ind + 1 - jobid) % njob != 0:
                continue
            meta, lc = snana.read_snana_ascii(os.path.join(data_dir, str(fn)))
            survey = str(meta.get('SURVEY', 'NULL'))          # reference :2688
            peak = meta['PEAKMJD'] if 'PEAKMJD' in meta else meta['SEARCH_PEAKMJD']
            zhel = float(meta['REDSHIFT_HELIO'])
            zcmb = float(meta.get('REDSHIFT_FINAL', meta.get('REDSHIFT_CMB', zhel)))
            zhel_err = float(meta.get('REDSHIFT_HELIO_ERR', 5e-4))
            zcmb_err = float(meta.get('REDSHIFT_FINAL_ERR', 5e-4))
            vpec = float(meta.get('VPEC', 0.0))
            zpec = np.sqrt((1 + vpec / C_KMS) / (1 - vpec / C_KMS)) - 1
            zhd = (1 + zcmb) / (1 + zpec) - 1
            obj = _one_light_curve(model, args, meta, lc, peak, zhel, fmap, used)
            obj.update(z=zhel, zerr=zhel_err, mu=float(model.cosmo.distmod(np.array([zhd]))[0]),
                       ebv=float(meta.get('MWEBV', 0.0)), mass=float(meta.get('HOSTGAL_LOGMASS', -9.0)))
            lcs.append(obj)
            names.append(str(meta.get('SNID', fn)))
            peaks.append(float(peak))
            add_fitres(meta, obj, zhel, zhel_err, zcmb, zcmb_err, survey_id(survey))
            if sim:
                for k, mk in SIM_COLUMNS:
                    sim_cols[k].append(meta.get(mk, -9))
    elif 'data_table' in args:
        if 'data_root' not in args:
            raise ValueError('If using data_table, please also pass data_root (which defines the location that the '
                             'paths in data_table are defined with respect to)')
        import pandas as pd
        table = pd.read_csv(os.path.join(args['data_root'], args['data_table']), comment='#', sep=r'\s+')
        for _, row in table.iterrows():
            parts, meta = [], {}
            for fn in str(row['files']).split(','):
                meta, lc = snana.read_snana_ascii(os.path.join(args['data_root'], fn))
                peak = row['SEARCH_PEAKMJD'] if 'SEARCH_PEAKMJD' in table.columns else meta['SEARCH_PEAKMJD']
                zhel = float(meta['REDSHIFT_HELIO'])
                parts.append(_one_light_curve(model, args, meta, lc, peak, zhel, fmap, used))
            obj = {k: np.concatenate([p_[k] for p_ in parts]) for k in parts[0]}
            zcmb, zcmb_err = float(row['REDSHIFT_CMB']), float(row['REDSHIFT_CMB_ERR'])
            obj.update(z=zhel, zerr=zcmb_err, mu=float(model.cosmo.distmod(np.array([zcmb]))[0]),
                       ebv=float(meta.get('MWEBV', 0.0)), mass=float(meta.get('HOSTGAL_LOGMASS', -9.0)))
            lcs.append(obj)
            names.append(str(row['SNID']))
            peaks.append(float(peak))
            add_fitres(meta, obj, zhel, zcmb_err, zcmb, zcmb_err, meta.get('IDSURVEY', 'NULL'))
            survey = str(meta.get('SURVEY', survey))
    else:
        raise ValueError('Please pass either version_photometry (a directory containing <name>.LIST and SNANA text files) '
                         'or data_table (a table of SNe and their files) with data_root')
    if not lcs:
        raise ValueError('no objects selected from the data set (check jobsplit / the LIST file)')
    band_inds = np.array(used)                                  # NULL band + bands in order of first use
    local = {g: i for i, g in enumerate(used)}
    N, n_obs = len(lcs), max(max(len(l['t']) for l in lcs), 1)
    obs = np.zeros((10, n_obs, N))
    obs[2] = 1 / np.sqrt(2 * np.pi)                              # padded sigma (:2703-2704)
    for i, l in enumerate(lcs):
        n = len(l['t'])
        flux, ferr = l['flux'], l['ferr']
        obs[0, :n, i] = l['t']
        if training:
            good = flux > 0
            with np.errstate(divide='ignore', invalid='ignore'):
                obs[1, :n, i] = np.where(good, 27.5 - 2.5 * np.log10(np.where(good, flux, 1.0)), 0.0)
                obs[2, :n, i] = np.where(good, (2.5 / np.log(10)) * ferr / np.where(good, flux, 1.0), 1 / np.sqrt(2 * np.pi))
            obs[9, :n, i] = good
            obs[0, :n, i] = np.where(good, l['t'], 0.0)
        else:
            obs[1, :n, i], obs[2, :n, i] = flux, ferr
            obs[9, :n, i] = 1
        obs[4, :n, i] = np.where(obs[9, :n, i] != 0, [local[g] for g in l['gl']], 0)
        obs[3, :, i], obs[5, :, i], obs[6, :, i], obs[7, :, i], obs[8, :, i] = l['mass'], l['z'], l['zerr'], l['mu'], l['ebv']
    lcplot = [dict(CID=names[i], MJD=l['mjd'], FLUXCAL=l['flux'], FLUXCALERR=l['ferr'], FLT=l['flt']) for i, l in enumerate(lcs)]
    fitres = {k: np.array(v) for k, v in cols.items()}
    if sim:
        fitres.update({k: np.array(v) for k, v in sim_cols.items()})
    return dict(obs=obs, band_inds=band_inds, sn_names=names, peak_mjds=np.array(peaks), lcplot=lcplot, survey=survey,
                sim=sim, fitres=fitres)


def lcplot_curves(model, ds, cons, lo, hi, save_fit_errors=False, chunk=512):
    """Model light curves for the LCPLOT file (reference :2256-2272): every SN of the shard [lo, hi) on a 2-day
    rest-frame grid in its own bands, from the device draws ``cons`` - the posterior 'mean' curve (which, like the
    reference's ``get_flux_from_chains(mean=True)``, is the first draw: SURVEY Appendix E) or mean / std over all
    draws with ``save_fit_errors``.  One ``bayesn_chain_flux`` launch per ``chunk`` SNe; only (n, bands, phases)
    means / stds leave the device."""
    import pandas as pd
    t = np.arange(model.tau_knots[0], model.tau_knots[-1], 2)
    z_hel, ebv = ds['obs'][5, 0, lo:hi], ds['obs'][8, 0, lo:hi]
    bands = [list(pd.unique(l['FLT'])) for l in ds['lcplot'][lo:hi]]
    nb = max(len(b) for b in ds_bands(ds))
    n = hi - lo
    fm, fe = np.zeros((n, nb, len(t))), np.zeros((n, nb, len(t)))
    flat = lambda v, a, b: v[a:b].transpose(1, 2).reshape(b - a, -1)          # (n, S*C): draw index c + C s
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        sl = (lambda v: flat(v, a, b)) if save_fit_errors else (lambda v: v[a:b, :1, 0])
        eps = cons['eps'][a:b]
        eps = eps.transpose(1, 2).reshape(b - a, -1, eps.shape[-1]) if save_fit_errors else eps[:, :1, 0]
        RV = sl(cons['RV']) if 'RV' in cons else None
        f = model.chain_flux_device(t, bands[a:b], sl(cons['theta']), sl(cons['AV']), sl(cons['tmax']),
                                    sl(cons['mu']) + sl(cons['delM']), eps.contiguous(), z_hel[a:b], ebv[a:b], RV=RV, mag=False)
        fm[a:b, :f.shape[2]] = f.double().mean(dim=1).cpu().numpy()
        fe[a:b, :f.shape[2]] = f.double().std(dim=1, unbiased=False).cpu().numpy()
    return t, fm, fe


def ds_bands(ds):
    import pandas as pd
    return [list(pd.unique(l['FLT'])) for l in ds['lcplot']]


def write_lcplot(model, path, ds, t, fm, fe):
    """``<prefix>.LCPLOT`` (reference :2256-2293): the photometry that was fitted (DATA_FLAG 1) followed
    by the model light curve of every SN (DATA_FLAG 0)."""
    import pandas as pd
    z_hel = ds['obs'][5, 0, :]
    bands = ds_bands(ds)
    frames = []
    for i, l in enumerate(ds['lcplot']):
        d = pd.DataFrame({'CID': l['CID'], 'MJD': l['MJD'], 'FLUXCAL': l['FLUXCAL'], 'FLUXCALERR': l['FLUXCALERR'],
                          'FLT': l['FLT'], 'DATA_FLAG': 1})
        nb = len(bands[i])
        fit = pd.DataFrame({'CID': l['CID'], 'MJD': (ds['peak_mjds'][i] + t * (1 + z_hel[i])).repeat(nb),
                            'FLUXCAL': fm[i, :nb, :].flatten(order='F'), 'FLUXCALERR': fe[i, :nb, :].flatten(order='F'),
                            'FLT': np.tile(bands[i], len(t)), 'DATA_FLAG': 0})
        frames += [d, fit]
    out = pd.concat(frames).sort_values(by=['CID', 'DATA_FLAG', 'MJD'])
    out.to_csv(path, index=False)
    return out


def survey_id(survey):
    """IDSURVEY from $SNDATA_ROOT/SURVEY.DEF (reference :2421-2431); 0 when SNANA is not installed."""
    root = os.environ.get('SNDATA_ROOT')
    if root and os.path.exists(os.path.join(root, 'SURVEY.DEF')):
        with open(os.path.join(root, 'SURVEY.DEF')) as fh:
            for line in fh:
                sp = line.split()
                if len(sp) >= 3 and sp[0] == 'SURVEY:' and sp[1] == survey:
                    return int(sp[2])
    return 0


def fit_dataset(model, ds, args, group=None):
    """The array-level core of ``run(mode='fitting')`` (reference :1809-1849 + ``postprocess`` :2247-2364): NUTS fits of
    every SN of ``ds`` (``obs`` (10, N_obs, N), ``band_inds``, ``sn_names``, ``peak_mjds``, optional ``lcplot`` /
    ``fitres`` frames), device-side summaries, FITRES / LCPLOT / chains files written by rank 0.

    With ``group`` the SNe are sharded over the ranks.  ``ds['obs']`` is either the whole data set on every rank
    (each takes ``shard_bounds(N, rank, world)``) or - ``ds['shard'] = (lo, hi, N)`` - already this rank's block
    (LSST-sized jobs simulate / read only their own block).  Returns (samples, rows dropped from FITRES)."""
    import torch
    from . import outputs
    from .distributed import shard_bounds, gather_rows
    rank, world = 0, 1
    if group is not None:
        rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
    if 'shard' in ds:
        lo, hi, N = ds['shard']
        obs_loc = ds['obs']
    else:
        N = ds['obs'].shape[-1]
        lo, hi = shard_bounds(N, rank, world)
        obs_loc = ds['obs'][:, :, lo:hi]
    t_fit = time.time()
    cons, raw = model.fit_batch(obs_loc, None, ds['band_inds'], num_chains=args['num_chains'],
                                num_warmup=args['num_warmup'], num_samples=args['num_samples'], max_tree_depth=10,
                                seed=args.get('seed', 0), return_device=True, sn_offset=lo,
                                warps_per_chain=args.get('warps_per_chain', 'auto'))
    ds['timing'] = {'fit_batch_s': time.time() - t_fit, 'leapfrogs': float(raw['chain_info'][..., 2].sum()),
                    'divergences': float(raw['chain_info'][..., 3].sum()), 'warps_per_chain': raw.get('warps_per_chain', 1)}
    muhat = obs_loc[7, 0, :]
    outputs.draw_mu_device(model, cons, muhat, seed=args.get('seed', 0), sn_offset=lo)          # reference :2250-2254
    summ = outputs.device_fit_summary(model, cons, muhat)
    keys = list(summ.keys())
    rows = np.stack([summ[k] for k in keys], axis=1)
    meta = np.stack([obs_loc[5, 0, :], obs_loc[8, 0, :], obs_loc[7, 0, :]], axis=1)      # zHEL, MWEBV, MUHAT
    # the LCPLOT table needs the photometry frames of process_dataset (not kept for pre-sharded array input)
    want_lcplot = args.get('lcplot', True) and ds.get('lcplot') is not None and 'shard' not in ds
    if want_lcplot:
        t, fm, fe = lcplot_curves(model, ds, cons, lo, hi, args.get('save_fit_errors', False))
    if world > 1:
        dev = cons['Ds'].device
        g = lambda a: gather_rows(torch.as_tensor(np.ascontiguousarray(a), device=dev), N, group).cpu().numpy()
        rows, meta = g(rows), g(meta)
        if want_lcplot:
            fm, fe = g(fm), g(fe)
    want_chains = args.get('save_chains', True) and not args.get('snana', False)
    if want_chains or world == 1:
        full = {k: (gather_rows(v.contiguous(), N, group) if world > 1 else v) for k, v in cons.items()}
        samples = outputs.host_samples(full) if (want_chains or args.get('return_samples', True)) else {}
        if 'tmax' in samples and ds.get('peak_mjds') is not None and len(ds['peak_mjds']) == N:
            samples['peak_MJD'] = ds['peak_mjds'][None, None, :] + samples['tmax'] * (1 + meta[:, 0][None, None, :])
    else:
        samples = outputs.host_samples(cons) if args.get('return_samples', True) else {}      # this rank's block only
    drop_count = 0
    if rank == 0:
        meta_cols = ds.get('fitres') or {}
        if meta_cols:
            # reference order (:2735-2753, :2918-2923): object metadata, then the fit columns, SIM_* columns last (:2336-2338)
            cols = {k: v for k, v in meta_cols.items() if 'SIM' not in k}
        else:
            cols = {'zHEL': meta[:, 0], 'MWEBV': meta[:, 1], 'MUHAT': meta[:, 2]}
        cols.update({k: rows[:, i] for i, k in enumerate(keys)})
        cols.update({k: v for k, v in meta_cols.items() if 'SIM' in k})
        pre = args['outfile_prefix']
        base = pre if os.path.isabs(pre) else os.path.join(args['outputdir'], pre)
        os.makedirs(os.path.dirname(base) or '.', exist_ok=True)
        names = ds.get('sn_names') or [str(i) for i in range(N)]
        drop_count = outputs.write_fitres(base + '.FITRES.TEXT', names, cols,
                                          version_photometry=args.get('version_photometry', 'NULL'))
        if want_lcplot:
            write_lcplot(model, base + '.LCPLOT', ds, t, fm, fe)
        if want_chains:
            outputs.write_chains(samples, args['outputdir'])
        ds['fitres_columns'] = cols
    return samples, drop_count


def run(model, args, cmd_args=None, group=None):
    """Reference ``SEDmodel.run`` for ``mode: fitting`` (per-object NUTS fits of every SN in the data
    set, :1809-1849) and ``mode: training_globalRV`` (:1766-1770, :1913-1924).  Returns the samples
    dict (reference site names; (chains, draws, ..., N) for fits).

    Multi-GPU (``torchrun``, one process per GPU): with an initialised ``torch.distributed`` default group (or an
    explicit ``group``) every rank fits the block ``shard_bounds(N, rank, world)`` of the data set - no data-path
    collective; per-SN summaries and light-curve tables are reduced on the device and only those rows are gathered
    for the FITRES / LCPLOT files that rank 0 writes.  The full draws are gathered (and ``chains.pkl`` /
    ``fit_summary.csv`` written) only with ``save_chains`` (default, as in the reference; switch it off for
    LSST-sized jobs: 100k SNe are 23 GB of draws)."""
    import torch
    import yaml
    args = parse_yaml_input(model, args, cmd_args)
    mode = args['mode'].lower()
    rank, world = 0, 1
    if group is None and args.get('distributed', True) and torch.distributed.is_available() \
            and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        group = torch.distributed.group.WORLD
    if group is not None:
        rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
    if 'training' in mode:
        apply_training_knots(model, args)
    ds = process_dataset(model, args)
    model.data, model.sn_list, model.peak_mjds = ds['obs'], ds['sn_names'], ds['peak_mjds']
    model.used_band_inds = ds['band_inds']
    model.survey = ds.get('survey', 'NULL')
    model.survey_id = survey_id(model.survey)
    if model.verbose:
        print(f'Current mode: {args["mode"]}\nRunning...')
    start = time.time()
    drop_count = 0
    if mode == 'fitting':
        samples, drop_count = fit_dataset(model, ds, args, group)
    elif mode == 'training_globalrv':
        init = args['initialisation']
        samples, _ = model.train(ds['obs'], None, ds['band_inds'], num_chains=args['num_chains'],
                                 num_warmup=args['num_warmup'], num_samples=args['num_samples'], initialisation=init,
                                 outputdir=args['outputdir'], mode=args['mode'], group=group)
    else:
        raise ValueError("Invalid mode, this implementation runs 'training_globalRV' and 'fitting'")
    model.end_time = time.time()
    if model.verbose:
        print(f'Total inference runtime: {model.end_time - start} seconds')
    if rank != 0:
        return samples
    if args['snana']:
        # SNANA batch job summary (reference :2342-2356)
        n = ds['obs'].shape[-1]
        out_dict = {'ABORT_IF_ZERO': 1, 'SURVEY': model.survey, 'IDSURVEY': int(model.survey_id), 'NEVT_TOT': n,
                    'NEVT_LC_CUTS': n, 'NEVT_LCFIT_CUTS': int(n - drop_count),
                    'CPU_MINUTES': round((model.end_time - model.start_time) / 60, 2)}
        with open(f'{args["outfile_prefix"]}.YAML', 'w') as fh:
            yaml.safe_dump(out_dict, fh)
    if not (mode == 'fitting' and args['snana']):
        with open(os.path.join(args['outputdir'], 'input.yaml'), 'w') as fh:
            yaml.safe_dump({k: (v if isinstance(v, (int, float, str, bool, list, dict, type(None))) else str(v)) for k, v in args.items()}, fh)
    return samples
