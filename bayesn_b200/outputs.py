"""Post-processing of batch fits (reference ``postprocess``, ``bayesn/bayesn_model.py:2247-2367``,
fitting branch): distance-modulus draws, per-SN summaries with split R-hat, ``FITRES.TEXT``,
``chains.pkl`` and ``fit_summary.csv``.

The per-SN reductions run on the device on the (N, C, S, D) draw tensor that the NUTS kernel
wrote, so that only ~10 floats per SN leave the GPU (100k SNe x 4 x 250 x 58 draws are 23 GB and
never need to reach the host for a FITRES table).  ``<prefix>.LCPLOT`` is written by
``bayesn_b200.driver.write_lcplot`` (it needs the photometry frame of ``process_dataset``).
"""
import math
import os
import pickle

import numpy as np


def split_rhat_device(x):
    """Classic split-R-hat (Gelman et al. 2013; numpyro.diagnostics.split_gelman_rubin, the ``r_hat`` column of
    ``print_summary``) of ``x`` (..., C, S) torch tensor, reduced over the last two axes on the device.  The FITRES
    columns use the rank-normalised statistic of ``arviz.summary`` instead (``summary.rank_rhat_t``)."""
    import torch
    C, S = x.shape[-2], x.shape[-1]
    h = S // 2
    halves = torch.cat([x[..., :h], x[..., S - h:]], dim=-2).double()          # (..., 2C, h)
    m = halves.mean(dim=-1)
    v = halves.var(dim=-1, unbiased=True)
    W = v.mean(dim=-1)
    B = h * m.var(dim=-1, unbiased=True)
    var_plus = (h - 1) / h * W + B / h
    return torch.sqrt(var_plus / W)


MU_CHUNK = 1024


def draw_mu_device(model, cons, muhat, seed=0, sn_offset=0):
    """Distance-modulus draws of the reference's ``postprocess`` (:2248-2254): ``mu ~ N((Ds 25 + muhat sigma0^2) /
    (25 + sigma0^2), sigma0^2 25 / (25 + sigma0^2))`` for every draw, ``delM = Ds - mu``; added to ``cons`` as
    (N, C, S) device tensors.  The reference uses the unseeded global numpy RNG; here the noise is drawn on the
    device in blocks of ``MU_CHUNK`` SNe keyed by (seed, global block index), so a shard of the data set
    (``sn_offset``) gets the same draws as the whole."""
    import torch
    Ds = cons['Ds']
    dev, (n, C, S) = Ds.device, Ds.shape
    noise = torch.empty((n, C, S), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    first, last = sn_offset // MU_CHUNK, (sn_offset + max(n, 1) - 1) // MU_CHUNK
    for b in range(first, last + 1):
        g.manual_seed((int(seed) << 24) + b)
        blk = torch.randn((MU_CHUNK, C, S), generator=g, device=dev, dtype=torch.float32)
        lo, hi = max(b * MU_CHUNK, sn_offset), min((b + 1) * MU_CHUNK, sn_offset + n)
        noise[lo - sn_offset:hi - sn_offset] = blk[lo - b * MU_CHUNK:hi - b * MU_CHUNK]
    mu_hat = torch.as_tensor(np.asarray(muhat, dtype=np.float64), device=dev)[:, None, None]
    s0, e2 = float(model.sigma0), 25.0
    Dd = Ds.double()
    cons['mu'] = ((Dd * e2 + mu_hat * s0 ** 2) / (e2 + s0 ** 2) + math.sqrt(s0 ** 2 * e2 / (e2 + s0 ** 2)) * noise.double())
    cons['delM'] = Dd - cons['mu']
    return cons


def device_fit_summary(model, cons, muhat, seed=0, sn_offset=0):
    """``cons``: dict of constrained device tensors (N, C, S) from ``SEDmodel.constrain`` /
    ``fit_batch(return_device=True)``.  Uses the ``mu`` / ``delM`` draws in ``cons`` (``draw_mu_device`` adds them
    when missing) and returns a dict of (N,) numpy arrays:
    MU_LCFIT, MUERR_LCFIT, THETA, THETAERR, AV, AVERR, MEANRHAT, MAXRHAT (+ TMAX, TMAXERR)."""
    import torch
    if 'mu' not in cons:
        draw_mu_device(model, cons, muhat, seed=seed, sn_offset=sn_offset)
    mu, Ds = cons['mu'], cons['Ds']
    # MEANRHAT / MAXRHAT: over the r_hat column of arviz.summary (:2315-2318) - rank-normalised split R-hat of every
    # site left after the *_tform ones are dropped: mu, delM, theta, AV, Ds, tmax, peak_MJD, eps (n_eps rows), RV when
    # it is sampled - evaluated for every (site, SN) on the device.  peak_MJD is an increasing affine map of tmax
    # (:2255-2257), so its rank statistic IS tmax's: that row is entered twice instead of building the draws.
    from .summary import rank_rhat_t
    sites = [mu, cons['delM'], cons['theta'], cons['AV'], Ds] + ([cons['RV']] if 'RV' in cons else [])
    rh = [rank_rhat_t(v) for v in sites]
    if 'tmax' in cons:
        rh += [rank_rhat_t(cons['tmax'])] * 2
    rhat = torch.stack(rh)
    if 'eps' in cons:
        e = cons['eps']                                                                     # (N, C, S, n_eps)
        r_eps = rank_rhat_t(e.permute(0, 3, 1, 2).reshape(-1, e.shape[1], e.shape[2])).reshape(e.shape[0], e.shape[3])
        rhat = torch.cat([rhat, r_eps.transpose(0, 1)], dim=0)                              # (rows, N)
    flat = lambda t: t.reshape(t.shape[0], -1).double()
    out = {'MU_LCFIT': flat(mu).mean(dim=1), 'MUERR_LCFIT': flat(mu).std(dim=1, unbiased=False),
           'THETA': flat(cons['theta']).mean(dim=1), 'THETAERR': flat(cons['theta']).std(dim=1, unbiased=False),
           'AV': flat(cons['AV']).mean(dim=1), 'AVERR': flat(cons['AV']).std(dim=1, unbiased=False),
           'MEANRHAT': rhat.mean(dim=0), 'MAXRHAT': rhat.max(dim=0).values}
    if 'tmax' in cons:
        out['TMAX'] = flat(cons['tmax']).mean(dim=1)
        out['TMAXERR'] = flat(cons['tmax']).std(dim=1, unbiased=False)
    return {k: v.cpu().numpy() for k, v in out.items()}


def host_samples(cons):
    """Device draws (N, C, S[, dim]) -> the reference's sample dict: (chains, draws, N) and, for the vector sites,
    (chains, draws, dim, N) (:1839-1848)."""
    out = {}
    for k, v in cons.items():
        v = v.cpu().numpy().astype(np.float64)
        out[k] = v.transpose(1, 2, 3, 0) if v.ndim == 4 else v.transpose(1, 2, 0)
    return out


def _fitres_cell(v):
    """Text of one FITRES cell as the reference produces it: the table is rounded to four decimals
    (``fitres_table.round(4)``, :2330) and every value written with ``str()`` - so 35.1, not 35.1000."""
    if isinstance(v, str):
        return v
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    return str(np.round(np.float64(v), 4))


def write_fitres(path, sn_names, columns, version_photometry='NULL', snana_version='NULL'):
    """SNANA FITRES text table (reference :2296-2340).  The reference hands an astropy table whose first column is
    named ``VARNAMES:`` and holds ``SN:`` (:2734-2742), with the header comment lines smuggled in as meta keys, to
    ``sncosmo.write_lc(..., metachar='')`` - i.e. blank-separated ``key value`` meta lines, one line of column names,
    one line per row: the header comment block, the ``VARNAMES:`` line and one ``SN:`` row per object written here.
    Rows with a NaN distance are dropped (:2332-2333).  Returns the number of rows dropped."""
    names = list(columns.keys())
    keep = ~np.isnan(np.asarray(columns['MU_LCFIT'], dtype=float))
    with open(path, 'w') as fh:
        fh.write(f'#\n# SNANA_VERSION: {snana_version}\n# VERSION_PHOTOMETRY: {version_photometry}\n# TABLE NAME: FITRES\n#\n')
        fh.write('VARNAMES: CID ' + ' '.join(names) + '\n')
        for i, sn in enumerate(sn_names):
            if not keep[i]:
                continue
            vals = ' '.join(_fitres_cell(columns[k][i]) for k in names)
            fh.write(f'SN: {sn} {vals}\n')
    return int((~keep).sum())


def postprocess_fits(model, cons, obs, sn_names=None, outputdir='.', outfile_prefix='output', seed=0,
                     save_chains=False, extra_columns=None, sn_offset=0):
    """Fitting branch of the reference's ``postprocess``: ``mu`` / ``delM`` draws, FITRES table (+ optionally
    ``chains.pkl`` / ``fit_summary.csv`` with the full draws in the reference layout, which is the only step that
    copies them to the host)."""
    obs = np.asarray(obs)
    N = obs.shape[-1]
    sn_names = [str(i) for i in range(N)] if sn_names is None else list(sn_names)
    summ = device_fit_summary(model, cons, obs[-3, 0, :], seed=seed, sn_offset=sn_offset)
    cols = {'zHEL': obs[-5, 0, :], 'MWEBV': obs[-2, 0, :], 'MUHAT': obs[-3, 0, :]}
    if extra_columns:
        cols.update(extra_columns)
    cols.update(summ)
    os.makedirs(outputdir, exist_ok=True)
    dropped = write_fitres(os.path.join(outputdir, f'{outfile_prefix}.FITRES.TEXT'), sn_names, cols)
    if save_chains:
        write_chains(host_samples(cons), outputdir)
    return cols, dropped


def write_chains(samples, outputdir):
    """``chains.pkl`` + ``fit_summary.csv`` (reference :2358-2364; arviz.summary -> bayesn_b200.summary)."""
    from . import summary
    keep = {k: v for k, v in samples.items() if k not in ('diverging', '_auto_latent')}
    summary.summary_table(keep).to_csv(os.path.join(outputdir, 'fit_summary.csv'))
    with open(os.path.join(outputdir, 'chains.pkl'), 'wb') as fh:
        pickle.dump(samples, fh)
