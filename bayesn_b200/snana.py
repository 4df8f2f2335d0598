"""Minimal SNANA text light-curve reader / writer.

Replaces the two sncosmo calls on the path's boundary: ``sncosmo.read_snana_ascii`` (call sites
``bayesn/bayesn_model.py:1976``, ``:2579``) and ``sncosmo.write_lc`` inside
``bayesn/bayesn_io.py:13-132`` (``write_snana_lcfile``).  Format: ``KEY: value [+- err]`` header
lines, one ``VARLIST:`` line naming the columns, ``OBS:`` rows, ``END:``.
"""
import os

import numpy as np


def _convert(tok):
    try:
        return int(tok)
    except ValueError:
        try:
            return float(tok)
        except ValueError:
            return tok


def read_snana_ascii(path):
    """Returns (meta dict, columns dict of numpy arrays)."""
    meta, names, rows = {}, None, []
    with open(path, 'r') as fh:
        for line in fh:
            line = line.split('#')[0].strip()
            if not line or ':' not in line:
                continue
            key, rest = line.split(':', 1)
            key, toks = key.strip(), rest.split()
            if key == 'VARLIST':
                names = toks
            elif key == 'OBS':
                rows.append(toks)
            elif key == 'END':
                break
            elif toks:
                meta[key] = _convert(toks[0])
                if len(toks) >= 3 and toks[1] == '+-':
                    meta[key + '_ERR'] = _convert(toks[2])
    if names is None:
        raise ValueError(f'{path}: no VARLIST line')
    cols = {}
    for j, name in enumerate(names):
        vals = [_convert(r[j]) for r in rows]
        if all(isinstance(v, (int, float)) for v in vals):
            cols[name] = np.array(vals, dtype=float)
        else:
            cols[name] = np.array([str(v) for v in vals])
    return meta, cols


def write_snana_lcfile(output_dir, snname, mjd, flt, mag, magerr, tmax, z_helio, z_cmb, z_cmb_err, ebv_mw, ra=None,
                       dec=None, author='anonymous', survey=None, paper=None, filename=None):
    """Same header keys and OBS columns as the reference writer (``bayesn/bayesn_io.py:13-132``):
    MJD FLT MAG MAGERR."""
    mjd, mag, magerr = np.asarray(mjd, float), np.asarray(mag, float), np.asarray(magerr, float)
    if not (len(mjd) == len(flt) == len(mag) == len(magerr)):
        raise ValueError('Provided columns are not the same length!')
    if filename is None:
        filename = f'{snname}.snana.dat'
    path = os.path.join(output_dir, filename)
    with open(path, 'w') as fh:
        fh.write(f'SNID: {snname}\n')
        fh.write(f'SOURCE: {author}\n' if author else '')
        if survey is not None:
            fh.write(f'SURVEY: {survey}\n')
        if paper is not None:
            fh.write(f'PAPER: {paper}\n')
        if ra is not None:
            fh.write(f'RA: {ra}\n')
        if dec is not None:
            fh.write(f'DEC: {dec}\n')
        fh.write(f'FILTERS: {",".join(sorted(set(flt)))}\n')
        fh.write(f'MWEBV: {ebv_mw}\n')
        fh.write(f'REDSHIFT_HELIO: {z_helio}\n')
        fh.write(f'REDSHIFT_CMB: {z_cmb}\n')
        fh.write(f'REDSHIFT_CMB_ERR: {z_cmb_err}\n')
        fh.write(f'REDSHIFT_FINAL: {z_cmb} +- {z_cmb_err}\n')
        fh.write(f'SEARCH_PEAKMJD: {tmax}\n')
        fh.write(f'NOBS: {len(mjd)}\nNVAR: 4\n')
        fh.write('VARLIST: MJD FLT MAG MAGERR\n')
        for a, b, c, d in zip(mjd, flt, mag, magerr):
            fh.write(f'OBS: {a:.5f} {b} {c:.5f} {d:.5f}\n')
        fh.write('END:\n')
    return path


def write_snana_fluxfile(output_dir, snname, mjd, flt, fluxcal, fluxcalerr, peak_mjd, z_helio, z_final, z_err=5e-4,
                         ebv_mw=0.0, filename=None, extra=None):
    """SNANA text file with FLUXCAL photometry (the layout ``process_dataset`` reads, reference
    :2577-2600: ``PEAKMJD``, ``REDSHIFT_HELIO``, ``REDSHIFT_FINAL``, columns MJD BAND FLUXCAL FLUXCALERR)."""
    if filename is None:
        filename = f'{snname}.DAT'
    path = os.path.join(output_dir, filename)
    with open(path, 'w') as fh:
        fh.write(f'SNID: {snname}\nMWEBV: {ebv_mw}\nREDSHIFT_HELIO: {z_helio} +- {z_err}\n')
        fh.write(f'REDSHIFT_FINAL: {z_final} +- {z_err}\nPEAKMJD: {peak_mjd}\n')
        for k, v in (extra or {}).items():
            fh.write(f'{k}: {v}\n')
        fh.write(f'NOBS: {len(mjd)}\nNVAR: 4\nVARLIST: MJD BAND FLUXCAL FLUXCALERR\n')
        for a, b, c, d in zip(mjd, flt, fluxcal, fluxcalerr):
            fh.write(f'OBS: {a:.5f} {b} {c:.6f} {d:.6f}\n')
        fh.write('END:\n')
    return path
