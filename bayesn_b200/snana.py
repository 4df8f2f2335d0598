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
    """Write user data to an SNANA-like light-curve file: the reference writer (``bayesn/bayesn_io.py:13-132``), same
    arguments, header keys in the same order (preamble comment, SNID, SOURCE = ``survey``, RA, DEC, MWEBV,
    REDSHIFT_HELIO, REDSHIFT_CMB, REDSHIFT_CMB_ERR, PEAKMJD, FILTERS, NOBS, NVAR), the same six columns
    ``MJD FLT FLUXCAL FLUXCALERR MAG MAGERR`` (zero point 27.5, fluxes rounded to four decimals, values written with
    ``str``), default file name ``snname[_survey][_paper].snana.dat``, the terminating ``END:``.  Like the reference it
    returns the FILE NAME, not the path, and refuses a missing directory.  (The reference builds an astropy table whose
    first column ``VARLIST:`` holds ``OBS:`` and lets sncosmo's plain ASCII writer print ``key value`` meta lines, the
    column names and the rows - the text below.)"""
    import time
    if not (len(mjd) == len(flt) == len(mag) == len(magerr)):
        raise ValueError('Provided columns are not the same length!')
    if not os.path.exists(output_dir):
        raise ValueError('Requested output directory does not exist!')
    mjd, mag, magerr = np.asarray(mjd), np.asarray(mag, dtype=float), np.asarray(magerr, dtype=float)
    fluxcal = 10 ** ((27.5 - mag) / 2.5)
    fluxcalerr = np.round(fluxcal * magerr * np.log(10) / 2.5, 4)
    fluxcal = np.round(fluxcal, 4)
    divider = '-' * 59
    datestamp = time.strftime('%Y.%m.%d', time.localtime())
    timestamp = time.strftime('%H.%M hrs (%Z)', time.localtime())
    lines = [f'# {snname} ', '# SNANA-like file generated from user-provided data',
             '# Zeropoint of the converted SNANA file: 27.5 mag', f'# {divider}', f'# Data table created by: {author}',
             f'# On date: {datestamp} (yyyy.mm.dd); {timestamp}.', '# Script used: BayeSNmodel.io.write_snana_lcfile.py',
             f'# {divider}', f'SNID: {snname}']
    if survey is not None:
        lines.append(f'SOURCE: {survey}')
    if ra is not None:
        lines.append(f'RA: {ra}')
    if dec is not None:
        lines.append(f'DEC: {dec}')
    filters = ','.join(np.unique(np.asarray(flt, dtype=str)))
    lines += [f'MWEBV: {ebv_mw}', f'REDSHIFT_HELIO: {z_helio}', f'REDSHIFT_CMB: {z_cmb}', f'REDSHIFT_CMB_ERR: {z_cmb_err}',
              f'PEAKMJD: {tmax}', f'FILTERS: {filters}', f'# {divider}', f'NOBS: {len(mjd)}', 'NVAR: 6',
              'VARLIST: MJD FLT FLUXCAL FLUXCALERR MAG MAGERR']
    for row in zip(mjd, flt, fluxcal, fluxcalerr, mag, magerr):
        lines.append('OBS: ' + ' '.join(str(v) for v in row))
    if filename is None:
        filename = snname + (survey is not None) * f'_{survey}' + (paper is not None) * f'_{paper}' + '.snana.dat'
    with open(os.path.join(output_dir, filename), 'w') as fh:
        fh.write('\n'.join(lines) + '\nEND:')
    return filename


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
