"""Raw readers for BayeSN's static data files (pure IO, no model arithmetic).

The reference reads these through h5py / astropy.io.fits / ruamel.yaml, none of which exist in
this image.  The layouts are fixed, so they are parsed directly:

* ``hsiao.h5`` (reference ``bayesn/bayesn_model.py:271-276``): superblock-v0 HDF5 holding three
  contiguous little-endian float64 datasets.  The dataset offsets are located by walking the
  file for the known dataspace sizes and are checked against the grid the template is defined
  on (phase -19..85 step 1, wave 1000..25000 step 10).
* Vega standard ``alpha_lyr_stis_010.fits`` (reference ``:380-383``): one BINTABLE extension,
  columns WAVELENGTH (f8) and FLUX (f4) big-endian; parsed from the FITS header cards.
* model ``BAYESN.YAML`` and ``filters.yaml`` (reference ``:223-228``, ``:319-320``): PyYAML.
"""
import os

import numpy as np
import yaml

DATA_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
MODEL_DIR = os.path.join(DATA_ROOT, 'model_files')
FILTER_DIR = os.path.join(DATA_ROOT, 'bayesn-filters')
HSIAO_PATH = os.path.join(DATA_ROOT, 'data', 'hsiao.h5')
EXAMPLE_LC = os.path.join(DATA_ROOT, 'data', 'example_lcs', 'Foundation_DR1_2016W.txt')


def load_yaml(path):
    with open(path, 'r') as fh:
        return yaml.safe_load(fh)


def builtin_models():
    return sorted(d for d in os.listdir(MODEL_DIR) if os.path.isdir(os.path.join(MODEL_DIR, d)))


def load_model_yaml(load_model):
    """Same resolution order as the reference (``:221-231``): an existing path wins, then a
    built-in model name."""
    if os.path.exists(load_model) and os.path.isfile(load_model):
        return load_yaml(load_model)
    if load_model in builtin_models():
        return load_yaml(os.path.join(MODEL_DIR, load_model, 'BAYESN.YAML'))
    raise FileNotFoundError(f'Specified model {load_model} does not exist and does not correspond to one '
                            f'of the built-in model {builtin_models()}')


def load_hsiao(path=HSIAO_PATH):
    """Return (phase[105], wave[2401], flux[105, 2401]) as float64."""
    with open(path, 'rb') as fh:
        buf = fh.read()
    if buf[:8] != b'\x89HDF\r\n\x1a\n':
        raise ValueError(f'{path} is not an HDF5 file')
    n_phase, n_wave = 105, 2401
    # The three datasets are stored contiguously in the order phase, wave, flux.  Find the
    # phase vector by its known content, then the others by their own grids.
    phase_ref = np.arange(-19.0, 86.0)
    wave_ref = np.arange(1000.0, 25010.0, 10.0)
    pb, wb = phase_ref.astype('<f8').tobytes(), wave_ref.astype('<f8').tobytes()
    p_off = buf.find(pb)
    w_off = buf.find(wb)
    if p_off < 0 or w_off < 0:
        raise ValueError('hsiao.h5: phase/wave grids not found at the expected values')
    f_bytes = n_phase * n_wave * 8
    # flux is the only remaining contiguous block of that size: it follows the wave vector,
    # aligned to 8 bytes, after the dataset object headers.  Its address is stored in the
    # layout message; rather than walk the b-tree we locate the block by size from the end.
    f_off = len(buf) - f_bytes
    # Some writers put metadata after the raw data; search backwards for a plausible block.
    cand = [f_off, 25408]
    flux = None
    for off in cand:
        if off < 0 or off + f_bytes > len(buf):
            continue
        arr = np.frombuffer(buf, dtype='<f8', count=n_phase * n_wave, offset=off).reshape(n_phase, n_wave)
        if np.all(np.isfinite(arr)) and arr.min() >= 0 and 1e-12 < arr.max() < 1e-6:
            flux = arr
            break
    if flux is None:
        raise ValueError('hsiao.h5: flux block not found')
    phase = np.frombuffer(buf, dtype='<f8', count=n_phase, offset=p_off)
    wave = np.frombuffer(buf, dtype='<f8', count=n_wave, offset=w_off)
    return phase.copy(), wave.copy(), flux.copy()


def read_fits_bintable(path, columns=('WAVELENGTH', 'FLUX')):
    """Minimal FITS BINTABLE reader (first extension), returns dict of float64 arrays."""
    with open(path, 'rb') as fh:
        buf = fh.read()

    def parse_header(start):
        cards = {}
        pos = start
        while True:
            block = buf[pos:pos + 2880]
            if len(block) < 2880:
                raise ValueError('truncated FITS header')
            done = False
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode('ascii', 'replace')
                key = card[:8].strip()
                if key == 'END':
                    done = True
                    break
                if card[8:10] == '= ':
                    val = card[10:].split('/')[0].strip()
                    cards[key] = val.strip("'").strip()
            pos += 2880
            if done:
                return cards, pos

    hdr0, pos = parse_header(0)
    naxis0 = int(hdr0.get('NAXIS', '0'))
    if naxis0 != 0:
        size = abs(int(hdr0['BITPIX'])) // 8
        for i in range(naxis0):
            size *= int(hdr0[f'NAXIS{i + 1}'])
        pos += ((size + 2879) // 2880) * 2880
    hdr, data_start = parse_header(pos)
    if hdr.get('XTENSION') != 'BINTABLE':
        raise ValueError('first extension is not a BINTABLE')
    row_bytes, n_rows, n_fields = int(hdr['NAXIS1']), int(hdr['NAXIS2']), int(hdr['TFIELDS'])
    fmt_map = {'D': ('>f8', 8), 'E': ('>f4', 4), 'I': ('>i2', 2), 'J': ('>i4', 4), 'K': ('>i8', 8),
               'B': ('u1', 1), 'L': ('u1', 1), 'A': ('S1', 1)}
    names, formats, offsets = [], [], []
    off = 0
    for i in range(1, n_fields + 1):
        tform = hdr[f'TFORM{i}']
        rep = ''.join(ch for ch in tform if ch.isdigit())
        rep = int(rep) if rep else 1
        code = ''.join(ch for ch in tform if ch.isalpha())[0]
        dt, nb = fmt_map[code]
        names.append(hdr[f'TTYPE{i}'])
        formats.append(dt if rep == 1 else (dt, (rep,)))
        offsets.append(off)
        off += nb * rep
    dtype = np.dtype({'names': names, 'formats': formats, 'offsets': offsets, 'itemsize': row_bytes})
    table = np.frombuffer(buf, dtype=dtype, count=n_rows, offset=data_start)
    return {c: table[c].astype(np.float64) for c in columns}


def load_filter_dict(filter_yaml=None):
    """Built-in filters.yaml with absolute paths, merged with an optional custom yaml following
    the reference's rules (``bayesn/bayesn_model.py:318-375``): ``standards_root`` /
    ``filters_root`` prefixes, ``$ENV`` expansion of the first path component, paths relative
    to the custom yaml's directory, custom entries overriding built-ins of the same name."""
    fd = load_yaml(os.path.join(FILTER_DIR, 'filters.yaml'))
    for sect in ('standards', 'filters'):
        for key, val in fd[sect].items():
            val['path'] = os.path.join(FILTER_DIR, val['path'])
    if filter_yaml is None:
        return fd
    if not os.path.exists(filter_yaml):
        raise FileNotFoundError(f'Specified filter yaml {filter_yaml} does not exist')
    custom = load_yaml(filter_yaml)
    ydir = os.path.split(os.path.abspath(filter_yaml))[0]

    def resolve(root, rel):
        path = os.path.join(root, rel)
        parts = os.path.normpath(path).split(os.path.sep)
        if parts[0][:1] == '$':
            env = os.getenv(parts[0][1:])
            if env is None:
                raise FileNotFoundError(f'The environment variable {parts[0]} was not found')
            return os.path.join(env, *parts[1:])
        if not os.path.isabs(path):
            return os.path.join(ydir, path)
        return path

    if 'standards' in custom:
        root = custom.get('standards_root', '')
        for key, val in custom['standards'].items():
            val['path'] = resolve(root, val['path'])
            fd['standards'][key] = val
    root = custom.get('filters_root', '')
    for key, val in (custom.get('filters') or {}).items():
        val['path'] = resolve(root, val['path'])
        fd['filters'][key] = val
    return fd


def load_standard(path):
    """(lam, f_lam) of a standard-star spectrum: FITS BINTABLE or two-column text
    (reference ``:378-388``)."""
    if '.fits' in path:
        tab = read_fits_bintable(path)
        return tab['WAVELENGTH'], tab['FLUX']
    txt = np.loadtxt(path)
    return txt[:, 0], txt[:, 1]


def load_filter_response(entry, name='?'):
    """Two-column response, wavelength converted to Angstrom (reference ``:411-421``)."""
    try:
        R = np.loadtxt(entry['path'])
    except Exception:
        raise FileNotFoundError(f'Filter response file {entry["path"]} not found for {name}')
    R = np.array(R[:, :2], dtype=np.float64)
    units = str(entry.get('lam_unit', 'AA')).lower()
    if units == 'nm':
        R[:, 0] = R[:, 0] * 10
    elif units == 'micron':
        R[:, 0] = R[:, 0] * 1e4
    return R
