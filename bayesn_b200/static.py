"""Host-side static precompute for the BayeSN hot path (fp64 numpy, run once per model/dataset).

Everything here is *setup*, not the hot path: it produces the tables the CUDA kernels consume
(spline matrices, Hsiao template on the model grid, Fitzpatrick99 basis, per-SN band weights,
zero points, distance moduli).  It mirrors what the reference computes in
``SEDmodel.__init__`` / ``_setup_band_weights`` / ``_load_hsiao_template`` /
``_calculate_band_weights`` (``bayesn/bayesn_model.py:199-554``) and
``bayesn/spline_utils.py:9-137`` but is written vectorised.
"""
import numpy as np
from scipy.integrate import simpson
from scipy.interpolate import interp1d

from . import datafiles

SPECTRUM_BINS = 300          # reference :295
BAND_OVERSAMPLING = 51       # reference :296
MAX_REDSHIFT = 4             # reference :297
UV_LIMIT_AA = 2700.0         # below this rest wavelength the reference switches to the FM90 UV law (:627-638)
C_AA_PER_S = 2.99792458e18   # astropy.constants.c in Angstrom/s (reference :391)
C_KM_S = 299792.458
ZPT = 27.5                   # reference :489
HSIAO_PHASE0_INDEX = 19      # index of phase 0 in the Hsiao grid (reference :920)


# ----------------------------------------------------------------------------------------
# natural cubic spline operators (bayesn/spline_utils.py:9-46, :68-137)
# ----------------------------------------------------------------------------------------
def inv_kd(x):
    """K^{-1} D for knots ``x``: maps knot values to knot second derivatives (zero at both ends)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    h = np.diff(x)
    K = np.zeros((n - 2, n - 2))
    D = np.zeros((n - 2, n))
    idx = np.arange(n - 2)
    K[idx, idx] = (h[:-1] + h[1:]) / 3.0
    K[idx[:-1], idx[:-1] + 1] = h[1:-1] / 6.0
    K[idx[1:], idx[1:] - 1] = h[1:-1] / 6.0
    D[idx, idx] = 1.0 / h[:-1]
    D[idx, idx + 1] = -(1.0 / h[:-1] + 1.0 / h[1:])
    D[idx, idx + 2] = 1.0 / h[1:]
    M = np.zeros((n, n))
    M[1:-1] = np.linalg.solve(K, D)
    return M


def spline_matrix(x_int, x, kd=None):
    """Dense interpolation matrix J (len(x_int), len(x)) of the natural cubic spline with linear
    extrapolation beyond the end knots."""
    x = np.asarray(x, dtype=np.float64)
    x_int = np.atleast_1d(np.asarray(x_int, dtype=np.float64))
    if kd is None:
        kd = inv_kd(x)
    n = x.size
    J = np.zeros((x_int.size, n))
    rows = np.arange(x_int.size)
    up = x_int > x[-1]
    dn = x_int < x[0]
    mid = ~(up | dn)
    # interior: interval q with x[q] <= x_now (last such), clipped so that q+1 exists
    q = np.clip(np.searchsorted(x, x_int, side='right') - 1, 0, n - 2)
    h = x[q + 1] - x[q]
    a = (x[q + 1] - x_int) / h
    b = 1.0 - a
    c = ((a ** 3 - a) / 6.0) * h ** 2
    d = ((b ** 3 - b) / 6.0) * h ** 2
    m = mid
    J[rows[m], q[m]] += a[m]
    J[rows[m], q[m] + 1] += b[m]
    J[m] += c[m, None] * kd[q[m]] + d[m, None] * kd[q[m] + 1]
    if up.any():
        h = x[-1] - x[-2]
        a = (x[-1] - x_int[up]) / h
        J[rows[up], n - 2] += a
        J[rows[up], n - 1] += 1.0 - a
        J[up] += ((x_int[up] - x[-1]) * h / 6.0)[:, None] * kd[-2]
    if dn.any():
        h = x[1] - x[0]
        b = (x_int[dn] - x[0]) / h
        J[rows[dn], 0] += 1.0 - b
        J[rows[dn], 1] += b
        J[dn] -= ((x_int[dn] - x[0]) * h / 6.0)[:, None] * kd[1]
    return J


def spline_row(t, x, kd):
    """Rows of the time-spline matrix at phases ``t`` (any shape) -> (..., T).
    Same branch structure as ``SEDmodel.spline_coeffs_irr_step`` (reference :761-816): strict
    ``>`` / ``<`` for the extrapolation branches, interior interval from the first knot strictly
    greater than ``t``."""
    t = np.asarray(t, dtype=np.float64)
    return spline_matrix(t.ravel(), x, kd).reshape(t.shape + (len(x),))


# ----------------------------------------------------------------------------------------
# cosmology (astropy FlatLambdaCDM(H0, Om0).distmod, reference :212, :2102)
# ----------------------------------------------------------------------------------------
class FlatLambdaCDM:
    def __init__(self, H0=73.24, Om0=0.28):
        self.H0 = float(H0)
        self.Om0 = float(Om0)
        self._gl = np.polynomial.legendre.leggauss(64)

    def _comoving_integral(self, z):
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        xg, wg = self._gl
        zz = 0.5 * z[:, None] * (xg[None, :] + 1.0)
        E = np.sqrt(self.Om0 * (1 + zz) ** 3 + (1.0 - self.Om0))
        return 0.5 * z * np.sum(wg[None, :] / E, axis=1)

    def luminosity_distance(self, z):
        z = np.asarray(z, dtype=np.float64)
        dc = (C_KM_S / self.H0) * self._comoving_integral(z).reshape(z.shape)
        return (1 + z) * dc

    def distmod(self, z):
        return 5.0 * np.log10(self.luminosity_distance(z)) + 25.0


# ----------------------------------------------------------------------------------------
# Fitzpatrick99 (reference :604-638 for host dust, :544 via `extinction` for Milky Way dust)
# ----------------------------------------------------------------------------------------
F99_XK = np.array([0.0, 1e4 / 26500., 1e4 / 12200., 1e4 / 6000., 1e4 / 5470., 1e4 / 4670., 1e4 / 4110.,
                   1e4 / 2700., 1e4 / 2600.])
F99_X0, F99_GAMMA, F99_C3, F99_C4, F99_C5 = 4.596, 0.99, 3.23, 0.41, 5.9


def f99_yk(RV):
    """Spline knot values k(lambda - V) of the F99 curve for R_V (any shape) -> (..., 9),
    and their derivative with respect to R_V."""
    RV = np.asarray(RV, dtype=np.float64)
    c2 = -0.824 + 4.717 / RV
    c1 = 2.030 - 3.007 * c2
    dc2 = -4.717 / RV ** 2
    dc1 = -3.007 * dc2
    x7, x8 = F99_XK[7], F99_XK[8]
    d1 = x7 ** 2 / ((x7 ** 2 - F99_X0 ** 2) ** 2 + (F99_GAMMA * x7) ** 2)
    d2 = x8 ** 2 / ((x8 ** 2 - F99_X0 ** 2) ** 2 + (F99_GAMMA * x8) ** 2)
    yk = np.stack([
        -RV,
        0.26469 * RV / 3.1 - RV,
        0.82925 * RV / 3.1 - RV,
        -0.422809 + 1.00270 * RV + 2.13572e-4 * RV ** 2 - RV,
        -5.13540e-2 + 1.00216 * RV - 7.35778e-5 * RV ** 2 - RV,
        0.700127 + 1.00184 * RV - 3.32598e-5 * RV ** 2 - RV,
        1.19456 + 1.01707 * RV - 5.46959e-3 * RV ** 2 + 7.97809e-4 * RV ** 3 - 4.45636e-5 * RV ** 4 - RV,
        c1 + c2 * x7 + F99_C3 * d1,
        c1 + c2 * x8 + F99_C3 * d2], axis=-1)
    dyk = np.stack([
        -np.ones_like(RV),
        0.26469 / 3.1 - np.ones_like(RV),
        0.82925 / 3.1 - np.ones_like(RV),
        1.00270 + 2 * 2.13572e-4 * RV - 1,
        1.00216 - 2 * 7.35778e-5 * RV - 1,
        1.00184 - 2 * 3.32598e-5 * RV - 1,
        1.01707 - 2 * 5.46959e-3 * RV + 3 * 7.97809e-4 * RV ** 2 - 4 * 4.45636e-5 * RV ** 3 - 1,
        dc1 + dc2 * x7,
        dc1 + dc2 * x8], axis=-1)
    return yk, dyk


def f99_alav(wave, RV):
    """A(lambda)/A_V of the Fitzpatrick (1999) law at wavelengths ``wave`` (Angstrom), scalar R_V:
    natural cubic spline through the 9 optical/IR anchor points for lambda >= 2700 A, the FM90
    UV parameterisation below (same algorithm as the `extinction` package the reference calls
    for Milky Way dust, and as the reference's own host-dust block)."""
    wave = np.asarray(wave, dtype=np.float64)
    x = 1e4 / wave
    yk, _ = f99_yk(np.float64(RV))
    k = spline_matrix(x.ravel(), F99_XK, inv_kd(F99_XK)) @ yk
    k = k.reshape(x.shape)
    uv = wave < 2700.0
    if np.any(uv):
        xu = x[uv]
        c2 = -0.824 + 4.717 / RV
        c1 = 2.030 - 3.007 * c2
        x2 = xu * xu
        y = x2 - F99_X0 ** 2
        d = x2 / (y * y + x2 * F99_GAMMA ** 2)
        ku = c1 + c2 * xu + F99_C3 * d
        far = xu >= F99_C5
        yy = xu[far] - F99_C5
        ku[far] += F99_C4 * (0.5392 * yy ** 2 + 0.05644 * yy ** 3)
        k[uv] = ku
    return 1.0 + k / RV


def fitz_matrix(wave):
    """(n, 9) matrix M with k(lambda - V)/E(B - V) = M yk(R_V) on the rest-frame wavelengths ``wave`` for EVERY R_V - the
    reference's ``M_fitz_block`` (:313-316) with its FM90 ultraviolet override (``get_spectra`` :627-638) folded in.

    At lambda >= 2700 A the rows are the natural-cubic-spline weights of the nine F99 anchors.  Below, the reference
    overwrites the spline value by k = c1 + c2 x + c3 D(x) (+ c4 F(x) at x >= 5.9), with c2 = -0.824 + 4.717 / R_V and
    c1 = 2.030 - 3.007 c2.  That is P(x) + Q(x) / R_V with
        P = 2.030 + c3 D + c4 F - 0.824 (x - 3.007),    Q = 4.717 (x - 3.007),
    and the two ultraviolet anchors yk[7], yk[8] are the same expression at x = 1e4/2700, 1e4/2600: they span {1, 1/R_V}.
    So the row (0, ..., 0, alpha, beta) with alpha (P7, Q7) + beta (P8, Q8) = (P, Q) reproduces the override exactly and
    every kernel that forms M yk - fixed, per-object or sampled R_V, forward or gradient - gets the branch for free."""
    wave = np.asarray(wave, dtype=np.float64)
    M = spline_matrix(1e4 / wave, F99_XK, inv_kd(F99_XK))
    uv = wave < UV_LIMIT_AA
    if np.any(uv):
        def pq(x):
            x2 = x * x
            y = x2 - F99_X0 ** 2
            P = 2.030 + F99_C3 * x2 / (y * y + x2 * F99_GAMMA ** 2) - 0.824 * (x - 3.007)
            return P, 4.717 * (x - 3.007)
        x = 1e4 / wave[uv]
        P, Q = pq(x)
        far = x >= F99_C5
        yy = x[far] - F99_C5
        P[far] += F99_C4 * (0.5392 * yy ** 2 + 0.05644 * yy ** 3)
        (P7, P8), (Q7, Q8) = pq(F99_XK[7:9])
        det = P7 * Q8 - P8 * Q7
        rows = np.zeros((x.shape[0], 9))
        rows[:, 7] = (P * Q8 - P8 * Q) / det
        rows[:, 8] = (P7 * Q - P * Q7) / det
        M[uv] = rows
    return M


# ----------------------------------------------------------------------------------------
# model grid, Hsiao template, filters
# ----------------------------------------------------------------------------------------
class ModelGrid:
    """Wavelength grids and spline tables for one set of knots (reference :293-316, :475-497)."""

    def __init__(self, l_knots, tau_knots):
        self.l_knots = np.asarray(l_knots, dtype=np.float64)
        self.tau_knots = np.asarray(tau_knots, dtype=np.float64)
        self.max_redshift = MAX_REDSHIFT
        lo, hi = np.log10(self.l_knots[0]), np.log10(self.l_knots[-1])
        self.model_log_wave = np.linspace(lo, hi, SPECTRUM_BINS)
        self.model_spacing = self.model_log_wave[1] - self.model_log_wave[0]
        self.band_spacing = self.model_spacing / BAND_OVERSAMPLING
        band_max = np.log10(self.l_knots[-1] * (1 + MAX_REDSHIFT)) + self.band_spacing
        self.band_log_wave = np.arange(lo, band_max, self.band_spacing)
        self.band_wave = 10 ** self.band_log_wave
        self.model_wave = 10 ** self.model_log_wave
        self.KD_l = inv_kd(self.l_knots)
        self.J_l = spline_matrix(self.model_wave, self.l_knots, self.KD_l)      # (300, L)
        self.KD_t = inv_kd(self.tau_knots)                                      # (T, T)
        self.M_fitz = fitz_matrix(self.model_wave)                               # (300, 9), FM90 UV rows below 2700 A
        ph, hw, hf = datafiles.load_hsiao()
        self.hsiao_phase = ph
        J_h = spline_matrix(self.model_wave, hw, inv_kd(hw))                    # (300, 2401)
        self.hsiao_flux = J_h @ hf.T                                            # (300, 105)

    def dust_basis(self, RV):
        """a[b] = 1 + (M_fitz yk(RV))[b] / RV and da/dRV, so that A[b] = A_V a[b] (reference :625)."""
        RV = np.asarray(RV, dtype=np.float64)
        yk, dyk = f99_yk(RV)
        my = yk @ self.M_fitz.T
        dmy = dyk @ self.M_fitz.T
        a = 1.0 + my / RV[..., None]
        da = dmy / RV[..., None] - my / RV[..., None] ** 2
        return a, da


def ab_standard_flam(lam):
    return (C_AA_PER_S / 1e23) * (lam ** -2) * 10 ** (-48.6 / 2.5) * 1e23


class FilterSet:
    """Band master curves on the oversampled grid, zero points and band limits
    (reference ``_setup_band_weights`` :318-461).  Index 0 is the NULL band."""

    def __init__(self, grid, filter_yaml=None):
        fd = datafiles.load_filter_dict(filter_yaml)
        standards = {}
        for key, val in fd['standards'].items():
            standards[key] = datafiles.load_standard(val['path'])
        bw = grid.band_wave
        dlam = np.diff(bw)
        dlam = np.r_[dlam, dlam[-1]]
        self.band_dict = {'NULL_BAND': 0}
        self.zp_dict = {'NULL_BAND': 10}
        self.band_lim_dict = {'NULL_BAND': (bw[0], bw[-1])}
        weights = [np.ones_like(bw)]
        zps, offsets = [10.0], [0.0]
        for i, (band, val) in enumerate(fd['filters'].items(), start=1):
            R = datafiles.load_filter_response(val, band)
            lam, thr = R[:, 0], R[:, 1]
            above = np.where(thr > 0.01 * thr.max())[0]
            trans = np.interp(bw, lam, thr, left=0, right=0)
            num = bw * trans * dlam
            with np.errstate(divide='ignore', invalid='ignore'):
                weights.append(num / num.sum())
            if val['magsys'] == 'ab':
                fstd = ab_standard_flam(lam)
            else:
                s_lam, s_f = standards[val['magsys']]
                fstd = interp1d(s_lam, s_f, kind='cubic')(lam)
            zp = 2.5 * np.log10(simpson(lam * fstd * thr, x=lam) / simpson(lam * thr, x=lam))
            self.band_dict[band] = i
            self.band_lim_dict[band] = [lam[above[0]], lam[above[-1]]]
            self.zp_dict[band] = zp
            zps.append(zp)
            offsets.append(val['magzero'])
        self.master = np.array(weights)                 # (n_bands+1, n_bandwave)
        self.zps = np.array(zps, dtype=np.float64)
        self.offsets = np.array(offsets, dtype=np.float64)
        self.inv_band_dict = {v: k for k, v in self.band_dict.items()}
        self.grid = grid

    def band_weights(self, redshifts, ebv, band_inds=None, RV_MW=3.1, dtype=np.float64, chunk=4096):
        """Observer-frame weights (N, 300, n_used) incl. Milky Way dust and the 1/(1+z) factor
        (reference ``_calculate_band_weights`` :499-554)."""
        z = np.atleast_1d(np.asarray(redshifts, dtype=np.float64))
        ebv = np.broadcast_to(np.atleast_1d(np.asarray(ebv, dtype=np.float64)), z.shape)
        if np.any(z > MAX_REDSHIFT) or np.any(z <= -1):
            raise ValueError(f'redshift outside the range covered by the band master curves (z <= {MAX_REDSHIFT}, '
                             f'reference :290)')
        if band_inds is None:
            band_inds = np.arange(self.master.shape[0])
        band_inds = np.asarray(band_inds)
        master = self.master[band_inds]
        g = self.grid
        base = np.arange(0, SPECTRUM_BINS * BAND_OVERSAMPLING, BAND_OVERSAMPLING, dtype=np.float64)
        out = np.empty((z.size, SPECTRUM_BINS, band_inds.size), dtype=dtype)
        for s0 in range(0, z.size, chunk):
            zz, ee = z[s0:s0 + chunk], ebv[s0:s0 + chunk]
            locs = base[None, :] + np.log10(1 + zz)[:, None] / g.band_spacing
            il = locs.astype(np.int32)
            rem = locs - il
            w = rem[None] * master[:, il + 1] + (1 - rem[None]) * master[:, il]   # (n_used, n, 300)
            w = np.transpose(w, (1, 2, 0))
            w = w / w.sum(axis=1)[:, None, :]
            lam_obs = g.model_wave[None, :] * (1 + zz[:, None])
            mw = f99_alav(lam_obs, RV_MW) * (RV_MW * ee)[:, None]
            w = w * (10 ** (-0.4 * mw))[..., None]
            w = w / (1 + zz)[:, None, None]
            out[s0:s0 + chunk] = w
        return out
