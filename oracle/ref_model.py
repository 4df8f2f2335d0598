"""ORACLE (test infrastructure, NOT product code): float64 CPU restatement of BayeSN's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package.  The product path (``bayesn_b200``) never does and
fails loudly if its CUDA library is missing.

The reference (bayesn v0.4.1, ``/root/reference``) is pure Python on numpyro/JAX, neither of
which can be installed here, so it cannot be executed.  This module restates, function by
function, the reference arithmetic in numpy (static tables) and torch float64 (everything that
needs a gradient; ``torch.autograd`` plays the role of JAX reverse-mode AD):

====================================  =======================================================
here                                  reference
====================================  =======================================================
``invKD_irr`` / ``spline_coeffs_irr``  ``bayesn/spline_utils.py:9-46`` / ``:68-137``
``RefSED.__init__``                   ``bayesn/bayesn_model.py:199-258``
``RefSED._setup_band_weights``        ``bayesn/bayesn_model.py:287-497``
``RefSED._load_hsiao_template``       ``bayesn/bayesn_model.py:260-285``
``RefSED._calculate_band_weights``    ``bayesn/bayesn_model.py:499-554``
``RefSED.spline_coeffs_irr_step``     ``bayesn/bayesn_model.py:761-816``
``RefSED.get_spectra``                ``bayesn/bayesn_model.py:556-643``
``RefSED.get_flux_batch``             ``bayesn/bayesn_model.py:645-709``
``RefSED.get_mag_batch``              ``bayesn/bayesn_model.py:711-759``
``RefSED.fit_logp``                   ``fit_model_globalRV :883-945`` (+ ``uniformRV :818-881``,
                                      ``popRV :996-1060``) as an unconstrained log-density
``RefSED.train_logp``                 ``train_model_globalRV :1115-1182``
``RefSED.prepare_fit_data``           ``fit :2061-2106``
``RefSED.simulate_light_curve``       ``simulate_light_curve :3103-3345``
``RefSED.spectrum_on_linear_grid``    ``simulate_spectrum :3067-3087`` (grid rebuild + get_spectra)
====================================  =======================================================

Third-party pieces that are absent from the tree are restated from their published algorithms:
``extinction.fitzpatrick99`` (Fitzpatrick 1999 spline through 9 anchor points, same
construction as the reference's own host-dust block ``:604-625``; **parity unpinned** - no
reference fixture pins the Milky Way curve), ``astropy.cosmology.FlatLambdaCDM.distmod``
(quadrature), numpyro's unconstraining transforms (``exp`` for positive support, affine-sigmoid
for intervals) and LKJ-Cholesky density.

Pinning: the reference has no tests.  What exists is pinned in ``tests/test_oracle.py``: the
trained-model constants, the SN 2016W notebook posterior (``example_fits.ipynb:166-197``) - the
oracle's posterior for the same data must agree within Monte-Carlo error - and finite-difference
checks of the autograd gradient.  log p / grad values themselves are **parity unpinned** by the
reference (it publishes none).
"""
import math

import numpy as np
import torch
from scipy.integrate import quad, simpson
from scipy.interpolate import interp1d

from bayesn_b200 import datafiles   # raw file readers only (pure IO)

torch.set_default_dtype(torch.float64)


# ---------------------------------------------------------------- spline_utils.py:9-46
def invKD_irr(x):
    n = len(x)
    K = np.zeros((n - 2, n - 2))
    D = np.zeros((n - 2, n))
    K[0, 0:2] = [(x[2] - x[0]) / 3, (x[2] - x[1]) / 6]
    K[-1, -2:n - 2] = [(x[n - 2] - x[n - 3]) / 6, (x[n - 1] - x[n - 3]) / 3]
    for j in np.arange(2, n - 2):
        row = j - 1
        K[row, row - 1:row + 2] = [(x[j] - x[j - 1]) / 6, (x[j + 1] - x[j - 1]) / 3, (x[j + 1] - x[j]) / 6]
    for j in np.arange(1, n - 1):
        row = j - 1
        D[row, row:row + 3] = [1. / (x[j] - x[j - 1]), -(1. / (x[j + 1] - x[j]) + 1. / (x[j] - x[j - 1])),
                               1. / (x[j + 1] - x[j])]
    M = np.zeros((n, n))
    M[1:-1, :] = np.linalg.solve(K, D)
    return M


# ---------------------------------------------------------------- spline_utils.py:68-137
def spline_coeffs_irr(x_int, x, invkd):
    n_x_int, n_x = len(x_int), len(x)
    X = np.zeros((n_x_int, n_x))
    xmax, xmin = max(x), min(x)
    for i in range(n_x_int):
        x_now = x_int[i]
        if x_now > xmax:
            h = x[-1] - x[-2]
            a = (x[-1] - x_now) / h
            b = 1 - a
            f = (x_now - x[-1]) * h / 6.0
            X[i, -2] = a
            X[i, -1] = b
            X[i, :] = X[i, :] + f * invkd[-2, :]
        elif x_now < xmin:
            h = x[1] - x[0]
            b = (x_now - x[0]) / h
            a = 1 - b
            f = (x_now - x[0]) * h / 6.0
            X[i, 0] = a
            X[i, 1] = b
            X[i, :] = X[i, :] - f * invkd[1, :]
        else:
            q = np.where(x[0:-1] <= x_now)[0][-1]
            h = x[q + 1] - x[q]
            a = (x[q + 1] - x_now) / h
            b = 1 - a
            c = ((a ** 3 - a) / 6) * h ** 2
            d = ((b ** 3 - b) / 6) * h ** 2
            X[i, q] = a
            X[i, q + 1] = b
            X[i, :] = X[i, :] + c * invkd[q, :] + d * invkd[q + 1, :]
    return X


def distmod(z, H0=73.24, Om0=0.28):
    """astropy FlatLambdaCDM(H0, Om0).distmod(z) (reference :212, :2102): Tcmb0 = 0, so
    E(z) = sqrt(Om0 (1+z)^3 + 1 - Om0)."""
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    out = np.empty_like(z)
    for i, zi in enumerate(z):
        integ, _ = quad(lambda zz: 1.0 / math.sqrt(Om0 * (1 + zz) ** 3 + 1 - Om0), 0.0, zi, epsabs=1e-13,
                        epsrel=1e-13)
        dl = (1 + zi) * 299792.458 / H0 * integ
        out[i] = 5 * math.log10(dl) + 25
    return out


def fitzpatrick99(wave, a_v, r_v=3.1):
    """Restatement of ``extinction.fitzpatrick99`` (Fitzpatrick 1999, PASP 111, 63): cubic spline
    in 1/lambda [um^-1] through the anchor points for lambda >= 2700 A, FM90 UV curve below."""
    x0, gamma, c3, c4, c5 = 4.596, 0.99, 3.23, 0.41, 5.9
    xk = np.array([0.0, 1e4 / 26500., 1e4 / 12200., 1e4 / 6000., 1e4 / 5470., 1e4 / 4670., 1e4 / 4110.,
                   1e4 / 2700., 1e4 / 2600.])
    c2 = -0.824 + 4.717 / r_v
    c1 = 2.030 - 3.007 * c2

    def uv(x):
        x2 = x * x
        y = x2 - x0 * x0
        d = x2 / (y * y + x2 * gamma * gamma)
        k = c1 + c2 * x + c3 * d
        if x >= c5:
            y = x - c5
            k += c4 * (0.5392 * y * y + 0.05644 * y * y * y)
        return k

    yk = np.array([-r_v, 0.26469 * r_v / 3.1 - r_v, 0.82925 * r_v / 3.1 - r_v,
                   -0.422809 + 1.00270 * r_v + 2.13572e-4 * r_v ** 2 - r_v,
                   -5.13540e-2 + 1.00216 * r_v - 7.35778e-5 * r_v ** 2 - r_v,
                   0.700127 + 1.00184 * r_v - 3.32598e-5 * r_v ** 2 - r_v,
                   1.19456 + 1.01707 * r_v - 5.46959e-3 * r_v ** 2 + 7.97809e-4 * r_v ** 3
                   - 4.45636e-5 * r_v ** 4 - r_v,
                   uv(xk[7]), uv(xk[8])])
    wave = np.asarray(wave, dtype=np.float64)
    x = 1e4 / wave
    J = spline_coeffs_irr(x, xk, invKD_irr(xk))
    k = J @ yk
    for i in range(len(x)):
        if wave[i] < 2700.0:
            k[i] = uv(x[i])
    return a_v * (1.0 + k / r_v)


class RefSED:
    """Float64 restatement of the parts of ``SEDmodel`` that are on (or feed) the hot path."""

    def __init__(self, load_model='T21_model', filter_yaml=None, fiducial_cosmology={"H0": 73.24, "Om0": 0.28},
                 bands=None):
        # bayesn_model.py:212-258
        self.cosmo = dict(fiducial_cosmology)
        self.RV_MW = 3.1
        self.sigma_pec = 150 / 3e5
        params = datafiles.load_model_yaml(load_model)
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
        self.filter_yaml = filter_yaml
        self._only_bands = bands     # optional subset (keeps the oracle light in tests)
        self._setup_band_weights()

    def distmod(self, z):
        return distmod(z, **self.cosmo)

    # ------------------------------------------------------------ bayesn_model.py:260-285
    def _load_hsiao_template(self):
        hsiao_phase, hsiao_wave, hsiao_flux = datafiles.load_hsiao()
        KD_l_hsiao = invKD_irr(hsiao_wave)
        self.J_l_T_hsiao = spline_coeffs_irr(self.model_wave, hsiao_wave, KD_l_hsiao)
        self.hsiao_t = hsiao_phase
        self.hsiao_l = hsiao_wave
        self.hsiao_flux = np.matmul(self.J_l_T_hsiao, hsiao_flux.T)      # (300, 105)

    # ------------------------------------------------------------ bayesn_model.py:287-497
    def _setup_band_weights(self):
        self.min_wave = self.l_knots[0]
        self.max_wave = self.l_knots[-1]
        self.spectrum_bins = 300
        self.band_oversampling = 51
        self.max_redshift = 4
        model_log_wave = np.linspace(np.log10(self.min_wave), np.log10(self.max_wave), self.spectrum_bins)
        model_spacing = model_log_wave[1] - model_log_wave[0]
        band_spacing = model_spacing / self.band_oversampling
        band_max_log_wave = np.log10(self.max_wave * (1 + self.max_redshift)) + band_spacing
        band_log_wave = np.arange(np.log10(self.min_wave), band_max_log_wave, band_spacing)
        band_wave = 10 ** band_log_wave

        filter_dict = datafiles.load_filter_dict(self.filter_yaml)
        for key, val in filter_dict['standards'].items():
            lam, f = datafiles.load_standard(val['path'])
            val['lam'], val['f_lam'] = lam, f

        def ab_standard_flam(l):
            return (2.99792458e18 / 1e23) * (l ** -2) * 10 ** (-48.6 / 2.5) * 1e23

        band_weights, zps, offsets = [], [], []
        self.band_dict, self.zp_dict, self.band_lim_dict = {}, {}, {}
        self.band_dict['NULL_BAND'] = 0
        self.zp_dict['NULL_BAND'] = 10
        self.band_lim_dict['NULL_BAND'] = band_wave[0], band_wave[-1]
        band_weights.append(np.ones_like(band_wave))
        zps.append(10)
        offsets.append(0)
        band_ind = 1
        for key, val in filter_dict['filters'].items():
            if self._only_bands is not None and key not in self._only_bands:
                continue
            band, magsys, offset = key, val['magsys'], val['magzero']
            R = datafiles.load_filter_response(val, key)
            band_low_lim = R[np.where(R[:, 1] > 0.01 * R[:, 1].max())[0][0], 0]
            band_up_lim = R[np.where(R[:, 1] > 0.01 * R[:, 1].max())[0][-1], 0]
            band_conv_transmission = np.interp(band_wave, R[:, 0], R[:, 1], left=0, right=0)
            dlamba = np.diff(band_wave)
            dlamba = np.r_[dlamba, dlamba[-1]]
            num = band_wave * band_conv_transmission * dlamba
            denom = np.sum(num)
            band_weights.append(num / denom)
            lam = R[:, 0]
            if magsys == 'ab':
                zp = ab_standard_flam(lam)
            else:
                standard = filter_dict['standards'][magsys]
                zp = interp1d(standard['lam'], standard['f_lam'], kind='cubic')(lam)
            int1 = simpson(lam * zp * R[:, 1], x=lam)
            int2 = simpson(lam * R[:, 1], x=lam)
            zp = 2.5 * np.log10(int1 / int2)
            self.band_dict[band] = band_ind
            self.band_lim_dict[band] = [band_low_lim, band_up_lim]
            self.zp_dict[band] = zp
            zps.append(zp)
            offsets.append(offset)
            band_ind += 1
        self.zps = np.array(zps, dtype=np.float64)
        self.offsets = np.array(offsets, dtype=np.float64)
        self.inv_band_dict = {val: key for key, val in self.band_dict.items()}
        self.band_interpolate_locations = np.arange(0, self.spectrum_bins * self.band_oversampling,
                                                    self.band_oversampling)
        self.band_interpolate_spacing = band_spacing
        self.band_interpolate_weights = np.array(band_weights)
        self.model_wave = 10 ** model_log_wave
        KD_l = invKD_irr(self.l_knots)
        self.J_l_T = spline_coeffs_irr(self.model_wave, self.l_knots, KD_l)
        self.KD_t = invKD_irr(self.tau_knots)
        self._load_hsiao_template()
        self.ZPT = 27.5
        self.xk = np.array([0.0, 1e4 / 26500., 1e4 / 12200., 1e4 / 6000., 1e4 / 5470., 1e4 / 4670., 1e4 / 4110.,
                            1e4 / 2700., 1e4 / 2600.])
        KD_x = invKD_irr(self.xk)
        self.M_fitz_block = spline_coeffs_irr(1e4 / self.model_wave, self.xk, KD_x)
        self._set_uv_indices()

    # ------------------------------------------------------------ bayesn_model.py:475-481
    def _set_uv_indices(self):
        """Bins below 2700 A take the FM90 ultraviolet form of the F99 law in ``get_spectra`` (:627-638)."""
        self.uv_ind1 = self.model_wave < 2700
        self.uv_ind2 = (self.model_wave < 2700) & ((1e4 / self.model_wave) >= 5.9)
        self.uv_ind3 = (1e4 / self.model_wave[self.uv_ind1]) >= 5.9
        self.uv_x = 1e4 / self.model_wave[self.uv_ind1]

    # ------------------------------------------------------------ bayesn_model.py:499-554
    def _calculate_band_weights(self, redshifts, ebv):
        redshifts = np.atleast_1d(np.asarray(redshifts, dtype=np.float64))
        ebv = np.atleast_1d(np.asarray(ebv, dtype=np.float64))
        locs = self.band_interpolate_locations + np.log10(1 + redshifts)[:, None] / self.band_interpolate_spacing
        flat_locs = locs.flatten()
        int_locs = flat_locs.astype(np.int32)
        remainders = flat_locs - int_locs
        start = self.band_interpolate_weights[..., int_locs]
        end = self.band_interpolate_weights[..., int_locs + 1]
        flat_result = remainders * end + (1 - remainders) * start
        weights = flat_result.reshape((-1,) + locs.shape).transpose(1, 2, 0)
        s = np.sum(weights, axis=1)
        weights = weights / s[:, None, :]
        av = self.RV_MW * ebv
        all_lam = np.array(self.model_wave[None, :] * (1 + redshifts[:, None]))
        all_lam = all_lam.flatten(order='F')
        mw_ext = fitzpatrick99(all_lam, 1, self.RV_MW)
        mw_ext = mw_ext.reshape((weights.shape[0], weights.shape[1]), order='F')
        mw_ext = mw_ext * av[:, None]
        mw_ext = np.power(10, -0.4 * mw_ext)
        weights = weights * mw_ext[..., None]
        weights = weights / (1 + redshifts)[:, None, None]
        return weights

    # ------------------------------------------------------------ bayesn_model.py:761-816
    @staticmethod
    def spline_coeffs_irr_step(x_now, x, invkd):
        """Vectorised over ``x_now`` (any shape) -> (..., T); differentiable in ``x_now``."""
        x = torch.as_tensor(x)
        invkd = torch.as_tensor(invkd)
        xn = x_now[..., None]
        T = x.shape[0]
        eye = torch.eye(T)
        up_extrap = (x_now > x[-1]).to(x.dtype)[..., None]
        down_extrap = (x_now < x[0]).to(x.dtype)[..., None]
        interp = 1 - up_extrap - down_extrap

        h = x[-1] - x[-2]
        a = (x[-1] - xn) / h
        b = 1 - a
        f = (xn - x[-1]) * h / 6.0
        X = (a * eye[T - 2] + b * eye[T - 1] + f * invkd[-2, :]) * up_extrap

        h = x[1] - x[0]
        b = (xn - x[0]) / h
        a = 1 - b
        f = (xn - x[0]) * h / 6.0
        X = X + (a * eye[0] + b * eye[1] - f * invkd[1, :]) * down_extrap

        q = torch.argmax((xn < x).to(torch.int64), dim=-1) - 1
        q = torch.where(q < 0, torch.zeros_like(q), q)    # only reachable in the extrapolation branches
        h = (x[q + 1] - x[q])[..., None]
        a = (x[q + 1][..., None] - xn) / h
        b = 1 - a
        c = ((a ** 3 - a) / 6) * h ** 2
        d = ((b ** 3 - b) / 6) * h ** 2
        X = X + (a * eye[q] + b * eye[q + 1] + c * invkd[q, :] + d * invkd[q + 1, :]) * interp
        return X

    def J_t_and_hsiao(self, t):
        """``t`` (N_obs, N) tensor -> J_t (N, T, N_obs), hsiao_interp (3, N_obs, N)   (:920-924)."""
        hsiao_interp = torch.stack([19 + torch.floor(t), 19 + torch.ceil(t), torch.remainder(t, 1)])
        J = self.spline_coeffs_irr_step(t, self.tau_knots, self.KD_t)      # (N_obs, N, T)
        return J.permute(1, 2, 0), hsiao_interp

    # ------------------------------------------------------------ bayesn_model.py:556-643
    def get_spectra(self, theta, AV, W0, W1, eps, RV, J_t, hsiao_interp):
        num_batch = theta.shape[0]
        J_l_T = torch.as_tensor(self.J_l_T)
        hs = torch.as_tensor(self.hsiao_flux)
        xk = self.xk
        W = W0 + theta[..., None, None] * W1 + eps
        WJt = torch.matmul(W, J_t)
        W_grid = torch.matmul(J_l_T, WJt)
        low_hsiao = hs[:, hsiao_interp[0, ...].to(torch.int64)]
        up_hsiao = hs[:, hsiao_interp[1, ...].to(torch.int64)]
        H_grid = ((1 - hsiao_interp[2, :]) * low_hsiao + hsiao_interp[2, :] * up_hsiao).permute(2, 0, 1)
        model_spectra = H_grid * 10 ** (-0.4 * W_grid)
        f99_x0, f99_gamma, f99_c3 = 4.596, 0.99, 3.23
        RV = torch.as_tensor(RV, dtype=torch.float64) * torch.ones(num_batch)
        f99_c2 = -0.824 + 4.717 / RV
        f99_c1 = 2.030 - 3.007 * f99_c2
        f99_d1 = xk[7] ** 2 / ((xk[7] ** 2 - f99_x0 ** 2) ** 2 + (f99_gamma * xk[7]) ** 2)
        f99_d2 = xk[8] ** 2 / ((xk[8] ** 2 - f99_x0 ** 2) ** 2 + (f99_gamma * xk[8]) ** 2)
        yk = torch.stack([
            -RV,
            0.26469 * RV / 3.1 - RV,
            0.82925 * RV / 3.1 - RV,
            -0.422809 + 1.00270 * RV + 2.13572e-4 * RV ** 2 - RV,
            -5.13540e-2 + 1.00216 * RV - 7.35778e-5 * RV ** 2 - RV,
            0.700127 + 1.00184 * RV - 3.32598e-5 * RV ** 2 - RV,
            1.19456 + 1.01707 * RV - 5.46959e-3 * RV ** 2 + 7.97809e-4 * RV ** 3 - 4.45636e-5 * RV ** 4 - RV,
            f99_c1 + f99_c2 * xk[7] + f99_c3 * f99_d1,
            f99_c1 + f99_c2 * xk[8] + f99_c3 * f99_d2], dim=1)            # (num_batch, 9)
        A = AV[..., None] * (1 + (torch.as_tensor(self.M_fitz_block) @ yk.T).T / RV[..., None])
        if self.uv_ind1.any():                                             # :627-638
            f99_c4, f99_c5 = 0.41, 5.9
            uv_x = torch.as_tensor(self.uv_x)
            c2 = -0.824 + 4.717 / RV[..., None]
            c1 = 2.030 - 3.007 * c2
            x2 = uv_x * uv_x
            y = x2 - f99_x0 * f99_x0
            d = x2 / (y * y + x2 * f99_gamma * f99_gamma)
            k = c1 + c2 * uv_x + f99_c3 * d
            A = A.clone()
            A[:, torch.as_tensor(self.uv_ind1)] = AV[..., None] * (1. + k / RV[..., None])
            y = uv_x - f99_c5
            y2 = y * y
            k = k + f99_c4 * (0.5392 * y2 + 0.05644 * y2 * y)
            A[:, torch.as_tensor(self.uv_ind2)] = AV[..., None] * (1. + k[..., torch.as_tensor(self.uv_ind3)] / RV[..., None])
        f_A = 10 ** (-0.4 * A)
        model_spectra = model_spectra * f_A[..., None]
        return model_spectra

    # ------------------------------------------------------------ bayesn_model.py:3067-3087
    def spectrum_on_linear_grid(self, t, dl, theta, AV, eps, RV):
        """The deterministic part of ``simulate_spectrum``: as the reference does, REPLACES this instance's wavelength
        grid and the tables built on it (J_l_T, M_fitz_block, Hsiao) by ones on a linear grid of spacing ``dl`` between
        the first and last wavelength knot, then evaluates ``get_spectra`` at phases ``t`` for every object.  Returns
        (l_r, spectra (N, n_wave, len(t)))."""
        N = len(theta)
        l_r = np.arange(min(self.l_knots), max(self.l_knots) + dl, dl, dtype=float)
        l_r = l_r[l_r <= max(self.l_knots)]
        self.model_wave = l_r
        self._set_uv_indices()
        self.J_l_T = spline_coeffs_irr(self.model_wave, self.l_knots, invKD_irr(self.l_knots))
        self.M_fitz_block = spline_coeffs_irr(1e4 / self.model_wave, self.xk, invKD_irr(self.xk))
        self._load_hsiao_template()
        tt = torch.as_tensor(np.repeat(np.asarray(t, dtype=float)[..., None], N, axis=1))
        J_t, hsiao_interp = self.J_t_and_hsiao(tt)
        tn = lambda a: torch.as_tensor(np.asarray(a, dtype=float))
        spectra = self.get_spectra(tn(theta), tn(AV), tn(self.W0), tn(self.W1), tn(eps), tn(RV), J_t, hsiao_interp)
        return l_r, spectra.numpy()

    # ------------------------------------------------------------ bayesn_model.py:645-709
    def get_flux_batch(self, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights,
                       zps=None, offsets=None):
        num_batch = theta.shape[0]
        num_observations = band_indices.shape[0]
        zps = torch.as_tensor(self.zps if zps is None else zps)
        offsets = torch.as_tensor(self.offsets if offsets is None else offsets)
        model_spectra = self.get_spectra(theta, AV, W0, W1, eps, RV, J_t, hsiao_interp)
        batch_indices = torch.arange(num_batch).repeat_interleave(num_observations)
        obs_band_weights = (weights[batch_indices, :, band_indices.T.flatten()]
                            .reshape((num_batch, num_observations, -1)).permute(0, 2, 1))
        model_flux = torch.sum(model_spectra * obs_band_weights, dim=1).T
        model_flux = model_flux * 10 ** (-0.4 * (M0 + Ds))
        zp_flux = 10 ** (zps[band_indices] / 2.5)
        model_flux = (model_flux / zp_flux) * 10 ** (0.4 * (27.5 - offsets[band_indices]))
        model_flux = model_flux * mask
        return model_flux

    # ------------------------------------------------------------ bayesn_model.py:711-759
    def get_mag_batch(self, *args, **kw):
        mask = args[9]
        model_flux = self.get_flux_batch(*args, **kw)
        model_flux = model_flux + (1 - mask.to(model_flux.dtype)) * 0.01
        model_mag = -2.5 * torch.log10(model_flux) + 27.5
        return model_mag * mask

    # ------------------------------------------------------------ fit :2061-2106
    def prepare_fit_data(self, t, flux, flux_err, filters, z, ebv_mw=0, peak_mjd=None, filt_map={}, drop_bands=[],
                         mag=False):
        drop_bands = list(drop_bands)
        t, flux, flux_err, filters = np.array(t, float), np.array(flux, float), np.array(flux_err, float), np.array(
            filters)
        if mag:
            flux = np.power(10, (27.5 - flux) / 2.5)
            flux_err = (np.log(10) / 2.5) * flux * flux_err
        if peak_mjd is not None:
            t = (t - peak_mjd) / (1 + z)
        keep = (t > self.tau_knots.min()) & (t < self.tau_knots.max())
        flux, flux_err, filters, t = flux[keep], flux_err[keep], filters[keep], t[keep]
        filters = np.array([filt_map.get(f, f) for f in filters])
        for f in np.unique(filters):
            if f not in self.band_dict:
                raise ValueError(f'Filter "{f}" not defined')
            if z > (self.band_lim_dict[f][0] / self.l_knots[0] - 1) or z < (
                    self.band_lim_dict[f][1] / self.l_knots[-1] - 1):
                drop_bands.append(f)
        for f in drop_bands:
            inds = filters != f
            flux, flux_err, filters, t = flux[inds], flux_err[inds], filters[inds], t[inds]
        band_indices = np.array([self.band_dict[f] for f in filters])
        n_data = len(t)
        if n_data == 0:
            raise ValueError('No data in rest-frame phase range covered by model')
        data = np.zeros((10, n_data, 1))
        data[0, :, 0] = t
        data[1, :, 0] = flux
        data[2, :, 0] = flux_err
        data[4, :, 0] = band_indices
        data[5, :, 0] = z
        data[7, :, 0] = self.distmod(z)[0]
        data[8, :, 0] = ebv_mw
        data[9, :, 0] = 1
        band_weights = self._calculate_band_weights(data[-5, 0, :], data[-2, 0, :])
        return data, band_weights

    # ------------------------------------------------------------ unconstrained log-densities
    def eps_from_tform(self, eps_tform, L_Sigma=None):
        """eps_tform (N, n_eps) -> eps (N, L, T) incl. F-order reshape and zero first/last rows
        (:926-932)."""
        L_Sigma = torch.as_tensor(self.L_Sigma) if L_Sigma is None else L_Sigma
        N = eps_tform.shape[0]
        L, T = len(self.l_knots), len(self.tau_knots)
        eps = torch.matmul(L_Sigma, eps_tform.T).T                       # (N, n_eps)
        eps = eps.reshape(N, T, L - 2).permute(0, 2, 1)                  # order='F': vec index j + (L-2) m
        eps_full = torch.zeros((N, L, T))
        eps_full[:, 1:-1, :] = eps
        return eps_full

    def fit_logp(self, q, obs, weights, rv_mode='fixed', fix_tmax=False, fix_theta=False, theta_val=0.0,
                 fix_AV=False, AV_val=0.0, return_flux=False):
        """Unconstrained log joint density of ``fit_model_globalRV`` (rv_mode='fixed', :883-945),
        ``fit_model_popRV`` ('pop', :996-1060) or ``fit_model_uniformRV`` ('uniform', :818-881).

        ``q`` (N, D) holds, per SN, ``[theta, u_AV, u_tmax, Ds, (u_RV,) eps_tform...]`` with
        ``AV = exp(u_AV)``, ``tmax = -10 + 20 sigmoid(u_tmax)``, ``RV`` / ``RV_tform`` =
        ``low + (high-low) sigmoid(u_RV)`` - numpyro's ``biject_to`` for positive / interval
        supports (upstream numpyro, not in the tree) - and the log|det J| terms added, i.e.
        minus numpyro's potential energy.  ``obs`` (10, N_obs, N), ``weights`` (N, 300, n_b).
        Returns log p per SN (N,)."""
        obs = torch.as_tensor(obs)
        weights = torch.as_tensor(weights)
        N = obs.shape[-1]
        n_sc = 4 if rv_mode == 'fixed' else 5
        theta_s, u_AV, u_t, Ds = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        eps_tform = q[:, n_sc:]
        lp = -0.5 * theta_s ** 2 - 0.5 * math.log(2 * math.pi)                         # theta ~ N(0,1)
        AV_s = torch.exp(u_AV)
        lp = lp + (-math.log(self.tauA) - AV_s / self.tauA) + u_AV                      # Exp(1/tauA) + log|J|
        sg = torch.sigmoid(u_t)
        tmax_s = -10 + 20 * sg
        lp = lp + (-math.log(20.0)) + (math.log(20.0) + torch.nn.functional.logsigmoid(u_t)
                                       + torch.nn.functional.logsigmoid(-u_t))         # U(-10,10) + log|J|
        if rv_mode == 'fixed':
            RV = torch.full((N,), float(self.RV))
        else:
            u_R = q[:, 4]
            sr = torch.sigmoid(u_R)
            ljac = torch.nn.functional.logsigmoid(u_R) + torch.nn.functional.logsigmoid(-u_R)
            if rv_mode == 'uniform':
                RV = 1 + 5 * sr                                                        # U(1,6)
                lp = lp - math.log(5.0) + math.log(5.0) + ljac
            else:
                nd = torch.distributions.Normal(0.0, 1.0)
                phi_alpha_R = nd.cdf(torch.tensor((self.trunc_val - self.mu_R) / self.sigma_R))
                RV = self.mu_R + self.sigma_R * nd.icdf(phi_alpha_R + sr * (1 - phi_alpha_R))
                lp = lp + ljac                                                         # U(0,1) + log|J|
        theta = theta_s * (1 - fix_theta) + theta_val * fix_theta
        AV = AV_s * (1 - fix_AV) + AV_val * fix_AV
        tmax = tmax_s * (1 - fix_tmax)
        t = obs[0, ...] - tmax[None, :]
        J_t, hsiao_interp = self.J_t_and_hsiao(t)
        lp = lp + torch.sum(-0.5 * eps_tform ** 2 - 0.5 * math.log(2 * math.pi), dim=1)  # MVN(0, I)
        eps = self.eps_from_tform(eps_tform)
        band_indices = obs[-6].to(torch.int64)
        muhat = obs[-3, 0, :]
        mask = obs[-1].to(torch.bool)
        Ds_err = math.sqrt(5.0 * 5.0 + self.sigma0 * self.sigma0)
        lp = lp + (-0.5 * ((Ds - muhat) / Ds_err) ** 2 - math.log(Ds_err) - 0.5 * math.log(2 * math.pi))
        flux = self.get_flux_batch(self.M0, theta, AV, torch.as_tensor(self.W0), torch.as_tensor(self.W1), eps, Ds, RV,
                                   band_indices, mask, J_t, hsiao_interp, weights)
        sig = obs[2]
        ll = -0.5 * ((obs[1] - flux) / sig) ** 2 - torch.log(sig) - 0.5 * math.log(2 * math.pi)
        lp = lp + torch.sum(torch.where(mask, ll, torch.zeros_like(ll)), dim=0)
        if return_flux:
            return lp, flux
        return lp

    def fit_logp_grad(self, q, obs, weights, **kw):
        q = torch.as_tensor(q, dtype=torch.float64).clone().requires_grad_(True)
        lp = self.fit_logp(q, obs, weights, **kw)
        g, = torch.autograd.grad(lp.sum(), q)
        return lp.detach().numpy(), g.numpy()

    # ---- training (:1115-1182).  Globals are passed in *constrained* form plus the
    # unconstrained pieces whose Jacobians matter; see tests for how parity is organised.
    def train_loglik_terms(self, W0, W1, L_Sigma, sigma0, RV, tauA, theta, AV, eps_tform, Ds, obs, weights,
                           J_t, hsiao_interp):
        """Sum over SNe of the per-SN terms of ``train_model_globalRV`` that depend on the
        globals: AV ~ Exp(1/tauA), Ds ~ N(muhat, sqrt(muhat_err^2 + sigma0^2)), and the
        magnitude-space likelihood (:1153-1182); theta / eps_tform standard-normal priors are
        included too.  All arguments are constrained-space tensors."""
        obs = torch.as_tensor(obs)
        N = obs.shape[-1]
        lp = torch.sum(-0.5 * theta ** 2 - 0.5 * math.log(2 * math.pi))
        lp = lp + torch.sum(-torch.log(tauA) - AV / tauA)
        lp = lp + torch.sum(-0.5 * eps_tform ** 2 - 0.5 * math.log(2 * math.pi))
        eps = self.eps_from_tform(eps_tform, L_Sigma)
        band_indices = obs[-6].to(torch.int64)
        redshift, redshift_error, muhat = obs[-5, 0, :], obs[-4, 0, :], obs[-3, 0, :]
        mask = obs[-1].to(torch.bool)
        muhat_err = 5 / (redshift * math.log(10)) * torch.sqrt(redshift_error ** 2 + self.sigma_pec ** 2)
        Ds_err = torch.sqrt(muhat_err * muhat_err + sigma0 * sigma0)
        lp = lp + torch.sum(-0.5 * ((Ds - muhat) / Ds_err) ** 2 - torch.log(Ds_err) - 0.5 * math.log(2 * math.pi))
        mag = self.get_mag_batch(self.M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp,
                                 torch.as_tensor(weights))
        sig = obs[2]
        ll = -0.5 * ((obs[1] - mag) / sig) ** 2 - torch.log(sig) - 0.5 * math.log(2 * math.pi)
        lp = lp + torch.sum(torch.where(mask, ll, torch.zeros_like(ll)))
        return lp

    # ---- unconstrained joint density of train_model_globalRV (:1115-1182) ---------------------
    def train_dims(self):
        L, T = len(self.l_knots), len(self.tau_knots)
        ne = (L - 2) * T
        return dict(L=L, T=T, n_eps=ne, n_corr=ne * (ne - 1) // 2, G=2 * L * T + ne + ne * (ne - 1) // 2 + 3,
                    Dl=ne + 3)

    @staticmethod
    def corr_cholesky(y, n):
        """numpyro ``CorrCholeskyTransform`` (upstream numpyro/distributions/transforms.py, not in
        the reference tree; restated from its published algorithm): unconstrained vector ``y`` of
        length n(n-1)/2 -> Cholesky factor of a correlation matrix by tanh + signed stick breaking
        (strict lower triangle filled row-major), and the log|det J| numpyro adds for it."""
        t = torch.tanh(y)
        r = torch.zeros((n, n), dtype=y.dtype)
        ii, jj = torch.tril_indices(n, n, -1)
        r = r.index_put((ii, jj), t)
        z = r * r
        z1m_cumprod = torch.cumprod(1 - z, dim=-1)
        shifted = torch.cat([torch.ones((n, 1), dtype=y.dtype), torch.sqrt(z1m_cumprod)[:, :-1]], dim=1)
        Lo = (r + torch.eye(n, dtype=y.dtype)) * shifted
        ii2, jj2 = torch.tril_indices(n, n, -2)
        stick = 0.5 * torch.sum(torch.log(z1m_cumprod[ii2, jj2]))
        tanh_ld = -2 * torch.sum(y + torch.nn.functional.softplus(-2 * y) - math.log(2.0))
        return Lo, stick + tanh_ld

    def train_unpack(self, qg):
        """Unconstrained global vector -> constrained globals and the log prior + log|det J| of the
        global sites.  Layout of ``qg``: [W0 vec_F (L T), W1 vec_F (L T), u_sigmaepsilon (n_eps),
        y_L_Omega (n_eps (n_eps-1)/2), u_sigma0, u_RV, u_tauA]; interval supports use numpyro's
        ``biject_to`` (affine sigmoid).  The LKJ normalising constant is omitted (parameter free)."""
        d = self.train_dims()
        L, T, ne, nc = d['L'], d['T'], d['n_eps'], d['n_corr']
        o = 0
        W0v = qg[o:o + L * T]; o += L * T
        W1v = qg[o:o + L * T]; o += L * T
        u_se = qg[o:o + ne]; o += ne
        y = qg[o:o + nc]; o += nc
        u_s0, u_rv, u_ta = qg[o], qg[o + 1], qg[o + 2]
        lsig = torch.nn.functional.logsigmoid
        lp = torch.sum(-0.5 * W0v ** 2) + torch.sum(-0.5 * W1v ** 2) - L * T * math.log(2 * math.pi)
        W0 = W0v.reshape(T, L).T                     # order='F'
        W1 = W1v.reshape(T, L).T
        half_pi = math.pi / 2
        se = torch.tan(half_pi * torch.sigmoid(u_se))                  # U(0, pi/2) then tan (:1137-1139)
        lp = lp + torch.sum(lsig(u_se) + lsig(-u_se))                   # -log(pi/2) + log|J|
        Lo, ld = self.corr_cholesky(y, ne)
        k = torch.arange(1, ne, dtype=qg.dtype)
        lp = lp + torch.sum((ne - 1 - k) * torch.log(torch.diagonal(Lo)[1:])) + ld     # LKJCholesky(eta = 1)
        L_Sigma = se[:, None] * Lo                                      # diag(sigma) L_Omega (:1141)
        sigma0 = 0.1 * torch.tan(half_pi * torch.sigmoid(u_s0))
        lp = lp + lsig(u_s0) + lsig(-u_s0)
        RV = 1 + 4 * torch.sigmoid(u_rv)                                # U(1, 5) (:1147)
        lp = lp + lsig(u_rv) + lsig(-u_rv)
        tauA = torch.tan(half_pi * torch.sigmoid(u_ta))
        lp = lp + lsig(u_ta) + lsig(-u_ta)
        return dict(W0=W0, W1=W1, sigmaepsilon=se, L_Omega=Lo, L_Sigma=L_Sigma, sigma0=sigma0, RV=RV, tauA=tauA), lp

    def train_logp(self, qg, ql, obs, weights):
        """Unconstrained log joint density of ``train_model_globalRV``.  ``qg`` (G,) as in
        :meth:`train_unpack`; ``ql`` (N, n_eps + 3) per SN ``[eps_tform..., theta, log AV, Ds]``.
        J_t and the Hsiao interpolation are fixed at the data phases (``self.J_t`` :2725)."""
        obs = torch.as_tensor(obs)
        ne = self.train_dims()['n_eps']
        g, lp = self.train_unpack(qg)
        eps_tform, theta, u_AV, Ds = ql[:, :ne], ql[:, ne], ql[:, ne + 1], ql[:, ne + 2]
        AV = torch.exp(u_AV)
        J_t, hsiao_interp = self.J_t_and_hsiao(obs[0])
        lp = lp + torch.sum(u_AV)                                        # log|J| of AV = exp(u)
        lp = lp + self.train_loglik_terms(g['W0'], g['W1'], g['L_Sigma'], g['sigma0'], g['RV'], g['tauA'], theta, AV,
                                          eps_tform, Ds, obs, weights, J_t, hsiao_interp)
        return lp

    def train_logp_grad(self, qg, ql, obs, weights):
        qg = torch.as_tensor(qg, dtype=torch.float64).clone().requires_grad_(True)
        ql = torch.as_tensor(ql, dtype=torch.float64).clone().requires_grad_(True)
        lp = self.train_logp(qg, ql, obs, weights)
        gg, gl = torch.autograd.grad(lp, (qg, ql))
        return float(lp.detach()), gg.numpy(), gl.numpy()

    # ------------------------------------------------------------ simulate_light_curve :3103-3345
    def sample_epsilon(self, N):
        n_sig = (len(self.l_knots) - 2) * len(self.tau_knots)
        eps_tform = np.random.multivariate_normal(np.zeros(n_sig), np.eye(n_sig), N).T
        eps = np.matmul(self.L_Sigma, eps_tform).T
        eps = np.reshape(eps, (N, len(self.l_knots) - 2, len(self.tau_knots)), order='F')
        eps_full = np.zeros((N, len(self.l_knots), len(self.tau_knots)))
        eps_full[:, 1:-1, :] = eps
        return eps_full

    def simulate_light_curve(self, t, N, bands, yerr=0, err_type='mag', z=0, mu=0, ebv_mw=0, RV=None, tmax=0,
                             del_M=None, AV=None, theta=None, eps=None, mag=True, noise=True):
        """Same draw order from the global numpy RNG as the reference (del_M, AV, theta, eps, then
        the observation noise), so seeded runs reproduce."""
        rep = lambda v: np.array(v, dtype=np.float64).repeat(N) if np.ndim(v) == 0 else np.array(v, dtype=np.float64)
        del_M = np.random.normal(0, self.sigma0, N) if del_M is None else rep(del_M)
        AV = np.random.exponential(self.tauA, N) if AV is None else rep(AV)
        theta = np.random.normal(0, 1, N) if theta is None else rep(theta)
        if eps is None:
            eps = self.sample_epsilon(N)
        elif np.ndim(eps) == 0:
            eps = np.zeros((N, len(self.l_knots), len(self.tau_knots)))
        ebv_mw, tmax, z = rep(ebv_mw), rep(tmax), rep(z)
        RV = rep(self.RV if RV is None else RV)
        mu = self.distmod(z) if (isinstance(mu, str) and mu == 'z') else rep(mu)
        param_dict = {'del_M': del_M, 'AV': AV, 'theta': theta, 'eps': eps, 'z': z, 'mu': mu, 'ebv_mw': ebv_mw,
                      'RV': RV}
        t = np.asarray(t, dtype=np.float64)
        if t.shape[0] == np.array(bands).shape[0]:
            band_indices = np.array([self.band_dict[b] for b in bands])
        else:
            num_per_band, num_bands = t.shape[0], len(bands)
            band_indices = np.zeros(num_bands * num_per_band)
            t = np.repeat(t[:, None], num_bands, axis=1).flatten(order='F')
            for i, band in enumerate(bands):
                band_indices[i * num_per_band:(i + 1) * num_per_band] = self.band_dict[band]
        band_indices = band_indices[:, None].repeat(N, axis=1).astype(int)
        mask = np.ones_like(band_indices)
        band_weights = self._calculate_band_weights(z, ebv_mw)
        tt = torch.as_tensor(np.repeat(t[:, None], N, axis=1) - tmax[None, :])
        J_t, hsiao_interp = self.J_t_and_hsiao(tt)
        args = (self.M0, torch.as_tensor(theta), torch.as_tensor(AV), torch.as_tensor(self.W0),
                torch.as_tensor(self.W1), torch.as_tensor(eps), torch.as_tensor(mu + del_M), torch.as_tensor(RV),
                torch.as_tensor(band_indices), torch.as_tensor(mask).to(torch.bool), J_t, hsiao_interp,
                torch.as_tensor(band_weights))
        data = (self.get_mag_batch(*args) if mag else self.get_flux_batch(*args)).numpy()
        yerr = np.array(yerr, dtype=np.float64)
        if err_type == 'mag' and not mag:
            yerr = yerr * (np.log(10) / 2.5) * data
        if yerr.ndim == 0:
            yerr = np.ones_like(data) * yerr
        elif yerr.ndim == 1:
            yerr = np.repeat(yerr[..., None], N, axis=1)
        if noise:
            data = np.random.normal(data, yerr)
        return data, yerr, param_dict, (np.repeat(t[:, None], N, axis=1), band_indices, band_weights)
