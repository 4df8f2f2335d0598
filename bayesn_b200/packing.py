"""Pack the reference's ``obs`` (10, N_obs, N) / ``weights`` (N, 300, n_b) arrays
(``bayesn/bayesn_model.py:2096-2106``, ``:2711-2730``) into the float32 device layout the
kernels read (``include/bayesn_b200.h``: ``bayesn_dataset_desc``)."""
import math

import numpy as np

from .native import NBINS, OVERRUN


def band_log10_scale(zps, offsets):
    """log10 of the FLUXCAL scale of each band: 0.4 (27.5 - offset) - zp / 2.5 (reference :704-707)."""
    return 0.4 * (27.5 - np.asarray(offsets, dtype=np.float64)) - np.asarray(zps, dtype=np.float64) / 2.5


def _tiles_on_host(weights, used, fold, bscale, M0, device, chunk):
    """Compact float32 weight tiles from the reference-layout (N, 300, n_b) float64 array."""
    import torch
    N, n_b = used.shape
    # exact non-zero extent of every (SN, band) weight row after rounding to float32; bands without
    # a valid observation of that SN (always the NULL band 0) get an empty row
    sup_h = np.zeros((N, n_b, 4), dtype=np.int32)
    tiles = []
    bins = np.arange(NBINS)[None, None, :]
    for s0 in range(0, N, chunk):
        w = np.asarray(weights[s0:s0 + chunk], dtype=np.float64)                  # (n, 300, n_b)
        sc = 10.0 ** (bscale[None, :] - 0.4 * (M0 + fold[s0:s0 + chunk, None]))   # (n, n_b)
        w32 = (np.transpose(w, (0, 2, 1)) * sc[:, :, None]).astype(np.float32)     # (n, n_b, 300)
        nz = (w32 != 0) & used[s0:s0 + chunk, :, None]
        any_nz = nz.any(axis=2)
        first = np.where(any_nz, nz.argmax(axis=2), 0)
        last = np.where(any_nz, NBINS - nz[:, :, ::-1].argmax(axis=2), 0)
        width = last - first
        off = np.cumsum(width, axis=1) - width
        sup_h[s0:s0 + chunk, :, 0], sup_h[s0:s0 + chunk, :, 1], sup_h[s0:s0 + chunk, :, 2] = first, last, off
        # compact tile: the rows of one SN back to back over their supports (C3: ~470 floats
        # instead of 9 x 304)
        n = w32.shape[0]
        inside = (bins >= first[:, :, None]) & (bins < last[:, :, None])
        dest = (off[:, :, None] + bins - first[:, :, None])[inside]
        srow = np.broadcast_to(np.arange(n)[:, None, None], inside.shape)[inside]
        tile = np.zeros((n, max(4, int(width.sum(axis=1).max()))), dtype=np.float32)
        tile[srow, dest] = w32[inside]
        tiles.append(tile)
    wt_stride = -(-max(t_.shape[1] for t_ in tiles) // 4) * 4     # 16-byte granular for the TMA copy
    # lanes past the end of a band's support keep reading (and discard) up to OVERRUN floats while
    # the other half-warp finishes a wider band: spare room after the last tile
    wt_store = torch.zeros(N * wt_stride + OVERRUN + 32, dtype=torch.float32, device=device)
    wt = wt_store[:N * wt_stride].view(N, wt_stride)
    s0 = 0
    for tile in tiles:
        wt[s0:s0 + tile.shape[0], :tile.shape[1]] = torch.from_numpy(tile).to(device)
        s0 += tile.shape[0]
    return sup_h, wt, wt_store, wt_stride


class DeviceTileBuilder:
    """Everything the device band-weight builder needs from the host model (reference
    ``_calculate_band_weights`` :499-554): the master curves of the bands in play, the model grid
    and the Milky Way F99 spline at R_V = 3.1."""

    def __init__(self, ctx, filters, band_inds, RV_MW=3.1):
        from .static import F99_XK, f99_yk, inv_kd
        self.ctx = ctx
        self.master = np.ascontiguousarray(filters.master[np.asarray(band_inds)], dtype=np.float64)
        self.model_wave = np.ascontiguousarray(filters.grid.model_wave, dtype=np.float64)
        self.band_spacing = float(filters.grid.band_spacing)
        self.RV_MW = float(RV_MW)
        self.max_redshift = float(getattr(filters.grid, 'max_redshift', 4))
        yk, _ = f99_yk(np.float64(RV_MW))
        self.xk, self.yk, self.mk = F99_XK.copy(), np.asarray(yk), inv_kd(F99_XK) @ yk


def device_tiles(builder, z, ebv, used, fold, bscale, device):
    """Compact tiles straight from (z, E(B-V)) on the GPU: no (N, 300, n_b) array anywhere.  ``used`` (N, n_b)
    bool marks the band rows to build, ``fold`` (N,) is the per-SN exponent (magnitudes) folded into the rows
    together with the band scale.  Returns (support (N, n_b, 4) int32 device tensor, tiles (N, wt_stride) view,
    backing storage, wt_stride)."""
    import torch
    N, n_b = used.shape
    z = np.asarray(z, dtype=np.float64)
    if np.any(z > builder.max_redshift) or np.any(z <= -1):
        raise ValueError(f'redshift outside the range covered by the band master curves (z <= {builder.max_redshift}, '
                         f'reference :290)')
    nbytes = [0]

    def f64(a):
        a = np.array(a, dtype=np.float64, order='C')          # a writable copy (inputs may be broadcast views)
        nbytes[0] += a.nbytes
        return torch.as_tensor(a, device=device)
    bits = (used.astype(np.uint64) << np.arange(n_b, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint32)
    dev = dict(master=f64(builder.master), z=f64(z), ebv=f64(ebv), fold=f64(fold),
               bscale=f64(bscale), used=torch.as_tensor(bits.view(np.int32), device=device),
               model_wave=f64(builder.model_wave))
    sup = torch.zeros((N, n_b, 4), dtype=torch.int32, device=device)
    total = torch.zeros(N, dtype=torch.int32, device=device)
    desc = builder.ctx.tiles_desc(N, n_b, dev, builder)
    builder.ctx.tiles_support(desc, sup, total)
    wt_stride = max(4, -(-int(total.max().item()) // 4) * 4)
    wt_store = torch.zeros(N * wt_stride + OVERRUN + 32, dtype=torch.float32, device=device)
    wt = wt_store[:N * wt_stride].view(N, wt_stride)
    builder.ctx.tiles_fill(desc, sup, wt_stride, wt)
    torch.cuda.synchronize()
    builder.h2d_bytes = nbytes[0] + bits.nbytes
    return sup, wt, wt_store, wt_stride


def _tiles_on_device(builder, obs, used, fold, bscale, device):
    sup, wt, wt_store, wt_stride = device_tiles(builder, obs[-5, 0, :], obs[-2, 0, :], used, fold, bscale, device)
    return sup.cpu().numpy(), wt, wt_store, wt_stride


def pack_dataset(obs, weights, zps, offsets, M0, Ds_prior_sd, tauA, n_eps, device='cuda', chunk=8192, train=False,
                 sigma_pec=150 / 3e5, builder=None):
    """Returns a dict of torch tensors on ``device``.

    * valid (mask != 0) observations are moved to the front of each SN; ``n_valid`` counts them;
    * the band scale 10^(0.4(27.5-offset) - zp/2.5 - 0.4(M0 + muhat)) is folded into the weight
      tile in float64 before rounding to float32, so the kernel never forms 10^27-sized factors;
    * ``support`` is the exact non-zero extent of each (SN, band) row and the offset of that row
      in the SN's compact weight tile (only bands the SN was observed in are stored);
    * ``lconst`` collects every parameter-independent term of the log joint density.

    ``weights=None`` with a :class:`DeviceTileBuilder` builds the tiles on the GPU from the redshifts
    and Milky Way reddenings in ``obs`` (``_calculate_band_weights`` + packing in one kernel pair).

    ``train=True`` packs a training set (``train_model_globalRV``, reference :1153-1182): ``obs[1]``
    / ``obs[2]`` are magnitudes and their errors, the Ds prior width and tau_A are sampled (their
    terms are evaluated on the device) and ``muhat_err2`` carries (5/(z ln10))^2 (z_err^2 + sigma_pec^2).
    """
    import torch
    obs = np.asarray(obs, dtype=np.float64)
    N_obs, N = obs.shape[1], obs.shape[2]
    n_b = len(zps)
    assert weights is None or (weights.shape[0] == N and weights.shape[1] == NBINS and weights.shape[2] == n_b)
    assert weights is not None or builder is not None
    t, y, sig = obs[0].T, obs[1].T, obs[2].T
    band = np.rint(obs[-6].T).astype(np.int32)
    mask = obs[-1].T != 0
    muhat = np.ascontiguousarray(obs[-3, 0, :])
    if np.any(mask & ((band < 0) | (band >= n_b))):
        raise ValueError('band index out of range of the weights array')
    bscale = band_log10_scale(zps, offsets)
    assert bscale.shape[0] == n_b
    # the kernels carry the distance as Ds - float32(muhat): fold exactly that rounded muhat into the
    # weights, otherwise every flux of the SN is off by the rounding (2e-6 mag, which the
    # magnitude-space training gradient amplifies by n_obs / sigma^2)
    muhat_f32 = muhat.astype(np.float32).astype(np.float64)
    if train:
        # magnitudes relative to the SN's mean magnitude (see eval_logp_grad, TRAIN branch)
        nval = np.maximum(mask.sum(axis=1), 1)
        ybar = (np.where(mask, y, 0.0).sum(axis=1) / nval).astype(np.float32).astype(np.float64)
        y = y - ybar[:, None]
        fold = muhat_f32 - (ybar - 27.5)
    else:
        fold = muhat_f32
    if weights is None:
        fold = fold + M0              # the device builder takes the whole per-SN exponent
    used = np.zeros((N, n_b), dtype=bool)
    rows = np.repeat(np.arange(N), N_obs)
    used[rows[mask.ravel()], band.ravel()[mask.ravel()]] = True
    if weights is None:
        sup_h, wt, wt_store, wt_stride = _tiles_on_device(builder, obs, used, fold, bscale, device)
    else:
        sup_h, wt, wt_store, wt_stride = _tiles_on_host(weights, used, fold, bscale, M0, device, chunk)
    # valid observations first, widest band support first: the kernel processes observations in
    # pairs (one per half-warp) and a pair costs as many rounds as its wider member
    band_c = np.clip(band, 0, n_b - 1)
    width = np.take_along_axis(sup_h[:, :, 1] - sup_h[:, :, 0], band_c, axis=1)
    key = np.where(mask, -width.astype(np.int64) * n_b - band_c, np.iinfo(np.int64).max)
    order = np.argsort(key, axis=1, kind='stable')
    take = lambda a: np.take_along_axis(a, order, axis=1)
    t, y, sig, band, mask = take(t), take(y), take(sig), take(band), take(mask)
    n_valid = mask.sum(axis=1).astype(np.int32)
    obs4 = np.zeros((N, N_obs, 4), dtype=np.float32)
    safe_sig = np.where(mask, sig, 1.0)
    obs4[..., 0] = np.where(mask, t, 0.0)
    obs4[..., 1] = np.where(mask, y, 0.0)
    obs4[..., 2] = np.where(mask, 1.0 / safe_sig, 0.0)
    obs4[..., 3] = np.where(mask, band, 0).astype(np.int32).view(np.float32)
    half_l2pi = 0.5 * math.log(2 * math.pi)
    lconst = np.sum(np.where(mask, -np.log(safe_sig) - half_l2pi, 0.0), axis=1)
    if train:
        lconst += -half_l2pi * (2 + n_eps)
        zs, zerr = obs[-5, 0, :], obs[-4, 0, :]
        muhat_err2 = (5 / (zs * math.log(10))) ** 2 * (zerr ** 2 + sigma_pec ** 2)
    else:
        lconst += -half_l2pi * (1 + n_eps) - math.log(tauA) - math.log(Ds_prior_sd) - half_l2pi
        muhat_err2 = None
    h2d = [int(getattr(builder, 'h2d_bytes', 0)) if weights is None else int(wt.numel() * 4)]

    def to(a):
        a = np.ascontiguousarray(a)
        h2d[0] += a.nbytes
        return torch.from_numpy(a).to(device)
    return dict(N=N, N_obs=N_obs, n_b=n_b, wt_stride=wt_stride, obs4=to(obs4), n_valid=to(n_valid), weights=wt,
                support=to(np.ascontiguousarray(sup_h)),
                muhat=to(muhat.astype(np.float32)), lconst=to(lconst.astype(np.float32)),
                muhat64=muhat, order=order, _weights_storage=wt_store,
                muhat_err2=None if muhat_err2 is None else to(muhat_err2.astype(np.float32)), h2d_bytes=h2d)
