"""Posterior summaries without numpyro / arviz: the table ``mcmc.print_summary()`` prints
(reference ``bayesn/bayesn_model.py:1921``, ``:2143``) and the ``arviz.summary`` frame written to
``fit_summary.csv`` (``:2161``, ``:2360``) and mined for the FITRES R-hat columns (``:2315-2318``).

Two families, because the reference uses two libraries (neither is installable here, so both are restatements of the
published algorithms - parity unpinned):
* ``split_rhat`` / ``effective_sample_size``: numpyro.diagnostics (``print_summary``): split Gelman-Rubin of BDA3 and
  Geyer's initial monotone sequence on the plain draws;
* ``rank_rhat`` / ``ess_bulk`` / ``ess_tail`` / ``mcse_*``: arviz.summary's defaults, i.e. Vehtari, Gelman, Simpson,
  Carpenter & Buerkner (2021): rank-normalised split R-hat (the larger of the bulk and the folded statistic), bulk / tail
  effective sample sizes and Monte-Carlo standard errors.  arviz's display rounding (r_hat to two decimals, ESS to
  integers) is NOT applied: the frames carry the raw numbers."""
import math

import numpy as np


def _autocov(x):
    """x (chains, draws) -> autocovariance along draws via FFT."""
    n = x.shape[1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=1)[:, :n]
    return ac / n


def split_rhat(x):
    """x (chains, draws) -> split Gelman-Rubin statistic."""
    c, n = x.shape
    h = n // 2
    xs = np.concatenate([x[:, :h], x[:, n - h:]], axis=0)
    n = h
    w = xs.var(axis=1, ddof=1).mean()
    b = n * xs.mean(axis=1).var(ddof=1)
    if w == 0:
        return np.nan
    return float(np.sqrt(((n - 1) / n * w + b / n) / w))


def effective_sample_size(x):
    """x (chains, draws) -> n_eff (numpyro.diagnostics.effective_sample_size)."""
    c, n = x.shape
    gamma = _autocov(x)
    var_within = (gamma[:, 0] * n / (n - 1.0)).mean() if n > 1 else gamma[:, 0].mean()
    var_est = var_within * (n - 1.0) / n
    if c > 1:
        var_est = var_est + x.mean(axis=1).var(ddof=1)
    if var_est == 0:
        return np.nan
    rho = 1.0 - (var_within - gamma.mean(axis=0) * n / (n - 1.0)) / var_est
    rho[0] = 1.0
    k = n // 2
    pair = rho[0:2 * k:2] + rho[1:2 * k:2]
    pair = np.clip(pair, 0, None)
    pair = np.minimum.accumulate(pair)
    tau = -1.0 + 2.0 * pair.sum()
    return float(c * n / tau)


# ------------------------------------------------------------------ arviz.summary's diagnostics (Vehtari et al. 2021)
def _split_chains(x):
    """(chains, draws) -> (2 chains, draws // 2): first and last halves (the middle draw of an odd count is dropped)."""
    h = x.shape[1] // 2
    return np.concatenate([x[:, :h], x[:, x.shape[1] - h:]], axis=0)


def _rank_average(a):
    """1-based ranks of a 1-D array, ties sharing their mean rank (scipy.stats.rankdata(method='average'))."""
    order = np.argsort(a, kind='stable')
    srt = a[order]
    run = np.cumsum(np.r_[True, srt[1:] != srt[:-1]]) - 1
    counts = np.bincount(run)
    ends = np.cumsum(counts)
    mean_rank = (2 * ends - counts + 1) / 2.0
    r = np.empty(a.shape[0])
    r[order] = mean_rank[run]
    return r


def _z_scale(x):
    """Rank-normalisation: normal scores of the fractional ranks (rank - 3/8) / (S + 1/4) over ALL draws."""
    import torch
    r = _rank_average(np.ravel(x))
    p = (r - 0.375) / (r.shape[0] + 0.25)
    return torch.special.ndtri(torch.from_numpy(p)).numpy().reshape(np.shape(x))


def _rhat_chains(x):
    n = x.shape[1]
    w = x.var(axis=1, ddof=1).mean()
    b = n * x.mean(axis=1).var(ddof=1)
    if not w > 0:
        return np.nan
    return float(np.sqrt((b / w + n - 1) / n))


def rank_rhat(x):
    """x (chains, draws) -> arviz's default ``r_hat`` (method 'rank')."""
    x = np.asarray(x, dtype=float)
    if not np.isfinite(x).all():
        return np.nan
    sp = _split_chains(x)
    bulk = _rhat_chains(_z_scale(sp))
    tail = _rhat_chains(_z_scale(np.abs(sp - np.median(sp))))
    return max(bulk, tail) if bulk == bulk and tail == tail else np.nan


def _ess_chains(x):
    """Effective sample size of (chains, draws) as arviz computes it: Geyer's initial positive sequence on the
    multi-chain autocorrelation estimate, then the monotone correction, with the odd-lag improvement of the tail."""
    c, n = x.shape
    if n < 4 or not np.isfinite(x).all():
        return np.nan
    if x.max() - x.min() < np.finfo(float).resolution:          # a constant site: every draw counts (as arviz does)
        return float(c * n)
    acov = _autocov(x)
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if c > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    if not var_plus > 0:
        return np.nan
    rho_all = 1.0 - (mean_var - acov.mean(axis=0)) / var_plus
    rho = np.zeros(n)
    even, odd = 1.0, rho_all[1]
    rho[0], rho[1] = even, odd
    t = 1
    while t < n - 3 and even + odd > 0.0:
        even, odd = rho_all[t + 1], rho_all[t + 2]
        if even + odd >= 0:
            rho[t + 1], rho[t + 2] = even, odd
        t += 2
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    total = c * n
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + rho[max_t + 1:max_t + 2].sum()
    tau = max(tau, 1.0 / np.log10(total))
    return float(total / tau)


def ess_bulk(x):
    return _ess_chains(_z_scale(_split_chains(np.asarray(x, dtype=float))))


def ess_tail(x, prob=(0.05, 0.95)):
    """The smaller of the effective sample sizes of the indicators of the lower and the upper tail quantile."""
    x = np.asarray(x, dtype=float)
    return min(_ess_chains(_split_chains((x <= np.quantile(x, q)).astype(float))) for q in prob)


def ess_mean(x):
    return _ess_chains(_split_chains(np.asarray(x, dtype=float)))


def ess_sd(x):
    sp = _split_chains(np.asarray(x, dtype=float))
    return min(_ess_chains(sp), _ess_chains(sp ** 2))


def mcse_mean(x):
    return float(np.std(x, ddof=1) / np.sqrt(ess_mean(x)))


def mcse_sd(x):
    e = ess_sd(x)
    return float(np.std(x, ddof=1) * np.sqrt(np.e * (1.0 - 1.0 / e) ** (e - 1.0) - 1.0))


def hpdi(x, prob=0.9):
    x = np.sort(np.ravel(x))
    n = x.size
    m = int(np.floor(prob * n))
    if m >= n:
        return x[0], x[-1]
    widths = x[m:] - x[:n - m]
    i = int(np.argmin(widths))
    return x[i], x[i + m]


def _flatten_sites(samples):
    """dict of arrays (chains, draws, ...) -> list of (label, (chains, draws))."""
    out = []
    for name in sorted(samples.keys()):
        v = np.asarray(samples[name])
        if v.ndim < 2:
            continue
        rest = v.shape[2:]
        if not rest:
            out.append((name, v))
            continue
        for idx in np.ndindex(*rest):
            out.append((f'{name}[{",".join(str(i) for i in idx)}]', v[(slice(None), slice(None)) + idx]))
    return out


def print_summary(samples, n_div=None, prob=0.9):
    lo, hi = f'{50 * (1 - prob):.1f}%', f'{50 * (1 + prob):.1f}%'
    rows = _flatten_sites(samples)
    width = max([len(r[0]) for r in rows] + [10])
    print()
    print(f'{"":>{width}} {"mean":>9} {"std":>9} {"median":>9} {lo:>9} {hi:>9} {"n_eff":>9} {"r_hat":>9}')
    for label, x in rows:
        x = np.asarray(x, dtype=float)
        a, b = hpdi(x, prob)
        print(f'{label:>{width}} {x.mean():9.2f} {x.std(ddof=1):9.2f} {np.median(x):9.2f} {a:9.2f} {b:9.2f} '
              f'{effective_sample_size(x):9.2f} {split_rhat(x):9.2f}')
    if n_div is not None:
        print(f'\nNumber of divergences: {n_div}')


# ------------------------------------------------------------------ the same diagnostics, batched over sites (torch)
# One row per site: x (B, chains, draws) float64 on any device.  These feed FITRES (MEANRHAT / MAXRHAT of every SN) and
# fit_summary.csv; the per-site numpy functions above are the readable statement of the algorithm and their checker.
ROW_CHUNK = 16384


def _split_t(x):
    import torch
    h = x.shape[-1] // 2
    return torch.cat([x[..., :h], x[..., x.shape[-1] - h:]], dim=-2)


def _z_scale_t(x):
    """(B, c, n) -> normal scores of the average-tie ranks over the c n draws of each row."""
    import torch
    B = x.shape[0]
    flat = x.reshape(B, -1)
    M = flat.shape[1]
    srt, order = torch.sort(flat, dim=1, stable=True)
    pos = torch.arange(M, device=x.device).expand(B, M)
    first = torch.ones_like(srt, dtype=torch.bool)
    first[:, 1:] = srt[:, 1:] != srt[:, :-1]                       # first element of a run of equal values
    last = torch.ones_like(first)
    last[:, :-1] = first[:, 1:]
    start = torch.cummax(torch.where(first, pos, torch.zeros_like(pos)), dim=1).values
    end = torch.cummin(torch.where(last, pos, torch.full_like(pos, M - 1)).flip(1), dim=1).values.flip(1)
    rank = torch.empty_like(srt)
    rank.scatter_(1, order, (start + end).to(srt.dtype) / 2 + 1)
    return torch.special.ndtri((rank - 0.375) / (M + 0.25)).reshape(x.shape)


def _rhat_t(z):
    import torch
    n = z.shape[-1]
    w = z.var(dim=-1, unbiased=True).mean(dim=-1)
    b = n * z.mean(dim=-1).var(dim=-1, unbiased=True)
    return torch.sqrt((b / w + n - 1) / n)


def _median_t(flat):
    srt = flat.sort(dim=1).values
    M = flat.shape[1]
    return (srt[:, (M - 1) // 2] + srt[:, M // 2]) / 2


def rank_rhat_t(x):
    """x (B, chains, draws) -> (B,) rank-normalised split R-hat (arviz's default ``r_hat``), float64."""
    import torch
    out = []
    for r0 in range(0, x.shape[0], ROW_CHUNK):
        sp = _split_t(x[r0:r0 + ROW_CHUNK].double())
        bulk = _rhat_t(_z_scale_t(sp))
        med = _median_t(sp.reshape(sp.shape[0], -1))
        tail = _rhat_t(_z_scale_t((sp - med[:, None, None]).abs()))
        out.append(torch.maximum(bulk, tail))
    return torch.cat(out) if out else x.new_zeros((0,), dtype=torch.float64)


def _ess_t(x):
    """(B, c, n) -> (B,) effective sample size, arviz's estimator (``_ess_chains``) with the two sequential passes
    over the lags run for all rows at once."""
    import torch
    B, c, n = x.shape
    if n < 4:
        return torch.full((B,), float('nan'), dtype=x.dtype, device=x.device)
    m = 1 << int(math.ceil(math.log2(2 * n)))
    xc = x - x.mean(dim=-1, keepdim=True)
    f = torch.fft.rfft(xc, n=m, dim=-1)
    acov = torch.fft.irfft(f * f.conj(), n=m, dim=-1)[..., :n] / n
    mean_var = acov[..., 0].mean(dim=-1) * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if c > 1:
        var_plus = var_plus + x.mean(dim=-1).var(dim=-1, unbiased=True)
    rho_all = 1.0 - (mean_var[:, None] - acov.mean(dim=1)) / var_plus[:, None]
    rho = torch.zeros_like(rho_all)
    even, odd = torch.ones_like(mean_var), rho_all[:, 1].clone()
    rho[:, 0], rho[:, 1] = even, odd
    tfin = torch.ones(B, dtype=torch.long, device=x.device)
    active = torch.ones(B, dtype=torch.bool, device=x.device)
    t = 1
    while t < n - 3:                                               # Geyer's initial positive sequence
        active = active & (even + odd > 0)
        if not bool(active.any()):
            break
        e2, o2 = rho_all[:, t + 1], rho_all[:, t + 2]
        even, odd = torch.where(active, e2, even), torch.where(active, o2, odd)
        ok = active & (e2 + o2 >= 0)
        rho[:, t + 1] = torch.where(ok, e2, rho[:, t + 1])
        rho[:, t + 2] = torch.where(ok, o2, rho[:, t + 2])
        tfin = torch.where(active, tfin + 2, tfin)
        t += 2
    max_t = tfin - 2
    idx = (max_t + 1).clamp(max=n - 1)[:, None]
    rho.scatter_(1, idx, torch.where(even > 0, even, rho.gather(1, idx)[:, 0])[:, None])
    t = 1
    while t <= int(max_t.max()) - 2:                               # initial monotone sequence
        on = (t <= max_t - 2) & (rho[:, t + 1] + rho[:, t + 2] > rho[:, t - 1] + rho[:, t])
        half = (rho[:, t - 1] + rho[:, t]) / 2.0
        rho[:, t + 1] = torch.where(on, half, rho[:, t + 1])
        rho[:, t + 2] = torch.where(on, half, rho[:, t + 2])
        t += 2
    cs = torch.cat([torch.zeros(B, 1, dtype=rho.dtype, device=x.device), rho.cumsum(dim=1)], dim=1)
    head = cs.gather(1, (max_t + 1).clamp(min=0)[:, None])[:, 0]                   # sum(rho[:max_t + 1])
    tau = -1.0 + 2.0 * head + rho.gather(1, idx)[:, 0]
    total = float(c * n)
    ess = total / torch.clamp(tau, min=1.0 / math.log10(total))
    flat = x.reshape(B, -1)
    ess = torch.where(torch.isfinite(tau), ess, torch.full_like(ess, float('nan')))
    const = flat.max(dim=1).values - flat.min(dim=1).values < torch.finfo(torch.float64).resolution
    return torch.where(const, torch.full_like(ess, total), ess)


def summary_columns_t(x, hdi_prob=0.94, device=None):
    """x (B, chains, draws) torch tensor -> dict of (B,) float64 tensors with arviz.summary's nine columns.  Rows are
    processed ``ROW_CHUNK`` at a time and only a chunk is moved to ``device`` (default: where ``x`` lives), so the
    draws of a large run can stay in host memory."""
    import torch
    cols = {}
    for r0 in range(0, x.shape[0], ROW_CHUNK):
        xd = x[r0:r0 + ROW_CHUNK].to(device if device is not None else x.device).double()
        B = xd.shape[0]
        flat = xd.reshape(B, -1)
        M = flat.shape[1]
        srt = flat.sort(dim=1).values
        k = int(math.floor(hdi_prob * M))
        if k >= M:
            lo, hi = srt[:, 0], srt[:, -1]
        else:
            i = (srt[:, k:] - srt[:, :M - k]).argmin(dim=1, keepdim=True)
            lo, hi = srt.gather(1, i)[:, 0], srt.gather(1, i + k)[:, 0]
        sd = flat.std(dim=1, unbiased=True)
        sp = _split_t(xd)
        e_mean = _ess_t(sp)
        e_sd = torch.minimum(e_mean, _ess_t(sp ** 2))
        def quant(p):                                              # numpy's default (linear) quantile, from the sorted rows
            pos = p * (M - 1)
            i0 = int(math.floor(pos))
            i1 = min(i0 + 1, M - 1)
            return srt[:, i0] + (pos - i0) * (srt[:, i1] - srt[:, i0])
        q = [quant(0.05), quant(0.95)]
        e_tail = torch.minimum(*[_ess_t(_split_t((xd <= q[j][:, None, None]).double())) for j in range(2)])
        med = (srt[:, (M - 1) // 2] + srt[:, M // 2]) / 2
        rhat = torch.maximum(_rhat_t(_z_scale_t(sp)), _rhat_t(_z_scale_t((sp - med[:, None, None]).abs())))
        part = {'mean': flat.mean(dim=1), 'sd': sd, 'hdi_lo': lo, 'hdi_hi': hi, 'mcse_mean': sd / torch.sqrt(e_mean),
                'mcse_sd': sd * torch.sqrt(math.e * (1.0 - 1.0 / e_sd) ** (e_sd - 1.0) - 1.0),
                'ess_bulk': _ess_t(_z_scale_t(sp)), 'ess_tail': e_tail, 'r_hat': rhat}
        for name, v in part.items():
            cols.setdefault(name, []).append(v)
    return {name: torch.cat(v) for name, v in cols.items()}


def summary_table(samples, hdi_prob=0.94, device=None):
    """pandas DataFrame with arviz.summary's rows (one per scalar of every site, ``name[i,j]`` labels) and columns.
    All sites are evaluated together by ``summary_columns_t`` - on ``device`` (default: the GPU when there is one;
    this is host-side post-processing in the reference, :2161 / :2360, and the CPU gives the same numbers)."""
    import pandas as pd
    import torch
    if device is None:
        device = 'cuda' if torch.cuda.is_available() else 'cpu'
    labels, blocks = [], []
    for name in sorted(samples.keys()):
        v = np.asarray(samples[name], dtype=np.float64)
        if v.ndim < 2:
            continue
        rest = v.shape[2:]
        blocks.append(np.ascontiguousarray(np.moveaxis(v.reshape(v.shape[0], v.shape[1], -1), 2, 0)))    # (rows, C, S)
        labels += [name] if not rest else [f'{name}[{",".join(str(i) for i in idx)}]' for idx in np.ndindex(*rest)]
    lo, hi = f'hdi_{100 * (1 - hdi_prob) / 2:g}%', f'hdi_{100 * (1 + hdi_prob) / 2:g}%'
    names = ['mean', 'sd', lo, hi, 'mcse_mean', 'mcse_sd', 'ess_bulk', 'ess_tail', 'r_hat']
    if not blocks:
        return pd.DataFrame(columns=names)
    cols = summary_columns_t(torch.from_numpy(np.concatenate(blocks, axis=0)), hdi_prob, device=device)
    keys = ['mean', 'sd', 'hdi_lo', 'hdi_hi', 'mcse_mean', 'mcse_sd', 'ess_bulk', 'ess_tail', 'r_hat']
    return pd.DataFrame({n: cols[k].cpu().numpy() for n, k in zip(names, keys)}, index=labels)


def summary_table_per_site(samples, hdi_prob=0.94):
    """``summary_table`` one site at a time through the numpy functions (the checker of the batched path)."""
    import pandas as pd
    recs, idx = [], []
    for label, x in _flatten_sites(samples):
        x = np.asarray(x, dtype=float)
        a, b = hpdi(x, hdi_prob)
        lo, hi = f'hdi_{100 * (1 - hdi_prob) / 2:g}%', f'hdi_{100 * (1 + hdi_prob) / 2:g}%'
        with np.errstate(all='ignore'):
            recs.append({'mean': x.mean(), 'sd': x.std(ddof=1), lo: a, hi: b, 'mcse_mean': mcse_mean(x),
                         'mcse_sd': mcse_sd(x), 'ess_bulk': ess_bulk(x), 'ess_tail': ess_tail(x), 'r_hat': rank_rhat(x)})
        idx.append(label)
    return pd.DataFrame.from_records(recs, index=idx)
