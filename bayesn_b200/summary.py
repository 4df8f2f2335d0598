"""Posterior summaries without numpyro / arviz: the table ``mcmc.print_summary()`` prints
(reference ``bayesn/bayesn_model.py:1921``, ``:2143``) and the ``arviz.summary`` frame written to
``fit_summary.csv`` (``:2161``, ``:2360``).  Split R-hat and effective sample size follow
Gelman et al. (BDA3) / Geyer's initial monotone sequence, as numpyro.diagnostics does."""
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


def summary_table(samples, hdi_prob=0.94):
    """pandas DataFrame with arviz.summary's column names."""
    import pandas as pd
    recs, idx = [], []
    for label, x in _flatten_sites(samples):
        x = np.asarray(x, dtype=float)
        a, b = hpdi(x, hdi_prob)
        ess = effective_sample_size(x)
        sd = x.std(ddof=1)
        recs.append({'mean': x.mean(), 'sd': sd, 'hdi_3%': a, 'hdi_97%': b,
                     'mcse_mean': sd / np.sqrt(ess) if ess and ess > 0 else np.nan,
                     'mcse_sd': sd / np.sqrt(2 * ess) if ess and ess > 0 else np.nan,
                     'ess_bulk': ess, 'ess_tail': ess, 'r_hat': split_rhat(x)})
        idx.append(label)
    return pd.DataFrame.from_records(recs, index=idx)
