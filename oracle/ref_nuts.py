"""ORACLE (test infrastructure, NOT product code): float64 CPU NUTS with numpyro's semantics.

numpyro is a third-party dependency of the reference (unpinned in ``setup.cfg:14-27``, not in
the tree, not installable here).  Its published algorithm (``numpyro/infer/hmc.py``,
``hmc_util.py``) is restated: multinomial NUTS with iterative tree
building and checkpointed U-turn tests, velocity Verlet, dual-averaging step size
(t0=10, kappa=0.75, gamma=0.05), Stan windows with a regularised diagonal mass matrix,
``init_to_median``.  **Parity unpinned** at the trajectory level (the reference pins only the
SN 2016W posterior summary, ``example_fits.ipynb:166-197``, which ``tests/test_oracle.py``
checks this sampler against).

The random numbers come from the same Philox4x32-10 streams as the CUDA sampler
(``bayesn_b200/csrc/bayesn_kernels.cuh``), so short runs can be compared transition by transition.
Each chain is a Python generator that yields the position it needs log p / gradient at; the
driver evaluates all live chains in one batched oracle call ("lock-step", like the reference's
``jax.vmap`` over whole MCMC runs, ``bayesn/bayesn_model.py:1836``).
"""
import math

import numpy as np

M32 = 0xFFFFFFFF


def philox4x32(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        hi0, lo0 = p0 >> 32, p0 & M32
        hi1, lo1 = p1 >> 32, p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def u01(x):
    return float(x >> 8) * 5.9604644775390625e-8


def adaptation_window_ends(num_steps):
    """numpyro.infer.hmc_util.build_adaptation_schedule."""
    if num_steps <= 0:
        return []
    if num_steps < 20:
        return [num_steps - 1]
    start_buffer, end_buffer, init_window = 75, 50, 25
    if start_buffer + end_buffer + init_window > num_steps:
        start_buffer = int(0.15 * num_steps)
        end_buffer = int(0.1 * num_steps)
        init_window = num_steps - start_buffer - end_buffer
    ends = [start_buffer - 1]
    end_window_start = num_steps - end_buffer
    next_size, next_start = init_window, start_buffer
    while next_start < end_window_start:
        cur_start, cur_size = next_start, next_size
        if 3 * cur_size <= end_window_start - cur_start:
            next_size = 2 * cur_size
        else:
            cur_size = end_window_start - cur_start
        next_start = cur_start + cur_size
        ends.append(next_start - 1)
    ends.append(num_steps - 1)
    return ends


def _is_turning(im, r_left, r_right, r_sum):
    rs = r_sum - 0.5 * (r_left + r_right)
    return (np.dot(im * r_left, rs) <= 0) or (np.dot(im * r_right, rs) <= 0)


def _logaddexp(a, b):
    m = max(a, b)
    if m == -math.inf:
        return -math.inf
    return m + math.log1p(math.exp(-abs(a - b)))


def philox_normals(ctr, key):
    """The four standard normals the device code makes from one Philox block (Box-Muller)."""
    ra = philox4x32(ctr, key)
    n0 = math.sqrt(-2.0 * math.log(u01(ra[0]) + 5.9604644775390625e-8))
    n1 = math.sqrt(-2.0 * math.log(u01(ra[2]) + 5.9604644775390625e-8))
    a0, a1 = 2 * math.pi * u01(ra[1]), 2 * math.pi * u01(ra[3])
    return (n0 * math.cos(a0), n0 * math.sin(a0), n1 * math.cos(a1), n1 * math.sin(a1))


def chain_nuts(z0, unit, num_warmup, num_samples, max_depth=10, target_accept=0.8, init_step=1.0, seed=0,
               adapt_step=True, adapt_mass=True, regularize=True, record=None, nrm_fn=None):
    """Generator: yields positions, receives (logp, grad).  Fills ``record`` (dict of lists).
    ``nrm_fn(it, key)`` (optional) supplies the standard normals of the momentum refresh when the
    state is laid out differently from the per-object fit (training: sharded over SNe)."""
    key = (seed & M32, (seed >> 32) & M32)
    D = z0.shape[0]
    z = z0.astype(np.float64).copy()
    lp, grad = yield z
    pe, g = -lp, -grad
    im = np.ones(D)
    step = init_step
    da_x = da_xavg = da_gavg = 0.0
    da_prox = math.log(10 * step)
    da_t = 0
    wf_n, wmean, wm2 = 0, np.zeros(D), np.zeros(D)
    window = 0
    win_end = adaptation_window_ends(num_warmup)
    n_windows = len(win_end)
    lanes = np.arange(D) % 32
    slots = np.arange(D) // 32
    for it in range(num_warmup + num_samples):
        rk = 0
        # momentum refresh
        nrm = np.empty(D)
        cache = {}
        if nrm_fn is not None:
            nrm = nrm_fn(it, key)
        for i in range(D if nrm_fn is None else 0):
            ln = int(lanes[i])
            if ln not in cache:
                ra = philox4x32((it, 0x40000000, ln, unit & M32), key)
                n0 = math.sqrt(-2.0 * math.log(u01(ra[0]) + 5.9604644775390625e-8))
                n1 = math.sqrt(-2.0 * math.log(u01(ra[2]) + 5.9604644775390625e-8))
                a0, a1 = 2 * math.pi * u01(ra[1]), 2 * math.pi * u01(ra[3])
                cache[ln] = (n0 * math.cos(a0), n0 * math.sin(a0), n1 * math.cos(a1), n1 * math.sin(a1))
            nrm[i] = cache[ln][int(slots[i]) & 3]
        r = nrm / np.sqrt(im)
        E0 = pe + 0.5 * np.dot(im * r, r)
        zl, rl, gl = z.copy(), r.copy(), g.copy()
        zr, rr, gr = z.copy(), r.copy(), g.copy()
        zp, gp, prop_pe = z.copy(), g.copy(), pe
        rsum_main = r.copy()
        w_main, sum_acc = 0.0, 0.0
        depth, n_prop = 0, 0
        turning = diverging = False
        r_ck = np.zeros((max_depth, D))
        rs_ck = np.zeros((max_depth, D))
        while depth < max_depth and not turning and not diverging:
            rd = philox4x32((it, rk, M32, unit & M32), key)
            rk += 1
            going_right = (rd[0] >> 31) != 0
            u_main = u01(rd[1])
            if going_right:
                z, r, g = zr.copy(), rr.copy(), gr.copy()
            else:
                z, r, g = zl.copy(), rl.copy(), gl.copy()
            eps = step if going_right else -step
            n_max, n_sub = 1 << depth, 0
            s_turn = s_div = False
            w_sub, s_acc, sprop_pe = -math.inf, 0.0, 0.0
            szp = sgp = None
            rsum = None
            while n_sub < n_max and not s_turn and not s_div:
                r = r - 0.5 * eps * g
                z = z + eps * im * r
                lp, grad = yield z
                pe_new, g = -lp, -grad
                r = r - 0.5 * eps * g
                E_new = pe_new + 0.5 * np.dot(im * r, r)
                dE = E_new - E0
                if not (dE == dE):
                    dE = math.inf
                leaf_w = -dE
                s_div = dE > 1000.0
                s_acc += min(1.0, math.exp(-dE)) if dE > -700 else 1.0
                if n_sub == 0:
                    rsum = r.copy()
                    w_sub = leaf_w
                    szp, sgp, sprop_pe = z.copy(), g.copy(), pe_new
                else:
                    rsum = rsum + r
                    rl_ = philox4x32((it, rk, M32, unit & M32), key)
                    rk += 1
                    d = leaf_w - w_sub
                    tp = 1.0 / (1.0 + math.exp(-d)) if d > -700 else 0.0
                    if u01(rl_[0]) < tp:
                        szp, sgp, sprop_pe = z.copy(), g.copy(), pe_new
                    w_sub = _logaddexp(w_sub, leaf_w)
                leaf = n_sub
                n_sub += 1
                idx_max = bin(leaf >> 1).count('1')
                n_sub_trees = 0
                x = leaf
                while x & 1:
                    n_sub_trees += 1
                    x >>= 1
                idx_min = idx_max - n_sub_trees + 1
                if leaf % 2 == 0:
                    r_ck[idx_max] = r
                    rs_ck[idx_max] = rsum
                i = idx_max
                while i >= idx_min and not s_turn:
                    sub_rsum = rsum - rs_ck[i] + r_ck[i]
                    s_turn = _is_turning(im, r_ck[i], r, sub_rsum)
                    i -= 1
            if going_right:
                zr, rr, gr = z.copy(), r.copy(), g.copy()
            else:
                zl, rl, gl = z.copy(), r.copy(), g.copy()
            rsum_main = rsum_main + rsum
            if s_turn or s_div:
                tp = 0.0
            else:
                d = w_sub - w_main
                tp = min(1.0, math.exp(d)) if d < 0 else 1.0
            if u_main < tp:
                zp, gp, prop_pe = szp, sgp, sprop_pe
            turning = True if s_turn else _is_turning(im, rl, rr, rsum_main)
            w_main = _logaddexp(w_main, w_sub)
            diverging = s_div
            sum_acc += s_acc
            n_prop += n_sub
            depth += 1
        accept = sum_acc / max(n_prop, 1)
        z, g, pe = zp.copy(), gp.copy(), prop_pe
        if it < num_warmup:
            if adapt_step:
                da_t += 1
                gg = target_accept - accept
                da_gavg = (1 - 1 / (da_t + 10)) * da_gavg + gg / (da_t + 10)
                da_x = da_prox - math.sqrt(da_t) / 0.05 * da_gavg
                wt = da_t ** (-0.75)
                da_xavg = (1 - wt) * da_xavg + wt * da_x
                step = math.exp(da_xavg) if it == num_warmup - 1 else math.exp(da_x)
                step = max(step, np.finfo(np.float32).tiny)
            middle = 0 < window < n_windows - 1
            if adapt_mass and middle:
                wf_n += 1
                dpre = z - wmean
                wmean = wmean + dpre / wf_n
                wm2 = wm2 + dpre * (z - wmean)
            at_end = it == win_end[window]
            if at_end:
                window += 1
            if at_end and middle:
                if adapt_mass:
                    cov = wm2 / (wf_n - 1)
                    if regularize:
                        cov = (wf_n / (wf_n + 5.0)) * cov + 1e-3 * (5.0 / (wf_n + 5.0))
                    im = cov
                    wf_n, wmean, wm2 = 0, np.zeros(D), np.zeros(D)
                if adapt_step:
                    da_prox = math.log(10 * step)
                    da_x = da_xavg = da_gavg = 0.0
                    da_t = 0
        if record is not None:
            record['z'].append(z.copy())
            record['accept'].append(accept)
            record['n_steps'].append(n_prop)
            record['diverging'].append(diverging)
            record['pe'].append(pe)
            record['step'].append(step)
    if record is not None:
        record['inv_mass'] = im.copy()


def train_momentum_layout(G, N, Dl, C, chain, sn_offset=0):
    """Momentum stream of the sharded training sampler (bayesn_train_nuts.cuh): the globals of
    chain c draw blocks of four from stream 0x7fff0000 + c, SN s draws lane-wise from stream
    (sn_offset + s) * C + c.  Returns (nrm_fn, unit) for :func:`chain_nuts` on the joint vector
    [globals (G), latents of SN 0 (Dl), SN 1, ...]."""
    gkey = (0x7fff0000 + chain) & M32

    def nrm_fn(it, key):
        out = np.empty(G + N * Dl)
        for b in range((G + 3) // 4):
            n4 = philox_normals((it, 0x40000000, b, gkey), key)
            for k in range(4):
                if 4 * b + k < G:
                    out[4 * b + k] = n4[k]
        for s in range(N):
            gunit = ((sn_offset + s) * C + chain) & M32
            cache = {}
            for e in range(Dl):
                ln, slot = e % 32, e // 32
                if ln not in cache:
                    cache[ln] = philox_normals((it, 0x40000000, ln, gunit), key)
                out[G + s * Dl + e] = cache[ln][slot & 3]
        return out
    return nrm_fn, gkey


def run_chains(logp_grad_batch, q_init, num_warmup, num_samples, unit_ids=None, **kw):
    """Drive B chains in lock-step.  ``logp_grad_batch(q (B, D), rows)`` must return
    (logp (B,), grad (B, D)) for the chains listed in ``rows``.  Returns (list of record dicts,
    number of batched evaluations ("ticks"), number of single-chain evaluations)."""
    B, D = q_init.shape
    if unit_ids is None:
        unit_ids = list(range(B))
    records = [dict(z=[], accept=[], n_steps=[], diverging=[], pe=[], step=[]) for _ in range(B)]
    gens = [chain_nuts(q_init[b], unit_ids[b], num_warmup, num_samples, record=records[b], **kw) for b in range(B)]
    pending = {}
    for b, gen in enumerate(gens):
        pending[b] = next(gen)
    ticks = evals = 0
    while pending:
        rows = sorted(pending.keys())
        q = np.stack([pending[b] for b in rows])
        lp, gr = logp_grad_batch(q, rows)
        ticks += 1
        evals += len(rows)
        nxt = {}
        for j, b in enumerate(rows):
            try:
                nxt[b] = gens[b].send((float(lp[j]), gr[j]))
            except StopIteration:
                pass
        pending = nxt
    return records, ticks, evals


def init_to_median_device_compat(muhat, n_chains, D, n_eps, tauA, sdD, seed=0):
    """CPU replica of ``init_median_kernel`` (same Philox streams): (N, C, D) float64."""
    N = len(muhat)
    key = ((seed & M32) ^ 0x5bd1e995, (((seed >> 32) & M32) + 0x1234567) & M32)
    out = np.empty((N, n_chains, D))
    for sn in range(N):
        for c in range(n_chains):
            unit = sn * n_chains + c
            for i in range(D):
                kind = 0 if i < n_eps else i - n_eps + 1
                v = []
                for b in range(4):
                    ra = philox4x32((b, i, 0x80000000, unit), key)
                    u = [u01(x) for x in ra]
                    n0 = math.sqrt(-2.0 * math.log(u[0] + 5.9604644775390625e-8))
                    n1 = math.sqrt(-2.0 * math.log(u[2] + 5.9604644775390625e-8))
                    nr = [n0 * math.cos(2 * math.pi * u[1]), n0 * math.sin(2 * math.pi * u[1]),
                          n1 * math.cos(2 * math.pi * u[3]), n1 * math.sin(2 * math.pi * u[3])]
                    for k in range(4):
                        if 4 * b + k < 15:
                            v.append(u[k] if kind in (2, 3, 5) else nr[k])
                med = sorted(v)[7]
                if kind in (0, 1):
                    o = med
                elif kind == 2:
                    o = math.log(-tauA * math.log(max(1.0 - med, 1e-7)))
                elif kind == 4:
                    o = muhat[sn] + sdD * med
                else:
                    u_ = min(max(med, 1e-6), 1 - 1e-6)
                    o = math.log(u_) - math.log1p(-u_)
                out[sn, c, i] = o
    return out
