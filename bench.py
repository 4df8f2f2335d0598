#!/usr/bin/env python
"""bench.py - headline benchmark of the BayeSN hot path on B200.

Metric (BASELINE.json): SNe fitted/s with NUTS at fixed draws (4 chains x (250 warm-up + 250
samples), max_tree_depth 10) on the W22 model, simulated SNe (ugrizYJH, 5 epochs -> 40
observations each) - SURVEY.md section 8d config C3; one "step" = one complete fit of the rank's SNe.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

The JSON line (rank 0) carries
  value      device-timed SNe/s, 1000 SNe PER GPU resident in HBM (weak scaling: the fit mode has no collective);
  strong     C3 as BASELINE.json states it: 1000 SNe IN TOTAL sharded over the N ranks (latency mode picks the
             warps per chain from the shard size);
  e2e        the same fit through the public API - SEDmodel.fit_batch from reference-layout numpy arrays
             (obs (10, N_obs, N) float64, weights=None) to the reference-layout host sample dict;
  roofline   the persistent NUTS kernel against the measured FP32-FMA peak;
  logp_grad  the C2 microbenchmark sweep; latency: single fit() and C1; training: the C4 leg;
  cpu_baseline  the float64 oracle port timed on the host cores (sampler-in-the-loop and batched evaluator).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BANDS = ['u_LSST', 'g_PS1', 'r_PS1', 'i_PS1', 'z_PS1', 'Y_WFCAM', 'J_WFCAM', 'H_WFCAM']
EPOCHS = np.array([-8.0, 2.0, 12.0, 22.0, 32.0])
MODEL = 'W22_model'
N_CHAINS, N_WARMUP, N_SAMPLES, MAX_DEPTH = 4, 250, 250, 10
CONSTS_PATH = os.path.join(ROOT, 'profiles', 'workload_constants.json')
EXECUTED_PATH = os.path.join(ROOT, 'profiles', 'r02_executed.json')
W22_FIXTURE = os.path.join(ROOT, 'tests', 'golden', 'w22_oracle_posterior.npz')
METRIC = 'SNe fitted/s (NUTS, 4 chains x (250 warmup + 250 draws))'


def flops_per_eval(L, T, n_obs, B=300):
    """Algorithmic FLOPs of one log p + gradient evaluation, dense reference formulation
    (SURVEY.md section 8d): N_obs (4BL + 6LT + 14B) + 36B + 2B + 4 n_eps^2 + 4LT."""
    n_eps = (L - 2) * T
    return n_obs * (4 * B * L + 6 * L * T + 14 * B) + 36 * B + 2 * B + 4 * n_eps ** 2 + 4 * L * T


def evals_per_sn_fit():
    """Leapfrog evaluations of one C3 SN fit (4 chains x (250 + 250)), needed to express a CPU evaluation rate
    in SNe/s.  Preferred source: the float64 oracle sampler's own count on the committed W22 fixture
    (tests/golden/w22_oracle_posterior.npz, 16 full CPU fits); else the count measured on the GPU
    (profiles/workload_constants.json).  No silent default: without either the CPU arm cannot be converted."""
    if os.path.exists(W22_FIXTURE):
        d = np.load(W22_FIXTURE)
        return float(np.asarray(d['all_steps']).sum(axis=1).mean()), 'oracle sampler, tests/golden/w22_oracle_posterior.npz'
    if os.path.exists(CONSTS_PATH):
        with open(CONSTS_PATH) as fh:
            return float(json.load(fh)['evals_per_sn_fit']), 'GPU run, profiles/workload_constants.json'
    raise RuntimeError('neither tests/golden/w22_oracle_posterior.npz nor profiles/workload_constants.json exists: '
                       'cannot convert CPU evaluations/s to SNe/s')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for row in self.rows:
            parts = [p.strip() for p in row.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def build_workload(model, n_sne, seed, distinct=None):
    """simulate_light_curve recipe of SURVEY.md section 8d (C3) -> reference-layout arrays.  ``distinct``: simulate
    only that many SNe and repeat them (large microbenchmark sizes; every replica still has its own tile)."""
    np.random.seed(seed)
    n_sim = n_sne if distinct is None else min(n_sne, distinct)
    z = np.random.uniform(0.015, 0.08, n_sim)
    data, yerr, params = model.simulate_light_curve(EPOCHS, n_sim, BANDS, yerr=0.05, err_type='mag', z=z, mu='z',
                                                    ebv_mw=0, mag=False)
    obs, w, bi = model.simulated_obs_array(data, yerr, params)
    if n_sim < n_sne:
        reps = -(-n_sne // n_sim)
        obs = np.tile(obs, (1, 1, reps))[:, :, :n_sne]
        w = None
    return obs, w, bi


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    # keep stdout for the ONE JSON line: everything libraries print on file descriptor 1 (NCCL's version banner, ...)
    # goes to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    import __graft_entry__ as ge
    ge.build()
    from bayesn_b200 import SEDmodel
    from bayesn_b200.distributed import shard_bounds
    dev = f'cuda:{local}'
    model = SEDmodel(load_model=MODEL, device=dev)
    n_sne = args.n_sne
    obs, weights, band_inds = build_workload(model, n_sne, 20241019 + rank)
    ctx = model.prepare(obs, weights, band_inds)
    packed = ctx._keep['data']
    D = ctx.D
    n_obs = int(packed['n_valid'].max().item())
    bufs = {}
    q0 = torch.empty((n_sne, N_CHAINS, D), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    kern_ms = []
    leap = torch.zeros(2, dtype=torch.float64, device=dev)      # leapfrogs / divergences summed over the timed steps

    def step(i, timed=False):
        ctx.init_to_median(N_CHAINS, seed=args.seed + i, out=q0)
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ctx.nuts(q0, num_warmup=N_WARMUP, num_samples=N_SAMPLES, max_tree_depth=MAX_DEPTH, seed=args.seed + i, out=bufs)
        if timed:
            e1.record()
            kern_ms.append((e0, e1))
            leap.add_(bufs['chain_info'][..., 2:4].sum(dim=(0, 1)).double())

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i, timed=True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches() - launches0
    clk = clocks.stop() if rank == 0 else None
    leap_total, div_total = (float(x) for x in leap.cpu())
    nuts_ms_total = float(np.sum([a.elapsed_time(b) for a, b in kern_ms]))
    ms_max = max_over_ranks(ms)
    value = world * n_sne * args.steps / (ms_max * 1e-3)

    # ---- end-to-end through the public API: reference-layout numpy -> SEDmodel.fit_batch -> host sample dict ----
    # (host packing, H2D, device tile build, init, NUTS, constrain + transposition on the device, D2H of every site)
    def step_e2e(i):
        samples, extras = model.fit_batch(obs, None, band_inds, num_chains=N_CHAINS, num_warmup=N_WARMUP,
                                          num_samples=N_SAMPLES, max_tree_depth=MAX_DEPTH, seed=args.seed + i)
        return samples, extras

    _, ex = step_e2e(0)
    del _
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s_, ex = step_e2e(args.warmup + i)          # the seeds of the timed loop
        del s_
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n_sne * args.steps / e2e_s
    h2d, d2h = ex['h2d_bytes'], ex['d2h_bytes']

    # ---- strong scaling: C3 as stated, 1000 SNe IN TOTAL sharded over the ranks ---------------------------------
    strong = None
    if world > 1 or args.strong_at_one:
        model_s = SEDmodel(load_model=MODEL, device=dev)
        obs_t, _, bi_t = build_workload(model_s, args.n_total, 20241019)          # the same data set on every rank
        lo, hi = shard_bounds(args.n_total, rank, world)
        ctx_s = model_s.prepare(obs_t[:, :, lo:hi], None, bi_t)
        K = ctx_s.auto_warps_per_chain((hi - lo) * N_CHAINS) if args.warps_per_chain == 'auto' else int(args.warps_per_chain)
        qs = ctx_s.init_to_median(N_CHAINS, seed=args.seed, sn_offset=lo)
        sb = {}
        for i in range(max(1, min(args.warmup, 2))):
            ctx_s.nuts(qs, N_WARMUP, N_SAMPLES, MAX_DEPTH, seed=args.seed, sn_offset=lo, warps_per_chain=K, out=sb)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            ctx_s.init_to_median(N_CHAINS, seed=args.seed + i, out=qs, sn_offset=lo)
            ctx_s.nuts(qs, N_WARMUP, N_SAMPLES, MAX_DEPTH, seed=args.seed + i, sn_offset=lo, warps_per_chain=K, out=sb)
        s1.record()
        barrier()
        s_ms = max_over_ranks(s0.elapsed_time(s1))
        longest = max_over_ranks(float(sb['chain_info'][..., 2].max().item()))
        strong = {'workload': f'C3 as stated: {args.n_total} SNe in total sharded over {world} GPU(s), {hi - lo} on rank 0',
                  'scaling': 'strong', 'value': args.n_total * args.steps / (s_ms * 1e-3), 'unit': 'SNe/s',
                  'ms_per_step': s_ms / args.steps, 'warps_per_chain': K,
                  'longest_chain_leapfrogs': longest,
                  'bound': 'a step cannot end before its longest chain: longest_chain_leapfrogs x the latency of one '
                           'leapfrog of one chain (the mean chain is ~0.65 of the longest)'}
        del ctx_s, sb

    extra = {}
    if not args.no_train:
        tr = train_leg(args, dev, rank, world)       # every rank takes part (SNe sharded, all-reduce per leapfrog)
        if rank == 0:
            extra['training'] = tr
    if rank == 0:
        # ---- roofline of the dominant kernel (the persistent NUTS kernel) -------------------------
        L, T = len(model.l_knots), len(model.tau_knots)
        F = flops_per_eval(L, T, n_obs)
        peak = ctx.ffma_peak_tflops()
        achieved = F * leap_total / (nuts_ms_total * 1e-3) / 1e12
        executed = None
        if os.path.exists(EXECUTED_PATH):
            with open(EXECUTED_PATH) as fh:
                executed = json.load(fh)
        roofline = {'bound': 'fp32_fma', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak,
                    'traffic': ((executed or {}).get('dram_bytes_per_eval') or 0) * leap_total / args.steps or None,
                    'kernel': 'nuts_kernel',
                    'executed_frac': (executed or {}).get('fma_pipe_frac'),
                    'executed_note': 'frac counts the DENSE algorithmic flops of SURVEY 8d; the kernel skips the exact-zero '
                                     'part of every band (C3 bands cover 22-62 of 300 bins), so the FMA pipe itself is busy '
                                     'executed_frac of the time (ncu sm__inst_executed_pipe_fma, ' +
                                     ((executed or {}).get('source', 'profiles/README.md')) + ')',
                    'traffic_note': 'bytes per launch = (dram__bytes_read + dram__bytes_write) per evaluation of a short launch of '
                                    'the same kernel under ncu --set full (2.5 B; ncu cannot replay the multi-second launch) x the '
                                    'evaluations of this launch; the algorithmic bytes are 11.8 KB per evaluation: tiles, tables '
                                    'and the sampler workspace stay in L2 / shared memory',
                    'executed': {k: (executed or {}).get(k) for k in ('fma_pipe_frac', 'fma_inst_frac', 'issue_active_frac',
                                                                      'lsu_pipe_frac', 'warp_instructions_per_eval')},
                    'peak_source': 'measured FFMA microbenchmark (bayesn_ffma_peak); MEASURED_PEAKS.json has no '
                                   'FP32 SIMT figure',
                    'algorithmic_flops_per_eval': F, 'evals_per_launch': leap_total / args.steps,
                    'kernel_ms': nuts_ms_total / args.steps}
        if not args.no_micro:
            extra['logp_grad'] = micro_logp(model, args, dev, F, peak)
        if not args.no_latency:
            extra['latency'] = latency_leg(dev)
        per_fit = leap_total / (args.steps * n_sne)
        cpu = cpu_baseline(obs, weights, band_inds, per_fit, seconds=args.cpu_seconds)
        if args.write_consts:
            with open(CONSTS_PATH, 'w') as fh:
                json.dump({'evals_per_sn_fit': per_fit, 'workload': 'C3 W22 1000 SNe x 40 obs, 4x(250+250)'}, fh)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'SNe/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'C3: {MODEL} per-object fit of {n_sne} simulated SNe per GPU (ugrizYJH, 5 epochs,'
                                   f' {n_obs} obs), {N_CHAINS} chains x ({N_WARMUP}+{N_SAMPLES}), max_tree_depth '
                                   f'{MAX_DEPTH}', 'sharding': f'SNe sharded over {world} GPU(s), no collective',
                       'l2': 'weight tiles + sampler workspace of 1000 SNe (~45 MB) are L2-resident by design; '
                             'no flush between steps (each step is a ~3 s persistent kernel)',
                       'parity': 'log p rel 1e-5; gradient 1e-4 NORM-WISE per SN (max|dg| / max|g|); per-component worst '
                                 'case over components >= 1 % of the largest: tests/test_gpu_round2.py, profiles/parity_r02.jsonl'},
            'e2e': {'value': e2e_value, 'unit': 'SNe/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'SEDmodel.fit_batch(obs (10, N_obs, N) float64, weights=None) -> reference-layout host samples',
                    'steps': args.steps},
            'gpu_launches': launches, 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
            'leapfrogs_per_sn_fit': per_fit, 'divergences_per_step': div_total / args.steps,
        }
        if strong is not None:
            line['strong'] = strong
        line.update(extra)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latency_leg(dev):
    """Few chains on a whole GPU: the reference's single-object fit() (SN 2016W, example_fits.ipynb) and config C1
    (10 T21 SNe); wall time of the public call, one warp per chain vs latency mode."""
    import torch
    from bayesn_b200 import SEDmodel
    res = {}
    m = SEDmodel(load_model='T21_model', device=dev)
    fm = {'g': 'g_PS1', 'r': 'r_PS1', 'i': 'i_PS1', 'z': 'z_PS1'}
    for K in (1, 'auto'):
        m.fit_from_file(m.example_lc, filt_map=fm, print_summary=False, warps_per_chain=K)       # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.fit_from_file(m.example_lc, filt_map=fm, print_summary=False, warps_per_chain=K)
        torch.cuda.synchronize()
        res[f'fit_2016W_seconds_K{K}'] = time.perf_counter() - t0
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'c1_oracle_posterior.npz'))
    bi = np.array([m.band_dict[str(n)] for n in d['band_names']])
    ctx = m.prepare(d['obs'], None, bi)
    q0 = ctx.init_to_median(4, seed=0)
    for K in (1, 2, 4, 8):
        b = {}
        ctx.nuts(q0, 250, 250, 10, seed=0, warps_per_chain=K, out=b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.nuts(q0, 250, 250, 10, seed=0, warps_per_chain=K, out=b)
        e1.record()
        torch.cuda.synchronize()
        lf = b['chain_info'][..., 2]
        res[f'C1_K{K}'] = {'seconds': e0.elapsed_time(e1) * 1e-3, 'leapfrogs': float(lf.sum()),
                           'us_per_leapfrog_of_longest_chain': e0.elapsed_time(e1) * 1e3 / float(lf.max())}
    res['note'] = ('C1: T21, 10 SNe x 48 obs, 40 chains; oracle sampler on 8 host cores: '
                   f'{float(d["seconds"]):.0f} s for the same ten fits (tests/golden/c1_oracle_posterior.npz)')
    return res


def train_leg(args, dev, rank, world):
    """Config C4: M20 population training, 500 simulated SNe x 84 obs (CSP/RC bands),
    4 chains, train_model_globalRV sampled by the device-resident sharded NUTS; SNe are sharded over the
    ranks (strong scaling) and the per-leapfrog sums are all-reduced with NCCL.  Shortened run unless
    --train-nominal (warm-up + draws stated in the output); the unit is one leapfrog step of all 4 chains."""
    import torch
    import torch.distributed as dist
    from bayesn_b200 import SEDmodel, train as btrain
    from bayesn_b200.distributed import shard_bounds
    n_total, C = args.train_n, 4
    W, S = (500, 500) if args.train_nominal else (args.train_warmup, args.train_samples)
    m = SEDmodel(load_model='M20_model', device=dev)
    obs, weights, band_inds = btrain.simulate_training_set(m, n_total, seed=20241019)
    lo, hi = shard_bounds(n_total, rank, world)
    ctx = m.prepare_training(obs[:, :, lo:hi], weights[lo:hi], band_inds)
    qg0, ql0 = btrain.initial_position(m, obs, C, 'M20_model', seed=1, device=dev)
    group = dist.group.WORLD if world > 1 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ctx.train_nuts(qg0, ql0[lo:hi].contiguous(), num_warmup=W, num_samples=S, max_tree_depth=10, seed=2, group=group,
                         sn_offset=lo, exchange=args.train_exchange)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    st = out['stats'].cpu().numpy()
    res = {'workload': f'C4: M20_model train_model_globalRV, {n_total} simulated SNe x {obs.shape[1]} obs sharded over {world} '
                       f'GPU(s), {C} chains x ({W}+{S}) (nominal 500+500{"" if args.train_nominal else "; shortened"}), '
                       f'max_tree_depth 10, NUTS step_size 0.1',
           # steady state of the sampler: CUDA-graph replays of whole passes, everything of NUTS included; the one-off
           # set-up (graph capture, symmetric-memory rendezvous of the peer exchange) is reported beside it
           'seconds': dt, 'passes': out['passes'], 'warps_per_unit': out.get('warps_per_unit', 1),
           'us_per_leapfrog_pass': out['loop_seconds'] / max(out['loop_passes'], 1) * 1e6,
           'us_per_leapfrog_pass_incl_setup': dt / out['passes'] * 1e6,
           'setup_seconds': dt - out['loop_seconds'],
           'leapfrog_passes_per_s': out['loop_passes'] / out['loop_seconds'],
           'sn_evals_per_s': out['loop_passes'] * n_total * C / out['loop_seconds'],
           'gpu_launches': out.get('launches_per_pass', 4) * out['passes'] + 2,
           'exchange': out.get('exchange', 'nccl' if world > 1 else 'none'),
           'allreduce_doubles_per_pass': (C * ctx.train_dims()[2]) if world > 1 else 0,
           'mean_accept_sampling': float(st[:, W:, 0].mean()), 'divergences': float(out['info'][:, 5].sum()),
           'finished': out['finished']}
    if rank == 0 and not args.no_train_cpu:
        from oracle.ref_model import RefSED
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        bands = ['B_CSP', 'V_CSP', 'r_CSP', 'i_CSP', 'Y_RC', 'J_RC1', 'H_RC']
        ref = RefSED('M20_model', bands=bands)
        from bayesn_b200 import datafiles
        fd = datafiles.load_yaml(os.path.join(datafiles.FILTER_DIR, 'filters.yaml'))
        all_names = ['NULL_BAND'] + list(fd['filters'].keys())
        remap = np.array([ref.band_dict[all_names[i]] for i in band_inds])
        o = np.array(obs)
        o[4] = remap[o[4].astype(int)]
        w = np.zeros((n_total, 300, len(ref.band_dict)))
        w[:, :, remap] = weights
        qg = qg0[0].cpu().numpy().astype(np.float64)
        ql = ql0[:, 0].cpu().numpy().astype(np.float64)
        ref.train_logp_grad(qg, ql, o, w)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 6.0:
            ref.train_logp_grad(qg, ql, o, w)
            n += 1
        cdt = time.perf_counter() - t0
        res['cpu_baseline'] = {'value': n / cdt / 1.0, 'unit': 'joint log p + gradient evaluations/s (one chain)',
                               'cores': torch.get_num_threads(), 'kind': 'port',
                               'sample': f'{n} evaluations of the float64 oracle train_logp_grad over {n_total} SNe',
                               'sn_evals_per_s': n * n_total / cdt}
        res['speedup_vs_cpu_port_sn_evals'] = res['sn_evals_per_s'] / res['cpu_baseline']['sn_evals_per_s']
    return res


def micro_logp(model, args, dev, F, peak):
    """Config C2: fused log p + gradient over N simulated SNe x 40 obs, one evaluation each, N = 1k ... 1M
    (BASELINE.json configs[1]).  Up to 16 384 distinct simulated SNe, repeated beyond that (every replica has its
    own weight tile and observations in HBM; at 1M SNe the inputs are 2.5 GB, far beyond L2)."""
    import torch
    rows = []
    for n in args.micro_sizes:
        obs, weights, band_inds = build_workload(model, n, 777, distinct=16384)
        ctx = model.prepare(obs, None, band_inds)
        q = ctx.init_to_median(1, seed=5)
        for _ in range(3):
            ctx.logp_grad(q)
        torch.cuda.synchronize()
        reps = max(3, min(20, int(4e6 // n)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.logp_grad(q)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        evals_s = n / (ms * 1e-3)
        ach = F * evals_s / 1e12
        in_bytes = ctx._keep['data']['weights'].numel() * 4 + ctx._keep['data']['obs4'].numel() * 4
        rows.append({'n_sne': n, 'n_chains': 1, 'ms_per_launch': ms, 'evals_per_s': evals_s, 'input_mb': in_bytes / 1e6,
                     'roofline': {'bound': 'fp32_fma', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak}})
        del ctx, q
        torch.cuda.empty_cache()
    best = max(rows, key=lambda r: r['evals_per_s'])
    return {'sweep': rows, 'n_sne': best['n_sne'], 'evals_per_s': best['evals_per_s'], 'roofline': best['roofline'],
            'note': 'inputs (compact weight tiles + observations) are re-read from L2 / HBM every launch; sizes beyond 16 384 '
                    'repeat the simulated SNe; no L2 flush needed from 131 072 SNe on (inputs > 126 MB L2)'}


def oracle_setup(obs, weights, band_inds, rows):
    """Oracle model + the product's arrays remapped to the oracle's band order."""
    import torch
    from oracle.ref_model import RefSED
    from bayesn_b200 import datafiles
    fd = datafiles.load_yaml(os.path.join(datafiles.FILTER_DIR, 'filters.yaml'))
    all_names = ['NULL_BAND'] + list(fd['filters'].keys())
    names = [all_names[i] for i in band_inds]
    ref = RefSED(MODEL, bands=BANDS)
    remap = np.array([ref.band_dict[n] for n in names])
    obs = np.array(obs[:, :, :rows])
    obs[4] = remap[obs[4].astype(int)]
    w = np.zeros((rows, 300, len(ref.band_dict)))
    w[:, :, remap] = weights[:rows]
    return ref, torch.as_tensor(obs), torch.as_tensor(w)


def oracle_evals_per_s(ref, obs_t, w_t, seconds):
    """Batched log p + gradient (torch autograd, float64) of the oracle on the host cores."""
    rows = obs_t.shape[-1]
    n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
    rng = np.random.RandomState(0)
    q = np.zeros((rows, n_eps + 4))
    q[:, 0] = rng.randn(rows)
    q[:, 1] = np.log(0.2)
    q[:, 3] = obs_t[7, 0, :].numpy()
    q[:, 4:] = rng.randn(rows, n_eps)
    ref.fit_logp_grad(q, obs_t, w_t)
    reps, n, t0 = [], 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        t1 = time.perf_counter()
        ref.fit_logp_grad(q, obs_t, w_t)
        reps.append(rows / (time.perf_counter() - t1))
        n += rows
    return n / (time.perf_counter() - t0), reps


def oracle_sampler_evals_per_s(ref, obs_t, w_t, n_sne, seconds, seed=0):
    """The oracle NUTS itself (oracle/ref_nuts.py: numpyro semantics, float64) on ``n_sne`` SNe x 4 chains driven in
    lock-step, for ``seconds`` of wall clock from the start of warm-up: chain-leapfrogs per second with all the
    sampler bookkeeping charged."""
    import torch
    from oracle import ref_nuts
    obs = obs_t.numpy()
    n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
    D = n_eps + 4
    qk = ref_nuts.init_to_median_device_compat(obs[7, 0, :n_sne], N_CHAINS, D, n_eps, ref.tauA,
                                               math.sqrt(25 + ref.sigma0 ** 2), seed=seed)
    q0 = np.concatenate([qk[..., n_eps:], qk[..., :n_eps]], axis=-1).reshape(n_sne * N_CHAINS, D)
    sn_of = np.repeat(np.arange(n_sne), N_CHAINS)
    gens = [ref_nuts.chain_nuts(q0[b], b, N_WARMUP, N_SAMPLES, max_depth=MAX_DEPTH, seed=seed,
                                record=dict(z=[], accept=[], n_steps=[], diverging=[], pe=[], step=[])) for b in range(len(q0))]
    pending = {b: next(g) for b, g in enumerate(gens)}
    evals, ticks, t0 = 0, 0, time.perf_counter()
    while pending and time.perf_counter() - t0 < seconds:
        rows = sorted(pending)
        q = torch.as_tensor(np.stack([pending[b] for b in rows])).clone().requires_grad_(True)
        idx = sn_of[rows]
        lp = ref.fit_logp(q, obs_t[:, :, idx], w_t[idx])
        g, = torch.autograd.grad(lp.sum(), q)
        lp, g = lp.detach().numpy(), g.numpy()
        evals += len(rows)
        ticks += 1
        nxt = {}
        for j, b in enumerate(rows):
            try:
                nxt[b] = gens[b].send((float(lp[j]), g[j]))
            except StopIteration:
                pass
        pending = nxt
    dt = time.perf_counter() - t0
    return evals / dt, ticks, dt


def cpu_baseline(obs, weights, band_inds, evals_per_fit, seconds=12.0, rows=256, sampler_sne=64):
    import torch
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)          # all host cores this process may use (torchrun pins OMP_NUM_THREADS=1)
    if weights is None:
        raise RuntimeError('cpu_baseline needs the host band weights of the sample')
    rows = min(rows, obs.shape[2])
    ref, obs_t, w_t = oracle_setup(obs, weights, band_inds, rows)
    evals_s, reps = oracle_evals_per_s(ref, obs_t, w_t, seconds * 0.4)
    samp_s, ticks, sdt = oracle_sampler_evals_per_s(ref, obs_t, w_t, min(sampler_sne, rows), seconds * 0.6)
    return {'value': samp_s / evals_per_fit, 'unit': 'SNe/s', 'cores': threads, 'kind': 'port',
            'sampler_leapfrogs_per_s': samp_s, 'evaluator_evals_per_s': evals_s,
            'evaluator_evals_per_s_spread': [float(np.min(reps)), float(np.max(reps))], 'evaluator_repeats': len(reps),
            'value_evaluator_only': evals_s / evals_per_fit,
            'sample': f'float64 torch restatement of the reference (numpyro/JAX not installable): {sdt:.0f} s ({ticks} lock-step '
                      f'ticks) of the oracle NUTS on {min(sampler_sne, rows)} SNe x {N_CHAINS} chains x 40 obs, sampler '
                      f'bookkeeping included, + {seconds * 0.4:.0f} s of batched log p + autograd gradient over {rows} SNe; '
                      f'converted with {evals_per_fit:.0f} leapfrogs per SN fit (this GPU run; the oracle sampler needs '
                      f'the same number: tests/golden/w22_oracle_posterior.npz)'}


def run_reference(args):
    """Reference arm: the float64 CPU restatement of the reference (kind 'port': numpyro/JAX cannot be installed
    here) on all host cores.  One step = a bounded sample of the C3 workload: ``--ref-ticks`` lock-step ticks of the
    oracle NUTS over 64 SNe x 4 chains (every tick is one batched log p + gradient evaluation of all live chains
    plus the sampler bookkeeping), converted to SNe/s with the leapfrogs one SN fit needs."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    per_fit, per_fit_src = evals_per_sn_fit()
    from oracle.ref_model import RefSED
    import torch
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)      # torchrun pins OMP_NUM_THREADS=1 by default
    ref = RefSED(MODEL, bands=BANDS)
    rows = 64
    np.random.seed(20241019)
    z = np.random.uniform(0.015, 0.08, rows)
    data, yerr, pd, (tt, bidx, bw) = ref.simulate_light_curve(EPOCHS, rows, BANDS, yerr=0.05, z=z, mu='z', mag=False)
    obs = np.zeros((10, data.shape[0], rows))
    obs[0], obs[1], obs[2], obs[4] = tt, data, yerr, bidx
    obs[5], obs[6], obs[7], obs[9] = z[None], 1e-4, ref.distmod(z)[None], 1
    obs_t, w_t = torch.as_tensor(obs), torch.as_tensor(bw)
    from oracle import ref_nuts
    n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
    D = n_eps + 4
    qk = ref_nuts.init_to_median_device_compat(obs[7, 0, :], N_CHAINS, D, n_eps, ref.tauA, math.sqrt(25 + ref.sigma0 ** 2), seed=0)
    q0 = np.concatenate([qk[..., n_eps:], qk[..., :n_eps]], axis=-1).reshape(rows * N_CHAINS, D)
    sn_of = np.repeat(np.arange(rows), N_CHAINS)
    gens = [ref_nuts.chain_nuts(q0[b], b, N_WARMUP, N_SAMPLES, max_depth=MAX_DEPTH, seed=0,
                                record=dict(z=[], accept=[], n_steps=[], diverging=[], pe=[], step=[])) for b in range(len(q0))]
    state = {'pending': {b: next(g) for b, g in enumerate(gens)}}

    def step():
        """args.ref_ticks lock-step ticks of the running fits; returns chain-leapfrogs done."""
        evals = 0
        for _ in range(args.ref_ticks):
            pending = state['pending']
            if not pending:
                break
            rws = sorted(pending)
            q = torch.as_tensor(np.stack([pending[b] for b in rws])).clone().requires_grad_(True)
            idx = sn_of[rws]
            lp = ref.fit_logp(q, obs_t[:, :, idx], w_t[idx])
            g, = torch.autograd.grad(lp.sum(), q)
            lp, g = lp.detach().numpy(), g.numpy()
            evals += len(rws)
            nxt = {}
            for j, b in enumerate(rws):
                try:
                    nxt[b] = gens[b].send((float(lp[j]), g[j]))
                except StopIteration:
                    pass
            state['pending'] = nxt
        return evals

    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        evals += step()
    dt = time.perf_counter() - t0
    evals_s = evals / dt
    value = evals_s / per_fit
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value,
        'unit': 'SNe/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', 1)), 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'C3: {MODEL} per-object fit, {N_CHAINS} chains x ({N_WARMUP}+{N_SAMPLES}); each step = '
                               f'{args.ref_ticks} lock-step ticks of the float64 oracle NUTS over {rows} SNe x {N_CHAINS} chains '
                               f'(bounded sample of the same fits, sampler bookkeeping included)'},
        'cpu_baseline': {'value': value, 'unit': 'SNe/s', 'cores': threads, 'kind': 'port', 'sampler_leapfrogs_per_s': evals_s,
                         'sample': f'{args.steps} steps x {args.ref_ticks} ticks = {evals} chain-leapfrogs of the oracle NUTS in {dt:.1f} s; '
                                   f'converted with {per_fit:.0f} leapfrogs per SN fit ({per_fit_src})'},
        'e2e': {'value': value, 'unit': 'SNe/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n-sne', type=int, default=1000, help='SNe per GPU of the headline (weak) measurement')
    ap.add_argument('--n-total', type=int, default=1000, help='SNe in total of the strong-scaling block')
    ap.add_argument('--warps-per-chain', default='auto')
    ap.add_argument('--strong-at-one', action='store_true', help='also run the strong block on a single GPU')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--micro-sizes', type=lambda s: [int(x) for x in s.split(',')], default=[1024, 16384, 131072, 1048576])
    ap.add_argument('--no-micro', action='store_true')
    ap.add_argument('--no-latency', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--ref-ticks', type=int, default=150)
    ap.add_argument('--write-consts', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-train-cpu', action='store_true')
    ap.add_argument('--train-n', type=int, default=500)
    ap.add_argument('--train-warmup', type=int, default=40)
    ap.add_argument('--train-samples', type=int, default=20)
    ap.add_argument('--train-nominal', action='store_true', help='run the nominal 4 x (500 + 500) C4 training (~90 s)')
    ap.add_argument('--train-exchange', default='auto', choices=['auto', 'nccl', 'peer'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
