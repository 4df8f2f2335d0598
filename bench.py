#!/usr/bin/env python
"""bench.py - headline benchmark of the BayeSN hot path on B200.

Metric (BASELINE.json): SNe fitted/s with NUTS at fixed draws (4 chains x (250 warm-up + 250
samples), max_tree_depth 10) on the W22 model, 1,000 simulated SNe (ugrizYJH, 5 epochs -> 40
observations each) per GPU - SURVEY.md section 8d config C3; one "step" = one complete fit of
the rank's 1,000 SNe.  Ranks shard SNe with no data-path collective (weak scaling).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

Prints ONE JSON line on rank 0.
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


def flops_per_eval(L, T, n_obs, B=300):
    """Algorithmic FLOPs of one log p + gradient evaluation, dense reference formulation
    (SURVEY.md section 8d): N_obs (4BL + 6LT + 14B) + 36B + 2B + 4 n_eps^2 + 4LT."""
    n_eps = (L - 2) * T
    return n_obs * (4 * B * L + 6 * L * T + 14 * B) + 36 * B + 2 * B + 4 * n_eps ** 2 + 4 * L * T


def load_consts():
    try:
        with open(CONSTS_PATH) as fh:
            return json.load(fh)
    except Exception:
        return {}


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


def build_workload(model, n_sne, seed):
    """simulate_light_curve recipe of SURVEY.md section 8d (C3) -> reference-layout arrays."""
    np.random.seed(seed)
    z = np.random.uniform(0.015, 0.08, n_sne)
    data, yerr, params = model.simulate_light_curve(EPOCHS, n_sne, BANDS, yerr=0.05, err_type='mag', z=z, mu='z',
                                                    ebv_mw=0, mag=False)
    return model.simulated_obs_array(data, yerr, params)


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout for the one JSON line: NCCL's own banner ("NCCL version ...", NCCL_DEBUG >= VERSION) goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    import __graft_entry__ as ge
    ge.build()
    from bayesn_b200 import SEDmodel
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

    kern_ms = []

    def step(i, time_kernel=False):
        ctx.init_to_median(N_CHAINS, seed=args.seed + i, out=q0)
        if time_kernel:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ctx.nuts(q0, num_warmup=N_WARMUP, num_samples=N_SAMPLES, max_tree_depth=MAX_DEPTH, seed=args.seed + i, out=bufs)
        if time_kernel:
            e1.record()
            kern_ms.append((e0, e1))

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_leapfrogs = 0.0
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i, time_kernel=True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches() - launches0
    clk = clocks.stop() if rank == 0 else None
    info = bufs['chain_info']
    leap_last = float(info[..., 2].sum().item())           # leapfrogs of the last timed step
    div_last = float(info[..., 3].sum().item())
    nuts_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ms]))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_sne * args.steps / (ms_max * 1e-3)

    # ---- end-to-end: host buffers in the C-ABI layout -> H2D -> init + NUTS -> D2H samples -------
    host_in = {k: packed[k].cpu().pin_memory() for k in ('obs4', 'n_valid', 'weights', 'support', 'muhat', 'lconst')}
    host_out = torch.empty(bufs['samples'].shape, dtype=torch.float32).pin_memory()
    host_stats = torch.empty(bufs['stats'].shape, dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    d2h = host_out.numel() * 4 + host_stats.numel() * 4

    def step_e2e(i):
        for k, v in host_in.items():
            packed[k].copy_(v, non_blocking=True)
        step(i)
        host_out.copy_(bufs['samples'], non_blocking=True)
        host_stats.copy_(bufs['stats'], non_blocking=True)
        torch.cuda.synchronize()

    step_e2e(0)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e(100 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_sne * e2e_steps / float(t.item())

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
        achieved = F * leap_last / (nuts_ms * 1e-3) / 1e12
        roofline = {'bound': 'fp32_fma', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak, 'traffic': None, 'kernel': 'nuts_kernel',
                    'traffic_note': 'ncu cannot replay this multi-second persistent launch (it stalls under '
                                    'instrumentation); a short launch of the same kernel (1184 SNe, 20+10 transitions, '
                                    'profiles/r01k_ncu_full_raw.csv) moved 5.4 MB read + 2.9 MB written over 4.84 M '
                                    'evaluations = 1.7 B/evaluation against 11.8 KB algorithmic: tiles, tables and the '
                                    'sampler workspace stay in L2/shared memory',
                    'peak_source': 'measured FFMA microbenchmark (bayesn_ffma_peak); MEASURED_PEAKS.json has no '
                                   'FP32 SIMT figure',
                    'algorithmic_flops_per_eval': F, 'evals_per_launch': leap_last, 'kernel_ms': nuts_ms}
        # ---- logp+grad microbenchmark (config C2) as an extra --------------------------------------
        if not args.no_micro:
            extra['logp_grad'] = micro_logp(model, ctx, args, dev, F, peak)
        cpu = cpu_baseline(obs, weights, band_inds, leap_last / n_sne, seconds=args.cpu_seconds)
        consts = {'evals_per_sn_fit': leap_last / n_sne, 'workload': 'C3 W22 1000 SNe x 40 obs, 4x(250+250)'}
        try:
            os.makedirs(os.path.dirname(CONSTS_PATH), exist_ok=True)
            if args.write_consts:
                with open(CONSTS_PATH, 'w') as fh:
                    json.dump(consts, fh)
        except Exception:
            pass
        line = {
            'metric': 'SNe fitted/s (NUTS, 4 chains x (250 warmup + 250 draws))', 'value': value, 'unit': 'SNe/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'C3: {MODEL} per-object fit of {n_sne} simulated SNe per GPU (ugrizYJH, 5 epochs,'
                                   f' {n_obs} obs), {N_CHAINS} chains x ({N_WARMUP}+{N_SAMPLES}), max_tree_depth '
                                   f'{MAX_DEPTH}', 'sharding': f'SNe sharded over {world} GPU(s), no collective',
                       'l2': 'weight tiles + sampler workspace of 1000 SNe (~45 MB) are L2-resident by design; '
                             'no flush between steps (each step is a ~1 s persistent kernel)'},
            'e2e': {'value': e2e_value, 'unit': 'SNe/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
            'leapfrogs_per_sn_fit': leap_last / n_sne, 'divergences_last_step': div_last,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def train_leg(args, dev, rank, world):
    """Config C4: M20 population training, 500 simulated SNe x 84 obs (CSP/RC bands),
    4 chains, train_model_globalRV sampled by the device-resident sharded NUTS; SNe are sharded over the
    ranks (strong scaling) and the per-leapfrog sums are all-reduced with NCCL.  Shortened run
    (warm-up + draws stated in the output); the unit is one leapfrog step of all 4 chains."""
    import torch
    import torch.distributed as dist
    from bayesn_b200 import SEDmodel, train as btrain
    from bayesn_b200.distributed import shard_bounds
    n_total, C, W, S = args.train_n, 4, args.train_warmup, args.train_samples
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
                         sn_offset=lo)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    st = out['stats'].cpu().numpy()
    res = {'workload': f'C4: M20_model train_model_globalRV, {n_total} simulated SNe x {obs.shape[1]} obs sharded over {world} '
                       f'GPU(s), {C} chains x ({W}+{S}) (nominal 500+500; shortened), max_tree_depth 10, NUTS step_size 0.1',
           'seconds': dt, 'passes': out['passes'], 'us_per_leapfrog_pass': dt / out['passes'] * 1e6,
           'leapfrog_passes_per_s': out['passes'] / dt, 'sn_evals_per_s': out['passes'] * n_total * C / dt,
           # SN kernel + two reduction stages + control kernel per pass (CUDA-graph replays included)
           'gpu_launches': 4 * out['passes'] + 2,
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


def micro_logp(model, ctx_fit, args, dev, F, peak):
    """Config C2: fused log p + gradient over N simulated SNe x 40 obs, one evaluation each."""
    import torch
    n = args.micro_n
    obs, weights, band_inds = build_workload(model, n, 777)
    ctx = model.prepare(obs, weights, band_inds)
    q = ctx.init_to_median(1, seed=5)
    for _ in range(3):
        ctx.logp_grad(q)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ctx.logp_grad(q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    evals_s = n / (ms * 1e-3)
    ach = F * evals_s / 1e12
    return {'n_sne': n, 'n_chains': 1, 'ms_per_launch': ms, 'evals_per_s': evals_s,
            'roofline': {'bound': 'fp32_fma', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak},
            'note': f'inputs: {n} compact weight tiles ({ctx._keep["data"]["weights"].numel() * 4 / 1e6:.0f} MB) + observations, '
                    f're-read from L2/HBM each launch; no L2 flush (a launch is 0.24 ms of compute-bound work)'}


def oracle_evals_per_s(obs, weights, band_inds_names, rows, seconds, threads=None):
    """Time the float64 oracle's batched log p + gradient (torch autograd) on the host cores."""
    import torch
    from oracle.ref_model import RefSED
    # all host cores this process may use (torchrun pins OMP_NUM_THREADS=1 by default)
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    ref = RefSED(MODEL, bands=BANDS)
    # oracle band order = YAML order; remap the product's compact band indices
    names = band_inds_names
    remap = np.array([ref.band_dict[n] for n in names])
    obs = np.array(obs[:, :, :rows])
    obs[4] = remap[obs[4].astype(int)]
    w = np.zeros((rows, 300, len(ref.band_dict)))
    w[:, :, remap] = weights[:rows]
    n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
    rng = np.random.RandomState(0)
    q = np.zeros((rows, n_eps + 4))
    q[:, 0] = rng.randn(rows)
    q[:, 1] = np.log(0.2)
    q[:, 3] = obs[7, 0, :]
    q[:, 4:] = rng.randn(rows, n_eps)
    obs_t, w_t = torch.as_tensor(obs), torch.as_tensor(w)
    ref.fit_logp_grad(q, obs_t, w_t)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        ref.fit_logp_grad(q, obs_t, w_t)
        n += rows
    dt = time.perf_counter() - t0
    return n / dt, torch.get_num_threads()


def cpu_baseline(obs, weights, band_inds, evals_per_fit, seconds=12.0, rows=256):
    from bayesn_b200 import datafiles
    fd = datafiles.load_yaml(os.path.join(datafiles.FILTER_DIR, 'filters.yaml'))
    all_names = ['NULL_BAND'] + list(fd['filters'].keys())
    names = [all_names[i] for i in band_inds]
    rows = min(rows, obs.shape[2])
    evals_s, threads = oracle_evals_per_s(obs, weights, names, rows, seconds)
    return {'value': evals_s / evals_per_fit, 'unit': 'SNe/s', 'cores': threads, 'kind': 'port',
            'evals_per_s': evals_s,
            'sample': f'float64 torch restatement of the reference (numpyro/JAX not installable): {seconds:.0f} s of '
                      f'batched log p + autograd gradient over {rows} SNe x 40 obs, converted with '
                      f'{evals_per_fit:.0f} leapfrog evaluations per SN fit (measured on the GPU run); NUTS '
                      f'bookkeeping not charged to the CPU'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    consts = load_consts()
    evals_per_fit = float(consts.get('evals_per_sn_fit', 40000.0))
    # inputs: the same simulate_light_curve recipe, produced by the oracle itself on the CPU
    from oracle.ref_model import RefSED
    import torch
    torch.set_num_threads(len(os.sched_getaffinity(0)))      # torchrun pins OMP_NUM_THREADS=1 by default
    ref = RefSED(MODEL, bands=BANDS)
    rows = 256
    np.random.seed(20241019)
    z = np.random.uniform(0.015, 0.08, rows)
    data, yerr, pd, (tt, bidx, bw) = ref.simulate_light_curve(EPOCHS, rows, BANDS, yerr=0.05, z=z, mu='z', mag=False)
    obs = np.zeros((10, data.shape[0], rows))
    obs[0], obs[1], obs[2], obs[4] = tt, data, yerr, bidx
    obs[5], obs[6], obs[7], obs[9] = z[None], 1e-4, ref.distmod(z)[None], 1
    n_eps = (len(ref.l_knots) - 2) * len(ref.tau_knots)
    rng = np.random.RandomState(0)
    q = np.zeros((rows, n_eps + 4))
    q[:, 0] = rng.randn(rows)
    q[:, 1] = np.log(0.2)
    q[:, 3] = obs[7, 0, :]
    q[:, 4:] = rng.randn(rows, n_eps)
    obs_t, w_t = torch.as_tensor(obs), torch.as_tensor(bw)
    evals_step = 20 * rows
    def step():
        for _ in range(20):
            ref.fit_logp_grad(q, obs_t, w_t)
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    evals_s = evals_step * args.steps / dt
    value = evals_s / evals_per_fit
    threads = torch.get_num_threads()
    print(json.dumps({
        'impl': 'reference', 'metric': 'SNe fitted/s (NUTS, 4 chains x (250 warmup + 250 draws))', 'value': value,
        'unit': 'SNe/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', 1)), 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'C3: {MODEL} per-object fit, {N_CHAINS} chains x ({N_WARMUP}+{N_SAMPLES}); each step '
                               f'= {evals_step} log p + gradient evaluations of the float64 CPU restatement'},
        'cpu_baseline': {'value': value, 'unit': 'SNe/s', 'cores': threads, 'kind': 'port', 'evals_per_s': evals_s,
                         'sample': f'{args.steps} steps x {evals_step} evaluations over {rows} SNe x 40 obs; converted '
                                   f'with {evals_per_fit:.0f} leapfrogs per SN fit (profiles/workload_constants.json)'},
        'e2e': {'value': value, 'unit': 'SNe/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n-sne', type=int, default=1000)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--micro-n', type=int, default=16384)
    ap.add_argument('--no-micro', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--write-consts', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-train-cpu', action='store_true')
    ap.add_argument('--train-n', type=int, default=500)
    ap.add_argument('--train-warmup', type=int, default=40)
    ap.add_argument('--train-samples', type=int, default=20)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
