"""Warm, back-to-back timing (CUDA events) of the stages of a training pass (fused reduction: SN kernel + first
reduction stage | control kernel incl. the second stage).  usage: python scripts/train_kernel_times.py [N_sne ...]
(N = 500: one GPU's C4; N = 63: a rank's shard of C4 on 8 GPUs, where the SN kernel runs in latency mode)"""
import sys, torch
sys.path.insert(0, '.')
from bayesn_b200 import SEDmodel, train as btrain
m = SEDmodel(load_model='M20_model')
for N in ([int(a) for a in sys.argv[1:]] or [500]):
  obs, weights, band_inds = btrain.simulate_training_set(m, N, seed=20241019)
  ctx = m.prepare_training(obs, None, band_inds)
  qg0, ql0 = btrain.initial_position(m, obs, num_chains=4, reference_model='M20_model', seed=1, device='cuda')
  out = ctx.train_nuts(qg0, ql0, num_warmup=20, num_samples=5, max_tree_depth=10, seed=2, max_passes=3000)   # leave the sampler mid-run
  lib, h = ctx.lib, ctx.h
  lib.bayesn_train_fuse_reduction(h, 1)
  red = ctx._keep['train_nuts'][1]
  st = torch.cuda.current_stream().cuda_stream
  def timeit(fn, n=200):
      for _ in range(20): fn()
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(n): fn()
      e1.record(); torch.cuda.synchronize()
      return e0.elapsed_time(e1) / n * 1e3
  t_sn = timeit(lambda: lib.bayesn_train_nuts_sn(h, red.data_ptr(), st))
  t_ctl = timeit(lambda: lib.bayesn_train_nuts_ctl(h, red.data_ptr(), st))
  def both():
      lib.bayesn_train_nuts_sn(h, red.data_ptr(), st); lib.bayesn_train_nuts_ctl(h, red.data_ptr(), st)
  t_pass = timeit(both)
  print(f'N={N} warps/unit={out["warps_per_unit"]}: SN kernel + first reduction stage: {t_sn:.1f} us; control kernel (+ second stage): {t_ctl:.1f} us; eager pass: {t_pass:.1f} us; graph-replay loop {out["loop_seconds"] / out["loop_passes"] * 1e6:.1f} us/pass')
