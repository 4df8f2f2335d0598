"""Multi-GPU partitioning of the per-object fit: supernovae are independent posteriors
(reference ``run`` :1809-1849 is literally a vmap over SNe), so ranks take contiguous blocks of
SNe, run with no data-path collective, and only the per-SN summaries are gathered at the end.
One process per GPU (torchrun); NCCL on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of rank ``rank``; block sizes differ by at most one."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_obs(obs, weights, rank, world):
    """Slice the reference-layout arrays (``obs`` (10, N_obs, N), ``weights`` (N, 300, n_b))."""
    lo, hi = shard_bounds(obs.shape[-1], rank, world)
    return obs[:, :, lo:hi], weights[lo:hi], (lo, hi)


def gather_rows(local, n_total, group=None):
    """All-gather per-SN rows (first axis = this rank's SNe, ragged across ranks) into the full
    (n_total, ...) array on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.as_tensor(local)
    tail = tuple(t.shape[1:])
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tail, dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    full = torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)
    return full if isinstance(local, torch.Tensor) else full.cpu().numpy()


def posterior_summary(samples):
    """samples (N, C, S, D) torch tensor (device) -> (N, D, 2) mean / std over chains x draws,
    reduced on the device so that only 2 D floats per SN ever leave the GPU."""
    flat = samples.reshape(samples.shape[0], -1, samples.shape[-1])
    return np.stack([flat.mean(dim=1).cpu().numpy(), flat.std(dim=1).cpu().numpy()], axis=-1)
