"""
One process per GPU: walkers are independent, so a batch of
parameter vectors is split contiguously over the ranks of a torch.distributed
group, every rank evaluates its rows on its own engine (constants replicated at
engine creation), and the per-walker lnL is gathered back -- B doubles, the only
cross-rank traffic. No collective touches the data path of the kernels.

Backend: ``nccl`` on GPU boxes (tensors on the rank's device), ``gloo`` for the
CPU test tier.
"""
import numpy as np


def shard_bounds(n_rows, world_size):
    """Contiguous split: rank r owns rows [bounds[r], bounds[r+1]); the first
    ``n_rows % world_size`` ranks get one extra row (same rule as the in-process
    multi-device split in csrc/engine.cu)."""
    base, extra = divmod(int(n_rows), int(world_size))
    sizes = [base + (1 if r < extra else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def sharded_lnlike(evaluate, thetas, group=None):
    """
    :param evaluate: callable (rows, D) -> (rows,) lnL for this rank's rows
        (e.g. ``model.log_likelihood_batch`` or ``model.log_posterior_batch``)
    :param thetas: the FULL (B, D) batch, identical on every rank
    :return: (B,) lnL on every rank
    """
    import torch
    import torch.distributed as dist
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(evaluate(thetas), dtype=np.float64)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(len(thetas), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    mine = np.asarray(evaluate(thetas[lo:hi]), dtype=np.float64) if hi > lo \
        else np.zeros(0)
    width = int(np.max(np.diff(bounds))) if len(thetas) else 0
    device = torch.device('cuda', torch.cuda.current_device()) \
        if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    send = torch.zeros(max(width, 1), dtype=torch.float64, device=device)
    send[:hi - lo] = torch.from_numpy(mine).to(device)
    gathered = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(gathered, send, group=group)
    out = np.empty(len(thetas), dtype=np.float64)
    for r in range(world):
        out[bounds[r]:bounds[r + 1]] = \
            gathered[r][:bounds[r + 1] - bounds[r]].cpu().numpy()
    return out


class ShardedPool(object):
    """``pool.map`` for emcee where every rank runs the same sampler (same seed)
    and each evaluates only its shard of every (half-)ensemble."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group

    def map(self, func, iterable):
        thetas = [np.asarray(p, dtype=np.float64) for p in iterable]
        if not thetas:
            return []
        lnpost = sharded_lnlike(self.model.log_posterior_batch, np.stack(thetas),
                                self.group)
        return [(v, {}) for v in lnpost.tolist()]

    def map_batch(self, func, block):
        """Array protocol of this package's sampler (cf. BatchPool.map_batch):
        (B, D) -> ((B,) lnpost, None)."""
        block = np.ascontiguousarray(block, dtype=np.float64)
        return sharded_lnlike(self.model.log_posterior_batch, block, self.group), None
