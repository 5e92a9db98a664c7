"""
One process per GPU: walkers are independent, so a batch of
parameter vectors is split contiguously over the ranks of a torch.distributed
group, every rank evaluates its rows on its own engine (constants replicated at
engine creation), and the per-walker lnL is gathered back -- B doubles, the only
cross-rank traffic. No collective touches the data path of the kernels.

This is the pool.map of /root/reference/psfMC/fitting.py:55-58 (the reference
builds its sampler without a pool; emcee's hook for parallel evaluation is
``pool.map``) for a job launched with torchrun: every rank runs the same seeded
sampler and evaluates only its shard of every (half-)ensemble.

Backend: ``nccl`` on GPU boxes (all_gather_into_tensor on the ranks' devices),
``gloo`` for the CPU test tier.
"""
import numpy as np


def shard_bounds(n_rows, world_size):
    """Contiguous split: rank r owns rows [bounds[r], bounds[r+1]); the first
    ``n_rows % world_size`` ranks get one extra row (same rule as the in-process
    multi-device split in csrc/engine.cu)."""
    base, extra = divmod(int(n_rows), int(world_size))
    sizes = [base + (1 if r < extra else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


class ShardedEvaluator(object):
    """Sharded evaluation of (B, D) batches with reusable staging buffers.

    :param evaluate: callable (rows, D) -> (rows,) for this rank's rows
        (e.g. ``model.log_likelihood_batch`` or ``model.log_posterior_batch``)
    :param group: torch.distributed process group (default: the world)
    """

    def __init__(self, evaluate, group=None):
        self.evaluate = evaluate
        self.group = group
        self._width = 0
        self._send = self._recv = self._send_host = self._recv_host = None

    def _buffers(self, width, world, device):
        import torch
        if width > self._width or self._send is None or self._send.device != device:
            self._width = max(width, 2 * self._width)
            pin = device.type == 'cuda'
            self._send = torch.zeros(self._width, dtype=torch.float64, device=device)
            self._recv = torch.zeros(world * self._width, dtype=torch.float64, device=device)
            self._send_host = torch.zeros(self._width, dtype=torch.float64, pin_memory=pin)
            self._recv_host = torch.zeros(world * self._width, dtype=torch.float64,
                                          pin_memory=pin)
        return self._send, self._recv, self._send_host, self._recv_host

    def __call__(self, thetas):
        """:param thetas: the FULL (B, D) batch, identical on every rank
        :return: (B,) results on every rank"""
        import torch
        import torch.distributed as dist
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        if not (dist.is_available() and dist.is_initialized()):
            return np.asarray(self.evaluate(thetas), dtype=np.float64)
        group = self.group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n_rows = len(thetas)
        if world == 1 or n_rows == 0:
            return np.asarray(self.evaluate(thetas), dtype=np.float64)
        bounds = shard_bounds(n_rows, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        mine = np.asarray(self.evaluate(thetas[lo:hi]), dtype=np.float64) if hi > lo \
            else np.zeros(0)
        width = int(bounds[1] - bounds[0])          # the largest shard
        nccl = dist.get_backend(group) == 'nccl'
        device = torch.device('cuda', torch.cuda.current_device()) if nccl \
            else torch.device('cpu')
        send, recv, send_host, recv_host = self._buffers(width, world, device)
        stride = self._width
        if nccl:
            send_host.numpy()[:hi - lo] = mine
            send.copy_(send_host, non_blocking=True)
            dist.all_gather_into_tensor(recv, send, group=group)
            recv_host.copy_(recv, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            flat = recv_host.numpy()
        else:
            send.numpy()[:hi - lo] = mine
            pieces = list(recv.view(world, stride).unbind(0))
            dist.all_gather(pieces, send, group=group)
            flat = recv.numpy()
        out = np.empty(n_rows, dtype=np.float64)
        for r in range(world):
            count = int(bounds[r + 1] - bounds[r])
            out[bounds[r]:bounds[r + 1]] = flat[r * stride:r * stride + count]
        return out


def sharded_lnlike(evaluate, thetas, group=None):
    """One-shot form of :class:`ShardedEvaluator` (allocates its buffers per call)."""
    return ShardedEvaluator(evaluate, group)(thetas)


def sharded_lnlike_device(engine, theta_dev, n_rows, ld, send, recv, stream, group=None,
                          row_offset=0):
    """Device-resident form (NCCL): ``theta_dev`` is a float64 CUDA tensor holding the
    FULL batch on every rank; this rank's engine evaluates its shard straight into
    ``send`` and one ``all_gather_into_tensor`` fills ``recv`` -- laid out
    [world][len(send)], i.e. the (B,) lnL vector itself when B divides evenly.
    Everything is enqueued on ``stream`` (the current torch stream); nothing blocks.

    :return: (bounds, width) of the split
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(n_rows, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if hi > lo:
        engine.lnlike_device(theta_dev.data_ptr() + (row_offset + lo) * ld * 8, hi - lo, ld,
                             send.data_ptr(), stream=stream.cuda_stream)
    if world > 1:
        dist.all_gather_into_tensor(recv, send, group=group)
    return bounds, int(bounds[1] - bounds[0])


class PeerExchange(object):
    """Device-resident sharded evaluation whose lnL gather goes over PEER MEMORY instead
    of NCCL (include/psfmc_b200.h, psfmc_lnlike_batch_exchange): every rank's engine owns
    a mailbox on its GPU, the ranks map one another's mailboxes through CUDA IPC (NVLink),
    and a call publishes this rank's lnL with plain stores into every mailbox, raises one
    flag per peer and waits for the peers' flags -- one small kernel behind the lnL
    kernels. torch.distributed is used once, to all-gather the 64-byte IPC handles.

    :param engine: this rank's LikelihoodEngine (one device)
    :param capacity: the largest gathered vector (walkers per call over all ranks)
    """

    def __init__(self, engine, capacity, group=None):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        handle = engine.peer_create(int(capacity))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        engine.peer_connect(self.rank, handles)
        dist.barrier(group=group)
        self.capacity = int(capacity)
        engine._peer_exchange = self      # (an engine has one mailbox: later users share it)

    def lnlike(self, theta_dev, n_rows, ld, gathered, stream, row_offset=0):
        """``theta_dev``: float64 CUDA tensor holding the FULL batch (identical on every
        rank), rows ``row_offset .. row_offset + n_rows`` of it are the batch of this call;
        ``gathered``: float64 CUDA tensor of at least ``n_rows`` elements that receives
        the lnL of all rows, or None: the values then stay in this rank's mailbox
        (``engine.peer_gathered()``). Everything is enqueued on ``stream``; nothing blocks."""
        bounds = shard_bounds(n_rows, self.world)
        lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
        self.engine.lnlike_exchange(
            theta_dev.data_ptr() + (row_offset + lo) * ld * 8, hi - lo, ld, lo, n_rows,
            gathered.data_ptr() if gathered is not None else 0, stream=stream.cuda_stream)
        return bounds


class ShardedPool(object):
    """``pool.map`` for emcee where every rank runs the same sampler (same seed)
    and each evaluates only its shard of every (half-)ensemble."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self._evaluator = ShardedEvaluator(model.log_posterior_batch, group)
        self._fp32_enough = None        # pool.float32_is_enough, probed once

    def map(self, func, iterable):
        thetas = iterable if isinstance(iterable, list) else list(iterable)
        if not thetas:
            return []
        from itertools import repeat
        from .pool import rows_as_block
        block = rows_as_block(thetas)
        lnpost = self._library_lnpost(block)
        if lnpost is None:
            lnpost = self._evaluator(block)
        return list(zip(lnpost.tolist(), repeat({})))

    def _library_lnpost(self, block):
        """The whole sharded evaluation in one library call per rank
        (``psfmc_lnpost_batch_sharded``: priors on every rank, this rank's share of the
        lnL, gathered over peer memory) once a prior plan has been validated and the ranks'
        mailboxes are connected; None: the torch.distributed path."""
        import os
        import torch.distributed as dist
        if os.environ.get('PSFMC_NATIVE_SAMPLER', '1') == '0' or len(block) < 2:
            return None
        if not (dist.is_available() and dist.is_initialized()) or \
                dist.get_backend(self.group) != 'nccl':
            return None
        engine = getattr(self.model, 'engine', None)
        if engine is None or not hasattr(engine, 'lnpost_sharded') or \
                not hasattr(self.model, 'native_sampler_plan'):
            return None
        holder = self.model.native_sampler_plan(block)
        if holder is None:
            return None
        if not self._float32_is_enough(block):
            return None
        peer = getattr(engine, '_peer_exchange', None)
        if peer is None:
            try:
                peer = PeerExchange(engine, max(len(block), 65536), self.group)
            except Exception:
                return None
        if len(block) > peer.capacity:
            return None
        return engine.lnpost_sharded(holder['plan'], block)

    def _float32_is_enough(self, rows):
        """The sharded library calls have no float64 repeat; the torch.distributed path
        has (every rank's host call). Probed once per pool on the first rows it sees, on
        every rank for ALL of them -- the same deterministic answer on every rank, no
        collective -- see :func:`psfmc_b200.pool.float32_is_enough`."""
        if self._fp32_enough is None:
            from .pool import float32_is_enough
            self._fp32_enough = float32_is_enough(self.model, rows)
        return self._fp32_enough

    def map_batch(self, func, block):
        """Array protocol of this package's sampler (cf. BatchPool.map_batch):
        (B, D) -> ((B,) lnpost, None)."""
        block = np.ascontiguousarray(block, dtype=np.float64)
        lnpost = self._library_lnpost(block)
        if lnpost is not None:
            return lnpost, None
        return self._evaluator(block), None

    def native_sampler(self, start_positions):
        """The sampler's loop inside the library for a one-process-per-GPU job
        (``psfmc_ensemble_run`` with ``PSFMC_ENS_SHARDED``): every rank runs the same
        seeded loop, evaluates its contiguous share of every half-ensemble and the lnL of
        all rows is gathered over peer memory (:class:`PeerExchange`). NCCL groups on
        CUDA devices only; otherwise -- and with ``PSFMC_NATIVE_SAMPLER=0`` -- None: the
        numpy loop with :meth:`map_batch`. Non-finite float32 results are -inf on this
        path (no float64 repeat): a model whose starting walkers need the repeat stays on
        the numpy loop and the torch.distributed gather (:meth:`_float32_is_enough`)."""
        import os
        import torch.distributed as dist
        if os.environ.get('PSFMC_NATIVE_SAMPLER', '1') == '0':
            return None
        if not (dist.is_available() and dist.is_initialized()) or \
                dist.get_backend(self.group) != 'nccl':
            return None
        engine = getattr(self.model, 'engine', None)
        if engine is None or not hasattr(engine, 'ensemble_run') or \
                not hasattr(self.model, 'native_sampler_plan'):
            return None
        holder = self.model.native_sampler_plan(start_positions)
        # (every rank must take the same branch: the plan is validated on the same rows)
        if holder is None:
            return None
        if not self._float32_is_enough(start_positions):
            return None
        need = max(int(len(start_positions)), 2)
        peer = getattr(engine, '_peer_exchange', None)
        if peer is None:
            try:
                peer = PeerExchange(engine, max(need, 65536), self.group)
            except Exception:            # no peer access between the ranks' devices
                return None
        if need // 2 > peer.capacity:
            return None
        sharded = dict(holder)
        sharded['sharded'] = True
        # proposals, priors and acceptance on every rank's device (PSFMC_ENS_DEVICE): with the
        # GPU share of a half-ensemble shrinking as 1 / ranks, the host part of the loop is
        # what limits a sharded run otherwise. PSFMC_DEVICE_LOOP=0 keeps it on the host.
        if os.environ.get('PSFMC_DEVICE_LOOP', 'auto') != '0' and not holder['python_columns']:
            sharded['device_loop'] = True
        return engine, sharded

    def close(self):
        pass

    def join(self):
        pass
