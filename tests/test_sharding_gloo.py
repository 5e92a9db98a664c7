"""
CPU tier: the N>1 path (one process per GPU) on the gloo backend with
world_size 2: contiguous sharding of a walker batch over ranks, per-rank engines
(the emulated kernels stand in for the GPUs), lnL gathered on every rank, and a
seeded sampler run that is identical with and without sharding.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds():
    from psfmc_b200.distributed import shard_bounds
    assert shard_bounds(10, 4).tolist() == [0, 3, 6, 8, 10]
    assert shard_bounds(2048, 8).tolist() == list(range(0, 2049, 256))
    assert shard_bounds(1, 2).tolist() == [0, 1, 1]
    assert shard_bounds(0, 2).tolist() == [0, 0, 0]


def _free_port():
    with socket.socket() as sock:
        sock.bind(('127.0.0.1', 0))
        return sock.getsockname()[1]


def _worker(rank, world, port, emu_lib, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.distributed import ShardedPool, sharded_lnlike
    from psfmc_b200.sampler import EnsembleSampler
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    model = MultiComponentModel(synthetic_components(32, 1, psf_size=16),
                                precision='fp32', library=emu_lib)
    thetas = draw_walkers_fast(model, 7, seed=1)          # odd: ragged shards
    calls = []

    def evaluate(rows):
        calls.append(len(rows))
        return model.log_likelihood_batch(rows)
    gathered = sharded_lnlike(evaluate, thetas)
    one_row = sharded_lnlike(evaluate, thetas[:1])        # rank 1 gets no rows
    # emcee's list protocol (row views in, (lnpost, blob) tuples out) and the array one
    pool = ShardedPool(model)
    listed = np.array([r[0] for r in pool.map(None, [thetas[i] for i in range(len(thetas))])])
    batched = pool.map_batch(None, thetas)[0]
    assert pool.map(None, []) == []
    nwalk = 2 * model.num_params + 2
    sampler = EnsembleSampler(nwalk, model.num_params, model.log_posterior,
                              kwargs={'model': model}, pool=ShardedPool(model))
    sampler._random.seed(3)
    pos, lnp = sampler.run_mcmc(draw_walkers_fast(model, nwalk, seed=2), 2)[:2]
    np.savez(os.path.join(outdir, 'rank{}.npz'.format(rank)), gathered=gathered,
             one_row=one_row, calls=np.array(calls), pos=pos, lnp=lnp, listed=listed,
             batched=batched)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_two_gloo(emu_library, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, emu_library, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(str(tmp_path / 'rank0.npz'))
    r1 = np.load(str(tmp_path / 'rank1.npz'))
    assert r0['calls'][0] == 4 and r1['calls'][0] == 3          # 7 rows -> 4 + 3
    assert np.array_equal(r0['gathered'], r1['gathered'])
    assert np.array_equal(r0['one_row'], r1['one_row'])
    # the same rows evaluated in one process
    from psfmc_b200 import BatchPool, MultiComponentModel
    from psfmc_b200.sampler import EnsembleSampler
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    model = MultiComponentModel(synthetic_components(32, 1, psf_size=16),
                                precision='fp32', library=emu_library)
    thetas = draw_walkers_fast(model, 7, seed=1)
    assert np.array_equal(model.log_likelihood_batch(thetas), r0['gathered'])
    want = model.log_posterior_batch(thetas)
    for data in (r0, r1):
        assert np.array_equal(data['listed'], want)
        assert np.array_equal(data['batched'], want)
    nwalk = 2 * model.num_params + 2
    sampler = EnsembleSampler(nwalk, model.num_params, model.log_posterior,
                              kwargs={'model': model}, pool=BatchPool(model))
    sampler._random.seed(3)
    pos, lnp = sampler.run_mcmc(draw_walkers_fast(model, nwalk, seed=2), 2)[:2]
    assert np.array_equal(pos, r0['pos']) and np.array_equal(pos, r1['pos'])
    assert np.array_equal(lnp, r0['lnp'])
