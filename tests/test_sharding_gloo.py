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


def _probe_worker(rank, world, port, emu_lib, outdir):
    """One rank, the backend reported as nccl: walks ShardedPool's library branches up to
    the peer-memory setup (which needs CUDA IPC and gives up here)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world), PSFMC_NATIVE_SAMPLER='1')
    os.environ.pop('PSFMC_DEVICE_LOOP', None)
    import warnings
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import PointSource
    from psfmc_b200.distributed import ShardedPool
    from psfmc_b200.distributions import Uniform
    from psfmc_b200.synthetic import synthetic_components
    real_backend = dist.get_backend
    dist.get_backend = lambda group=None: 'nccl'
    size = 32
    centre, box = np.array((size / 2.0, size / 2.0)), np.array((3.0, 3.0))

    def model_with(point_mag):
        comps = synthetic_components(size, 1, dtype=np.float64, psf_size=16)
        comps = [c for c in comps if not isinstance(c, PointSource)]
        comps.append(PointSource(xy=Uniform(loc=centre - box, scale=2 * box),
                                 mag=Uniform(loc=point_mag, scale=0.5)))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return MultiComponentModel(comps, precision='fp32', library=emu_lib)
    out = {}
    for name, mag in (('ordinary', 21.0), ('bright', 8.0)):
        model = model_with(mag)
        start = model.init_params_from_priors(64)
        pool = ShardedPool(model)
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter('always')
            lnpost, _ = pool.map_batch(None, start)
            native = pool.native_sampler(start)
            listed = np.array([r[0] for r in pool.map(None, [row for row in start])])
        out[name + '_enough'] = pool._fp32_enough
        out[name + '_warned'] = sum('float64 repeat' in str(w.message) for w in caught)
        out[name + '_native_none'] = native is None
        out[name + '_same'] = bool(np.array_equal(lnpost, model.log_posterior_batch(start))
                                   and np.array_equal(listed, lnpost))
        out[name + '_rescued'] = model.engine.info()['rescued_total']
    dist.get_backend = real_backend
    np.savez(os.path.join(outdir, 'probe.npz'), **out)
    dist.destroy_process_group()


def test_sharded_pool_keeps_the_float64_repeat_when_float32_is_short(emu_library, tmp_path):
    """ShardedPool's library paths (psfmc_lnpost_batch_sharded, the sharded loops) have no
    float64 repeat: a model whose walkers need it stays on the torch.distributed gather of
    host calls, decided once per pool by a probe every rank runs on all rows."""
    import torch.multiprocessing as mp
    mp.spawn(_probe_worker, args=(1, _free_port(), emu_library, str(tmp_path)), nprocs=1,
             join=True)
    got = np.load(str(tmp_path / 'probe.npz'))
    assert bool(got['ordinary_enough']) and int(got['ordinary_warned']) == 0
    assert bool(got['ordinary_same']) and int(got['ordinary_rescued']) == 0
    assert not bool(got['bright_enough']) and int(got['bright_warned']) == 1
    assert bool(got['bright_native_none']) and bool(got['bright_same'])
    assert int(got['bright_rescued']) > 0
