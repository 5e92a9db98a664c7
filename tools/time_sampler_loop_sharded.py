#!/usr/bin/env python
"""The library's sampler loop with one process per GPU (run under torchrun): rate and, with
PSFMC_ENS_PROFILE=1, rank 0's host-time profile.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/time_sampler_loop_sharded.py [walkers] [iterations]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.distributed import ShardedPool
    from psfmc_b200.sampler import EnsembleSampler
    local = int(os.environ['LOCAL_RANK'])
    if local != 0:
        os.environ.pop('PSFMC_ENS_PROFILE', None)
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    model = MultiComponentModel(os.path.join(ROOT, 'examples', 'model_J0005-0006.py'),
                                devices=[local])
    with open(os.path.join(ROOT, 'tests', 'golden', 'c1_golden.json')) as fobj:
        centre = np.array(json.load(fobj)['theta'][0])
    start = centre + 1e-3 * np.random.RandomState(2).standard_normal((walkers, len(centre))) * \
        np.maximum(np.abs(centre), 1.0)
    smp = EnsembleSampler(walkers, len(centre), model.log_posterior, kwargs={'model': model},
                          pool=ShardedPool(model), live_dangerously=True)
    smp._random.seed(7)
    pos, lnp, _ = smp.run_mcmc(start, iters)
    smp.reset()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smp.run_mcmc(pos, iters, lnprob0=lnp)
    dt = time.perf_counter() - t0
    if dist.get_rank() == 0:
        print('{} ranks, {} walkers: {:.3f} M evals/s, {:.1f} us per half-ensemble'.format(
            dist.get_world_size(), walkers, walkers * iters / dt / 1e6, 1e6 * dt / iters / 2))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
