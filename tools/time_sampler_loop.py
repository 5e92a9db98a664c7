#!/usr/bin/env python
"""The stretch-move loop inside the library (psfmc_ensemble_run) on the C1 model: rate,
host-time profile (PSFMC_ENS_PROFILE), float64 repeats and graph replays per iteration.
    python tools/time_sampler_loop.py [walkers] [iterations] [start: prior|ball] [n_devices]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('PSFMC_ENS_PROFILE', '1')


def main():
    import json
    from psfmc_b200 import BatchPool, MultiComponentModel
    from psfmc_b200.sampler import EnsembleSampler
    from psfmc_b200.synthetic import draw_walkers_fast
    walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    how = sys.argv[3] if len(sys.argv) > 3 else 'prior'
    ndev = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    model = MultiComponentModel(os.path.join(ROOT, 'examples', 'model_J0005-0006.py'),
                                devices=list(range(ndev)))
    ndim = model.num_params
    if how == 'prior':
        start = draw_walkers_fast(model, walkers, seed=1)
    else:
        with open(os.path.join(ROOT, 'tests', 'golden', 'c1_golden.json')) as fobj:
            centre = np.array(json.load(fobj)['theta'][0])
        start = centre + 1e-3 * np.random.RandomState(2).standard_normal((walkers, ndim)) * \
            np.maximum(np.abs(centre), 1.0)
    smp = EnsembleSampler(walkers, ndim, model.log_posterior, kwargs={'model': model},
                          pool=BatchPool(model), live_dangerously=True)
    smp._random.seed(7)
    pos, lnp, _ = smp.run_mcmc(start, iters)      # (buffers sized, graphs captured)
    smp.reset()
    info0 = model.engine.info()
    t0 = time.perf_counter()
    pos, lnp, _ = smp.run_mcmc(pos, iters, lnprob0=lnp)
    dt = time.perf_counter() - t0
    info1 = model.engine.info()
    print('{} device(s), '.format(ndev), end='')
    print('{} walkers, {} start: {:.3f} M evals/s, {:.1f} us per half-ensemble; per iteration: '
          '{:.2f} float64 repeats, {:.2f} graph replays, {:.1f} kernel launches; acceptance {:.3f}'
          .format(walkers, how, walkers * iters / dt / 1e6, 1e6 * dt / iters / 2,
                  (info1['rescued_total'] - info0['rescued_total']) / iters,
                  (info1['graph_replays'] - info0['graph_replays']) / iters,
                  (info1['launches_total'] - info0['launches_total']) / iters,
                  smp.acceptance_fraction.mean()))


if __name__ == '__main__':
    main()
