"""Where the time of one BatchPool call goes (C1, 2048 walkers per call): the C-ABI
call alone, the priors alone, both overlapped (log_posterior_batch), and the
thread hand-over itself. Usage: python tools/time_pool_path.py [calls]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_components  # noqa: E402
from psfmc_b200 import BatchPool, MultiComponentModel  # noqa: E402
from psfmc_b200.synthetic import draw_walkers_fast  # noqa: E402


def timed(func, calls):
    for _ in range(5):
        func()
    t0 = time.perf_counter()
    for _ in range(calls):
        func()
    return (time.perf_counter() - t0) / calls * 1e6


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    model = MultiComponentModel(build_components('c1'), precision='fp32', devices=[0])
    th = draw_walkers_fast(model, 2048, seed=5)
    model.log_posterior_batch(th)
    model.log_posterior_batch(th)
    print('prior mode', getattr(model, '_prior_mode', None))
    print('engine.lnlike (numpy in/out)   {:8.1f} us'.format(
        timed(lambda: model.engine.lnlike(th), calls)))
    print('log_priors_batch               {:8.1f} us'.format(
        timed(lambda: model.log_priors_batch(th), calls)))
    print('log_posterior_batch            {:8.1f} us'.format(
        timed(lambda: model.log_posterior_batch(th), calls)))
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1)
    print('thread submit + result (noop)  {:8.1f} us'.format(
        timed(lambda: pool.submit(int).result(), calls)))
    bp = BatchPool(model)
    rows = [th[i] for i in range(len(th))]
    print('BatchPool.map (list in, list of tuples out) {:8.1f} us'.format(
        timed(lambda: bp.map(None, rows), calls)))
    print('BatchPool.map_batch (array in, array out)   {:8.1f} us'.format(
        timed(lambda: bp.map_batch(None, th), calls)))

if __name__ == '__main__':
    main()
