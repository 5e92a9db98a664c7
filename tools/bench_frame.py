#!/usr/bin/env python
"""lnL evaluations/s of the float32 engine on an arbitrary synthetic frame:
python tools/bench_frame.py HEIGHT WIDTH PSF_H PSF_W [WALKERS]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from conftest import arbitrary_frame_model
    from psfmc_b200.synthetic import draw_walkers_fast
    dims = tuple(int(v) for v in sys.argv[1:5])
    walkers = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
    model = arbitrary_frame_model(*dims, precision='fp32')
    thetas = draw_walkers_fast(model, walkers, seed=1)
    half = walkers // 2
    for _ in range(3):
        model.engine.lnlike(thetas[:half])
    steps = 20
    t0 = time.perf_counter()
    for _ in range(steps):
        model.engine.lnlike(thetas[:half])
        model.engine.lnlike(thetas[half:])
    rate = walkers * steps / (time.perf_counter() - t0)
    info = model.engine.info()
    print(dims, 'path', info['path'], 'walkers', walkers,
          'end-to-end evals/s %.0f' % rate)


if __name__ == '__main__':
    main()
