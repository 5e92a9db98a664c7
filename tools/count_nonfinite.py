#!/usr/bin/env python
"""How many prior-drawn walkers come back non-finite from the raw float32 kernels
(= how many the float64 rescue re-evaluates): python tools/count_nonfinite.py [workload]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast
    workload = sys.argv[1] if len(sys.argv) > 1 else 'c1'
    raw = MultiComponentModel(bench.build_components(workload), precision='fp32',
                              fp64_rescue=False)
    ref = MultiComponentModel(bench.build_components(workload), precision='fp64')
    for seed in range(4):
        thetas = draw_walkers_fast(raw, 4096, seed=seed)
        lnl = raw.log_likelihood_batch(thetas)
        rows = np.flatnonzero(~np.isfinite(lnl))
        print('seed', seed, 'non-finite', len(rows), 'float64:',
              ref.log_likelihood_batch(thetas[rows]) if len(rows) else [])
        for row in rows[:6]:
            print('   ', np.array2string(thetas[row], precision=5, max_line_width=200))


if __name__ == '__main__':
    main()
