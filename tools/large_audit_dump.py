"""Dump the worst walkers of tools/large_audit.py (theta, float32 and float64 lnL) for an
offline look at their bounds with the oracle.  python tools/large_audit_dump.py > out.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast
    n = 65536
    m64 = MultiComponentModel(bench.build_components('c1'), precision='fp64')
    m32 = MultiComponentModel(bench.build_components('c1'), precision='fp32')
    raw = MultiComponentModel(bench.build_components('c1'), precision='fp32', fp64_rescue=False)
    thetas = draw_walkers_fast(m64, n, seed=2026)
    l64 = np.concatenate([m64.log_likelihood_batch(thetas[i:i + 8192]) for i in range(0, n, 8192)])
    l32 = np.concatenate([m32.log_likelihood_batch(thetas[i:i + 2048]) for i in range(0, n, 2048)])
    lraw = np.concatenate([raw.log_likelihood_batch(thetas[i:i + 2048]) for i in range(0, n, 2048)])
    mismatch = np.flatnonzero(np.isfinite(l32) != np.isfinite(l64))
    both = np.isfinite(l32) & np.isfinite(l64)
    err = np.where(both, np.abs(l32 - l64), 0.0)
    worst = np.argsort(-err)[:12]
    out = {'names': m64.param_names, 'mismatch': [], 'worst': []}
    for r in mismatch:
        out['mismatch'].append({'row': int(r), 'theta': thetas[r].tolist(), 'l64': float(l64[r]),
                                'l32': float(l32[r]), 'lraw': float(lraw[r])})
    for r in worst:
        out['worst'].append({'row': int(r), 'theta': thetas[r].tolist(), 'l64': float(l64[r]),
                             'l32': float(l32[r]), 'lraw': float(lraw[r]), 'err': float(err[r])})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
