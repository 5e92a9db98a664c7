#!/usr/bin/env python
"""
Audit of the stated float32 tolerance on large prior-drawn
ensembles: float32 engine against the float64 engine (itself gated at 1e-10 against
the reference in the tests), with the per-walker bound of tests/conftest.py computed
from the oracle's images.   python tools/tolerance_audit.py > profiles/rN_fp32_tolerance_audit.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import bench
    from conftest import fp32_bounds, oracle_from_model
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast
    out = []
    for workload, nwalk, nbound in (('c1', 4096, 512), ('c3', 512, 64), ('c4', 128, 24)):
        m64 = MultiComponentModel(bench.build_components(workload), precision='fp64')
        m32 = MultiComponentModel(bench.build_components(workload), precision='fp32')
        thetas = draw_walkers_fast(m64, nwalk, seed=77)
        l64, l32 = m64.log_likelihood_batch(thetas), m32.log_likelihood_batch(thetas)
        finite = np.isfinite(l64)
        assert np.array_equal(np.isfinite(l32), finite)
        # the raw float32 kernels, without the float64 rescue of non-finite results
        raw = MultiComponentModel(bench.build_components(workload), precision='fp32',
                                  fp64_rescue=False).log_likelihood_batch(thetas)
        err = np.abs(l32 - l64)[finite]
        rows = np.flatnonzero(finite)[:nbound]
        bounds = fp32_bounds(m32, thetas[rows], oracle_from_model(m32))
        ratio = np.abs(l32 - l64)[rows] / bounds
        out.append({'workload': workload, 'frame': list(m32.engine.shape),
                    'walkers': int(nwalk), 'finite': int(finite.sum()),
                    'rescued_in_fp64': int(m32.engine.info()['rescued_total']),
                    'raw_fp32_nonfinite_where_fp64_finite':
                        int(np.sum(finite & ~np.isfinite(raw))),
                    'engine_path': {1: 'fused', 2: 'fused-cluster4', 3: 'tiled-4x4'}.get(
                        m32.engine.info()['path'], 'staged'),
                    'max_abs_dlnl': float(err.max()), 'median_abs_dlnl': float(np.median(err)),
                    'max_rel_dlnl': float((err / np.abs(l64[finite])).max()),
                    'bound_checked_on': int(len(rows)),
                    'max_err_over_bound': float(ratio.max()),
                    'median_err_over_bound': float(np.median(ratio))})
        print(json.dumps(out[-1]), file=sys.stderr)
    print(json.dumps({'audit': out, 'bound': '|dlnL| <= 0.01 + FP32_ULPS * 2^-24 * sum_good |resid| '
                      '* ivm * |model| (FP32_ULPS of tests/conftest.py; err_over_bound is '
                      'relative to it)', 'fp32_ulps': __import__('conftest').FP32_ULPS}, indent=1))


if __name__ == '__main__':
    main()
