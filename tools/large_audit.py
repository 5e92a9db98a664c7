"""float32 engine (with the float64 repeat) against the float64 engine on a large
prior-drawn C1 ensemble: finiteness agreement and |dlnL| statistics (no per-walker bound:
that needs the oracle's images, see tools/tolerance_audit.py for the bounded subset).
    python tools/large_audit.py [walkers = 65536] > profiles/rN_fp32_large_audit.json"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    m64 = MultiComponentModel(bench.build_components('c1'), precision='fp64')
    m32 = MultiComponentModel(bench.build_components('c1'), precision='fp32')
    thetas = draw_walkers_fast(m64, n, seed=2026)
    l64 = np.concatenate([m64.log_likelihood_batch(thetas[i:i + 8192])
                          for i in range(0, n, 8192)])
    l32 = np.concatenate([m32.log_likelihood_batch(thetas[i:i + 2048])
                          for i in range(0, n, 2048)])
    finite = np.isfinite(l64)
    err = np.abs(l32 - l64)[finite]
    rel = err / np.abs(l64[finite])
    info = m32.engine.info()
    # |dlnL| by |lnL| (the absolute statement of DESIGN.md 4.5 is read off these rows)
    mag = np.abs(l64[finite])
    bins = []
    for lo_edge, hi_edge in ((0, 5e4), (5e4, 2e5), (2e5, 1e6), (1e6, np.inf)):
        pick = (mag >= lo_edge) & (mag < hi_edge)
        if pick.any():
            bins.append({'abs_lnl': [lo_edge, None if np.isinf(hi_edge) else hi_edge],
                         'walkers': int(pick.sum()),
                         'max_abs_dlnl': float(err[pick].max()),
                         'p99_abs_dlnl': float(np.percentile(err[pick], 99)),
                         'median_abs_dlnl': float(np.median(err[pick])),
                         'max_rel_dlnl': float(rel[pick].max())})
    form = err / (0.05 + 2e-6 * mag)
    print(json.dumps({
        'workload': 'c1', 'walkers': n, 'finite_fp64': int(finite.sum()),
        'finiteness_agrees': bool(np.array_equal(np.isfinite(l32), finite)),
        'repeated_in_fp64': int(info['rescued_total']),
        'repeated_inside_graphs': int(info['rescued_on_device']),
        'graph_replays': int(info['graph_replays']),
        'max_abs_dlnl': float(err.max()), 'median_abs_dlnl': float(np.median(err)),
        'p99_abs_dlnl': float(np.percentile(err, 99)),
        'max_rel_dlnl': float(rel.max()), 'median_rel_dlnl': float(np.median(rel)),
        'lnl_range': [float(l64[finite].min()), float(l64[finite].max())],
        'by_abs_lnl': bins,
        'max_err_over_0.05_plus_2e-6_abs_lnl': float(form.max()),
        'p999_err_over_0.05_plus_2e-6_abs_lnl': float(np.percentile(form, 99.9))},
        indent=1))


if __name__ == '__main__':
    main()
