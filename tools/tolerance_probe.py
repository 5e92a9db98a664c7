#!/usr/bin/env python
"""
Per-walker data behind tools/tolerance_audit.py for one workload: float32 (raw, no
rescue) and float64 lnL of prior-drawn walkers plus the oracle-image statistics a
reliability criterion could use. Writes an .npz for offline analysis.
    python tools/tolerance_probe.py c1 1024 gpurun_out/probe_c1.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import bench
    from conftest import oracle_from_model, FP32_ATOL, FP32_ULPS
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast
    workload, nwalk, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    m64 = MultiComponentModel(bench.build_components(workload), precision='fp64')
    m32 = MultiComponentModel(bench.build_components(workload), precision='fp32',
                              fp64_rescue=False)
    thetas = draw_walkers_fast(m64, nwalk, seed=77)
    l64, l32 = m64.log_likelihood_batch(thetas), m32.log_likelihood_batch(thetas)
    oracle = oracle_from_model(m32)
    good = ~np.asarray(m32.config.bad_px, dtype=bool)
    ovar = np.asarray(m32.config.obs_var, dtype=np.float64)
    stats = np.zeros((nwalk, 6))
    for num, theta in enumerate(thetas):
        im = oracle.images(theta, with_point_source_subtracted=False)
        with np.errstate(all='ignore'):
            res, ivm, conv = (np.asarray(im[k], dtype=np.float64) for k in
                              ('residual', 'composite_ivm', 'convolved_model'))
            cv = 1.0 / ivm - ovar
            stats[num] = (
                FP32_ATOL + FP32_ULPS * 2.0 ** -24 * np.sum(np.abs(res[good]) * ivm[good] * np.abs(conv[good])),
                np.max(np.abs(conv)), np.max(np.abs(cv[np.isfinite(cv)])),
                np.sqrt(np.sum((res[good] * ivm[good]) ** 2)),
                0.5 * np.sqrt(np.sum(((1.0 - res[good] ** 2 * ivm[good]) * ivm[good]) ** 2)),
                np.max(np.abs(im['raw_model'])))
    np.savez(out, thetas=thetas, l64=l64, l32=l32, stats=stats,
             ovar_min=np.min(ovar[good]))
    err = np.abs(l32 - l64)
    ratio = err / stats[:, 0]
    print('ratio quantiles 50/90/99/max', np.nanquantile(ratio, [0.5, 0.9, 0.99, 1.0]),
          'n>1:', int(np.sum(~(ratio <= 1))))


if __name__ == '__main__':
    main()
