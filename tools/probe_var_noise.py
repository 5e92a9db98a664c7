"""Probe: most negative convolved model variance (in units of the observation variance)
of float32 walkers -- a detector for transforms whose rounding noise is no longer small
against the data's variance. Worst walkers of tools/large_audit_dump.py against a random
prior-drawn sample.   python tools/probe_var_noise.py gpurun_out/large_audit_dump.json"""
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
    dump = json.load(open(sys.argv[1]))
    m32 = MultiComponentModel(bench.build_components('c1'), precision='fp32', fp64_rescue=False)
    m64 = MultiComponentModel(bench.build_components('c1'), precision='fp64')
    good = ~np.asarray(m32.config.bad_px, dtype=bool)
    ovar = np.asarray(m32.config.obs_var, dtype=np.float64)

    def stats(thetas):
        out = []
        for i in range(0, len(thetas), 32):
            th = thetas[i:i + 32]
            ivm = m32.engine.render(th, ('composite_ivm',))['composite_ivm']
            l32 = m32.log_likelihood_batch(th)
            l64 = m64.log_likelihood_batch(th)
            for k in range(len(th)):
                with np.errstate(all='ignore'):
                    mvar = 1.0 / ivm[k][good] - ovar[good]
                    ratio = mvar / ovar[good]
                out.append((float(np.nanmin(ratio)), float(abs(l32[k] - l64[k])),
                            float(l64[k])))
        return out

    worst = np.array([m['theta'] for m in dump['worst']])
    print('worst walkers of the large audit: min(mvar/ovar), |dlnL|, lnL64')
    for row in stats(worst):
        print('  %12.4e  %10.3f  %14.1f' % row)
    rand = draw_walkers_fast(m64, 1024, seed=99)
    rows = np.array(stats(rand))
    print('random prior-drawn sample (1024): percentiles of min(mvar/ovar)')
    for q in (0, 0.1, 1, 5, 50):
        print('  p%-4s %12.4e' % (q, np.nanpercentile(rows[:, 0], q)))
    for tol in (1e-1, 1e-2, 1e-3, 1e-4, 1e-5):
        sel = rows[:, 0] < -tol
        print('  tol %.0e: flagged %4d of 1024, max |dlnL| among unflagged %.3f' % (
            tol, int(sel.sum()), float(np.nanmax(rows[~sel, 1]))))


if __name__ == '__main__':
    main()
