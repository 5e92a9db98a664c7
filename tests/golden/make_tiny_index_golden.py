"""
Regression vectors for very small Sersic indices (n < 0.01), found by
tools/large_audit.py on a 65536-walker prior-drawn ensemble: the reference's gradient
term g * (sdr / 12 * g) overflows in float64 at the far pixels, the raw model holds a NaN
and lnL is -inf (psfMC/ModelComponents/Sersic.py:129-133, psfMC/models.py:238-241). The
first float32 kernels clamped the term and returned finite values for these walkers.

Imports the UNMODIFIED reference through oracle/refshim.py (needs /root/reference) and
freezes its lnL next to the oracle's:   python tests/golden/make_tiny_index_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import refshim                       # noqa: E402
import make_golden as mg                         # noqa: E402

# theta in the order of model_c1.py (Sky.adu | PS.mag, x, y | Sersic angle, index, mag,
# reff, reff_b, x, y | Sersic ...)
THETAS = [
    [0.0081, 21.6593, 59.7452, 62.8589, 93.6629, 2.4593, 21.6046, 10.2533, 6.6365, 65.9981,
     63.4822, 127.6153, 0.0057, 24.7698, 5.7355, 3.9553, 47.3793, 81.0333],
    [-0.0061, 20.48, 60.8411, 61.6789, 90.5635, 0.0051, 26.9211, 5.4187, 4.1519, 58.3535,
     62.1335, 152.7025, 5.4182, 24.8433, 4.5301, 2.6818, 47.3149, 84.038],
    [0.0191, 21.6675, 58.5104, 69.8417, 131.9984, 0.0058, 21.7065, 8.784, 2.7513, 68.1496,
     61.1156, 163.9371, 1.1937, 23.6645, 7.4787, 5.555, 43.6681, 82.5395],
    [0.0033, 20.6758, 65.0871, 63.5589, 16.5292, 5.6187, 23.5934, 9.1283, 2.8546, 62.8737,
     57.629, 31.944, 0.0065, 25.4418, 7.257, 6.4982, 47.44, 84.4643],
    # the same walkers with an index just large enough for the term to stay finite
    [0.0081, 21.6593, 59.7452, 62.8589, 93.6629, 2.4593, 21.6046, 10.2533, 6.6365, 65.9981,
     63.4822, 127.6153, 0.03, 24.7698, 5.7355, 3.9553, 47.3793, 81.0333],
    [-0.0061, 20.48, 60.8411, 61.6789, 90.5635, 0.03, 26.9211, 5.4187, 4.1519, 58.3535,
     62.1335, 152.7025, 5.4182, 24.8433, 4.5301, 2.6818, 47.3149, 84.038],
]


def main():
    if not refshim.reference_available():
        raise SystemExit('the reference is not present; cannot regenerate')
    mfile = os.path.join(HERE, 'j0005', 'model_c1.py')
    out = {'theta': THETAS}
    for mode in ('M2', 'M3'):
        model = refshim.build_reference_model(mfile, mode)
        lnls = [mg.ref_images_and_lnl(model, theta)[0] for theta in THETAS]
        oracle = mg.oracle_for(model, mode, mg.raw_inputs_j0005(False))
        with np.errstate(all='ignore'):
            check = oracle.lnlike_batch(np.array(THETAS))
        assert np.array_equal(np.array(lnls), check), (lnls, check)   # oracle pinned here too
        out['lnl_' + mode] = [None if not np.isfinite(v) else v for v in lnls]
    with open(os.path.join(HERE, 'c1_tiny_index.json'), 'w') as fobj:
        json.dump(out, fobj, indent=1)
    print(json.dumps(out['lnl_M3']))


if __name__ == '__main__':
    main()
