#!/usr/bin/env python
"""
Fuzz of the kernel SOURCES on the CPU emulator (tests/emu) against the oracle, on the C1
frames with parameter vectors drawn from a box far WIDER than the model's priors (centres
outside the frame, reff 0.05 ... 300 px, axis ratios down to 0.005, indices 0.05 ... 12,
magnitudes 14 ... 32): what a user's model with other priors could hand the engine.

    python tools/emu_fuzz.py [n_walkers] [seed] [wide|typical] [brightest_mag]

Prints, per precision mode, the rows whose result disagrees with the oracle: finiteness,
fp64 beyond FP64_RTOL, fp32 beyond the stated bound of tests/conftest.py (fp32_bounds).
Test tool: imports the oracle; no GPU needed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)


BOXES = {
    # name: (brightest magnitude, centre margin outside the frame, reff range, smallest
    #        axis ratio, index range)
    'wide': (14.0, 20.0, (0.05, 300.0), 0.005, (0.05, 12.0)),
    # what psfMC models of real frames use (the J0005-0006 example: reff 1.5 ... 12 px,
    # index 0.5 ... 8, magnitudes 20.5 ... 27.5), with a margin on every side
    'typical': (18.0, 0.0, (0.5, 60.0), 0.05, (0.3, 10.0)),
}


def draw(rng, count, box='wide', bright=None):
    mag0, margin, (r_lo, r_hi), q_lo, (n_lo, n_hi) = BOXES[box]
    if bright is not None:
        mag0 = bright

    def logu(lo, hi, size):
        return np.exp(rng.uniform(np.log(lo), np.log(hi), size))
    cols = [rng.uniform(-0.05, 0.05, count), rng.uniform(mag0 + 1.0, 30.0, count),
            rng.uniform(-0.5 * margin, 128.0 + 0.5 * margin, count),
            rng.uniform(-0.5 * margin, 128.0 + 0.5 * margin, count)]
    for _ in range(2):
        reff = logu(r_lo, r_hi, count)
        cols += [rng.uniform(-360.0, 720.0, count), logu(n_lo, n_hi, count),
                 rng.uniform(mag0, 32.0, count), reff,
                 reff * rng.uniform(q_lo, 1.0, count),
                 rng.uniform(-margin, 128.0 + margin, count),
                 rng.uniform(-margin, 128.0 + margin, count)]
    return np.ascontiguousarray(np.stack(cols, axis=1))


def extended_bounds(model, thetas, oracle):
    """fp32_bounds plus the variance channel: the convolved model variance V = raw^2 (*) PSF
    variance comes out of the same float32 transform with an error relative to ITS largest
    value on the frame (masked pixels and wrapped-around wings included), and moves lnL by
    1/2 sum resid^2 ivm^2 dV -- negligible for a walker that fits (resid^2 ivm ~ 1, V << the
    pixel variance), not for a bright component on top of a frame it does not fit."""
    from conftest import FP32_ATOL, FP32_ULPS
    good = ~np.asarray(model.config.bad_px, dtype=bool)
    obs_var = np.asarray(model.config.obs_var, dtype=np.float64)
    out = []
    for theta in thetas:
        img = oracle.images(theta, with_point_source_subtracted=False)
        res = np.asarray(img['residual'], dtype=np.float64)[good]
        ivm_full = np.asarray(img['composite_ivm'], dtype=np.float64)
        ivm = ivm_full[good]
        with np.errstate(all='ignore'):
            var = 1.0 / ivm_full - obs_var
            vmax = np.nanmax(np.where(np.isfinite(var), np.abs(var), 0.0))
            local = np.sum(np.abs(res) * ivm *
                           np.abs(np.asarray(img['convolved_model'], dtype=np.float64)[good]))
            out.append(FP32_ATOL + FP32_ULPS * 2.0 ** -24 *
                       (local + 0.5 * np.sum(res * res * ivm * ivm) * vmax))
    return np.array(out)


def main():
    from conftest import (EMU_LIB, FP64_RTOL, fp32_bounds, model_from_file,
                          oracle_from_model)
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    box = sys.argv[3] if len(sys.argv) > 3 else 'wide'
    bright = float(sys.argv[4]) if len(sys.argv) > 4 else None
    thetas = draw(np.random.RandomState(seed), count, box, bright)
    names = None
    worst = {}
    for precision, env in (('fp64', {}), ('fp32', {}), ('fp32', {'PSFMC_FORCE_STAGED': '1'})):
        for key, val in env.items():
            os.environ[key] = val
        model = model_from_file('j0005/model_c1.py', precision, library=EMU_LIB,
                                obs_dtype=np.float64)
        for key in env:
            os.environ.pop(key)
        if names is None:
            names = model.param_names
            oracle = oracle_from_model(model)
            with np.errstate(all='ignore'):
                expect = oracle.lnlike_batch(thetas)
            bounds = fp32_bounds(model, thetas, oracle)
            extended = extended_bounds(model, thetas, oracle)
        rescued0 = model.engine.info()['rescued_total']
        got = model.log_likelihood_batch(thetas)
        rescued = model.engine.info()['rescued_total'] - rescued0
        tag = precision + ('/staged' if env else '') + ' path {}'.format(
            model.engine.info()['path'])
        fin_e, fin_g = np.isfinite(expect), np.isfinite(got)
        bad_fin = np.flatnonzero(fin_e != fin_g)
        both = fin_e & fin_g
        err = np.abs(got - expect)
        limit = bounds if precision == 'fp32' else FP64_RTOL * np.abs(expect)
        with np.errstate(all='ignore'):
            ratio = np.where(both, err / limit, 0.0)
        over = np.flatnonzero(ratio > 1.0)
        print('{}: {} rows, {} finite in the oracle, {} finiteness mismatches, {} beyond the '
              'bound, worst err/bound {:.3g}, float64 repeats {}'.format(
                  tag, count, int(fin_e.sum()), len(bad_fin), len(over), ratio.max(), rescued))
        if precision == 'fp32':
            with np.errstate(all='ignore'):
                rel = np.where(both, err / np.abs(expect), 0.0)
            print('   relative error |dlnL| / |lnL|: median {:.2g}, 99 % {:.2g}, max {:.2g}; beyond '
                  'the extended bound (variance-channel term added): {}'.format(
                      np.median(rel), np.percentile(rel, 99), rel.max(),
                      int(np.sum(both & (err > extended)))))
        for row in list(bad_fin[:5]) + list(over[np.argsort(-ratio[over])][:5]):
            print('   row {}: got {!r} expect {!r} bound {:.3g}'.format(
                row, got[row], expect[row], limit[row]))
            print('      theta', np.array2string(thetas[row], precision=5, max_line_width=200))
        worst[tag] = float(ratio.max())
    return worst


if __name__ == '__main__':
    main()
