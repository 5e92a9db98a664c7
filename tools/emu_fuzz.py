#!/usr/bin/env python
"""
Fuzz of the kernel SOURCES on the CPU emulator (tests/emu) against the oracle, on the C1
frames with parameter vectors drawn from a box far WIDER than the model's priors (centres
outside the frame, reff 0.05 ... 300 px, axis ratios down to 0.005, indices 0.05 ... 12,
magnitudes 14 ... 32): what a user's model with other priors could hand the engine.

    python tools/emu_fuzz.py [n_walkers] [seed] [wide|typical|hot] [c1|mixed128|mixed256|mixed512|frame75x100]
                              [brightest_mag]

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


def draw(rng, count, box='wide', bright=None, names=None, lens=None, shape=(128, 128),
         n_psf=1):
    """(count, D) parameter vectors for a model whose parameters are ``names`` / ``lens``
    (MultiComponentModel.param_names / param_lens; default: the C1 layout)."""
    mag0, margin, (r_lo, r_hi), q_lo, (n_lo, n_hi) = BOXES[box]
    if bright is not None:
        mag0 = bright
    if names is None:
        names = ['0_Sky_adu', '1_PointSource_mag', '1_PointSource_xy']
        lens = [1, 1, 2]
        for comp in (2, 3):
            names += ['%d_Sersic_%s' % (comp, attr) for attr in
                      ('angle', 'index', 'mag', 'reff', 'reff_b', 'xy')]
            lens += [1, 1, 1, 1, 1, 2]
    height, width = shape
    scale = max(height, width) / 128.0

    def logu(lo, hi):
        return np.exp(rng.uniform(np.log(lo), np.log(hi), count))
    cols, reff_of = [], {}
    for name, length in zip(names, lens):
        comp, attr = name.split('_', 1)[0], name.split('_', 2)[-1]
        point = 'PointSource' in name
        if attr == 'adu':
            cols.append(rng.uniform(-0.05, 0.05, count))
        elif attr == 'mag':
            cols.append(rng.uniform(mag0 + (1.0 if point else 0.0),
                                    30.0 if point else 32.0, count))
        elif attr == 'xy':
            edge = 0.5 * margin if point else margin
            cols.append(rng.uniform(-edge, width + edge, count))
            cols.append(rng.uniform(-edge, height + edge, count))
        elif attr == 'angle':
            cols.append(rng.uniform(-360.0, 720.0, count))
        elif attr == 'index':
            cols.append(logu(n_lo, n_hi))
        elif attr == 'reff':
            reff_of[comp] = logu(r_lo, r_hi * scale)
            cols.append(reff_of[comp])
        elif attr == 'reff_b':
            major = reff_of.get(comp)
            if major is None:                 # reff fixed in the model: a range of its own
                major = logu(r_lo, r_hi * scale)
            cols.append(major * rng.uniform(q_lo, 1.0, count))
        elif 'psf' in name.lower():
            # in range, half-integers (rint is half-to-even) and a little beyond both ends
            cols.append(rng.uniform(-0.75, n_psf - 0.25, count))
        else:
            raise ValueError('emu_fuzz: no rule for parameter ' + name)
        assert length == (2 if attr == 'xy' else 1), name
    return np.ascontiguousarray(np.stack(cols, axis=1))


def hot_centres(rng, model, count):
    """Prior draws whose Sersic centres are moved to within 0.05 ... 5e-6 px of a pixel
    centre at indices 1 ... 10: the reference's centroid correction makes ONE pixel outshine
    the frame by up to 10^13 there (the hot pixels of DESIGN.md 4.5)."""
    np.random.seed(int(rng.randint(2 ** 31 - 1)))       # (the priors draw from numpy's global state)
    thetas = model.init_params_from_priors(count)
    names, column = [], 0
    for name, length in zip(model.param_names, model.param_lens):
        names.append((name, column))
        column += length
    where = dict(names)
    for name, column in names:
        if 'Sersic' in name and name.endswith('_xy'):
            sel = rng.rand(count) < 0.7
            for axis in (0, 1):
                shrink = rng.choice([1.0, 0.1, 0.01, 1e-4], sel.sum())
                thetas[sel, column + axis] = np.rint(thetas[sel, column + axis]) + \
                    rng.uniform(-0.05, 0.05, sel.sum()) * shrink
            index = where.get(name[:-3] + '_index')
            if index is not None:
                thetas[sel, index] = rng.uniform(1.0, 10.0, sel.sum())
    return thetas


def build_model(which, precision, env=()):
    """The fuzzed models: c1 (fused 128^2 / staged), mixed128 (bilinear and clipped point
    sources, fixed parameters, radians), mixed256 (two PSFs, the four-CTA cluster kernel), mixed512 (the same scene on the tiled
    512 x 512 path),
    frame75x100 (75 x 100 frame, 31 x 17 PSF: zero-padded transform frame + fold)."""
    import conftest
    if os.environ.get('PSFMC_EMU_LIB'):          # an experimental emulator build
        conftest.EMU_LIB = os.environ['PSFMC_EMU_LIB']
    for key, val in env:
        os.environ[key] = val
    try:
        if which == 'c1':
            return conftest.model_from_file('j0005/model_c1.py', precision,
                                            library=conftest.EMU_LIB, obs_dtype=np.float64)
        if which == 'mixed128':
            return conftest.mixed_model_128(precision, library=conftest.EMU_LIB)
        if which == 'mixed256':
            return conftest.mixed_model_256(precision, library=conftest.EMU_LIB)
        if which == 'mixed512':
            return conftest.mixed_model_256(precision, library=conftest.EMU_LIB, n=512)
        if which == 'frame75x100':
            return conftest.arbitrary_frame_model(75, 100, 31, 17, precision=precision,
                                                  library=conftest.EMU_LIB, fp64_rescue=True)
        raise SystemExit('emu_fuzz: unknown model ' + which)
    finally:
        for key, _ in env:
            os.environ.pop(key, None)


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
    from conftest import FP64_RTOL, fp32_bounds, oracle_from_model
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    box = sys.argv[3] if len(sys.argv) > 3 else 'wide'
    which = sys.argv[4] if len(sys.argv) > 4 else 'c1'
    bright = float(sys.argv[5]) if len(sys.argv) > 5 else None
    thetas = None
    worst = {}
    modes = [('fp64', ()), ('fp32', ()), ('fp32', (('PSFMC_FORCE_STAGED', '1'),))]
    for precision, env in modes:
        model = build_model(which, precision, env)
        if thetas is None:
            n_psf = len(model.config.psf_selector.psf_images)
            if box == 'hot':
                thetas = hot_centres(np.random.RandomState(seed), model, count)
            else:
                thetas = draw(np.random.RandomState(seed), count, box, bright,
                              model.param_names, model.param_lens,
                              tuple(model.engine.shape), n_psf)
            oracle = oracle_from_model(model)
            # a PSF index that rounds outside the list: the prior is -inf there, the reference
            # never reaches the likelihood (psfMC/models.py:209-211; its PSF list would
            # raise IndexError); the engine answers -inf
            inside = np.ones(count, dtype=bool)
            usable = thetas
            if model.psf_index_slot[0] == 'theta':
                column = model.psf_index_slot[1]
                sel = np.rint(thetas[:, column])
                inside = (sel >= 0) & (sel < n_psf)
                usable = thetas.copy()
                usable[~inside, column] = 0.0
            with np.errstate(all='ignore'):
                expect = oracle.lnlike_batch(usable)
            expect[~inside] = -np.inf
            bounds = fp32_bounds(model, usable, oracle)
            extended = extended_bounds(model, usable, oracle)
        rescued0 = model.engine.info()['rescued_total']
        got = model.log_likelihood_batch(thetas)
        rescued = model.engine.info()['rescued_total'] - rescued0
        tag = '{} {}{} path {}'.format(which, precision, '/staged' if env else '',
                                       model.engine.info()['path'])
        fin_e, fin_g = np.isfinite(expect), np.isfinite(got)
        bad_fin = np.flatnonzero(fin_e != fin_g)
        both = fin_e & fin_g
        with np.errstate(all='ignore'):
            err = np.where(both, np.abs(got - expect), 0.0)
            limit = bounds if precision == 'fp32' else FP64_RTOL * np.abs(expect)
            ratio = np.where(both, err / limit, 0.0)
        over = np.flatnonzero(ratio > 1.0)
        print('{}: {} rows, {} finite in the oracle, {} finiteness mismatches, {} beyond the '
              'bound, worst err/bound {:.3g}, float64 repeats {}'.format(
                  tag, count, int(fin_e.sum()), len(bad_fin), len(over), ratio.max(), rescued))
        with np.errstate(all='ignore'):
            rel = np.where(both, err / np.abs(expect), 0.0)
        line = '   relative error |dlnL| / |lnL|: median {:.2g}, 99 % {:.2g}, max {:.2g}'.format(
            np.median(rel[both]), np.percentile(rel[both], 99), rel.max())
        if precision == 'fp32':
            line += '; beyond the extended bound (variance-channel term added): {}'.format(
                int(np.sum(both & (err > extended))))
        print(line)
        for row in list(bad_fin[:5]) + list(over[np.argsort(-ratio[over])][:5]):
            print('   row {}: got {!r} expect {!r} bound {:.3g}'.format(
                row, got[row], expect[row], limit[row]))
            print('      theta', np.array2string(thetas[row], precision=5, max_line_width=200))
        worst[tag] = float(ratio.max())
        model.engine.close()
    return worst


if __name__ == '__main__':
    main()
