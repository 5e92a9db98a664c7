#!/usr/bin/env python
"""
Differential fuzz of the ORACLE against the UNMODIFIED reference (build container only:
needs /root/reference, imported through oracle/refshim.py): the golden vectors pin the oracle
on prior draws and named edge cases; this drives both with parameter vectors from the wide
box of tools/emu_fuzz.py (centres outside the frame, reff 0.05 ... 300 px, index 0.05 ... 12,
any angle, 60 000 ADU components) through the reference's own raw_model / convolved_model /
residual / composite_ivm / point_source_subtracted (psfMC/models.py:245-306) and compares
the five images and lnL BITWISE in the
three precision modes of SURVEY.md 8c.

    python tools/ref_fuzz.py [n_thetas] [seed] [wide|typical|hot] [c1|c1_2psf|crop100|crop75x100|c2_n0.5|c2_n4.0|...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def main():
    import make_golden as mg
    from emu_fuzz import draw
    from oracle import refshim
    if not refshim.reference_available():
        raise SystemExit('the reference is not present')
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    box = sys.argv[3] if len(sys.argv) > 3 else 'wide'
    which = sys.argv[4] if len(sys.argv) > 4 else 'c1'
    two_psf = which == 'c1_2psf'
    gdir = os.path.join(ROOT, 'tests', 'golden')
    shape = (128, 128)
    if which in ('c1', 'c1_2psf'):
        model_file = os.path.join(gdir, 'j0005',
                                  'model_c1_2psf.py' if two_psf else 'model_c1.py')
        raw_inputs = mg.raw_inputs_j0005(two_psf)
    elif which in ('crop100', 'crop75x100'):
        # the J0005-0006 frames cropped to 100 x 100 / 75 x 100 (frames that are not
        # powers of two: the reference convolves circularly at the image size)
        from psfmc_b200 import fitsio
        jdir = os.path.join(gdir, 'j0005')
        model_file = os.path.join(jdir, 'model_c1_{}.py'.format(which))
        psfs, ivms = mg.raw_inputs_j0005(False)[3:5]
        obs = fitsio.getdata(os.path.join(jdir, 'sci_{}.fits'.format(which)))
        raw_inputs = (obs, fitsio.getdata(os.path.join(jdir, 'ivm_{}.fits'.format(which))),
                      fitsio.getdata(os.path.join(jdir, 'mask_{}.fits'.format(which))) != 0,
                      psfs, ivms, 25.9463)
        shape = obs.shape
    elif which.startswith('c2_n'):
        # the reference's GALFIT fixtures: one Sersic, delta PSF, constant weight
        from psfmc_b200 import fitsio
        qdir = os.path.join(gdir, 'galfit')
        model_file = os.path.join(qdir, 'model_n{}.py'.format(which[4:]))
        gfile = os.path.join(qdir, 'gfsim_n{}.fits.gz'.format(which[4:]))
        raw_inputs = (fitsio.getdata(gfile), fitsio.getdata(os.path.join(qdir, 'ivm_const.fits')),
                      None, [fitsio.getdata(os.path.join(qdir, 'psf_delta.fits'))],
                      [fitsio.getdata(os.path.join(qdir, 'psfivm_delta.fits'))],
                      float(fitsio.getheader(gfile)['MAGZPT']))
    else:
        raise SystemExit('ref_fuzz: unknown model ' + which)
    layout = refshim.build_reference_model(model_file, 'M1')
    names, lens = [], []        # (the reference's param_names sums ragged lists: numpy 1 only)
    for comp in layout.components:
        names += list(comp.stochastic_names())
        lens += list(comp.stochastic_lens())
    rng = np.random.RandomState(seed)
    thetas = draw(rng, count, 'typical' if box == 'hot' else box, None, names, lens, shape,
                  2 if two_psf else 1)
    if box == 'hot':
        # Sersic centres 0.05 ... 5e-6 px from a pixel centre (or exactly on it), index
        # 1 ... 10: one pixel of up to 10^13 ADU, the reference's own float64 limits
        column = 0
        for name, length in zip(names, lens):
            if 'Sersic' in name and name.endswith('xy'):
                for axis in (0, 1):
                    shrink = rng.choice([1.0, 0.1, 0.01, 1e-4, 0.0], count)
                    thetas[:, column + axis] = np.rint(thetas[:, column + axis]) + \
                        rng.uniform(-0.05, 0.05, count) * shrink
            if 'Sersic' in name and name.endswith('index'):
                thetas[:, column] = rng.uniform(1.0, 10.0, count)
            column += length
    assert thetas.shape[1] == layout.num_params
    if two_psf:
        # PSF index last: both PSFs, half-integers included (rint is half-to-even). An index
        # that rounds outside the list raises IndexError in the reference -- or wraps around
        # for -1; its prior is -inf there and the likelihood is never reached
        index = np.random.RandomState(seed + 1).uniform(-0.49, 1.49, count)
        index[:4] = (0.5, 1.4999, -0.4999, 0.49999)
        thetas[:, -1] = index
    for mode in mg.MODES:
        model = refshim.build_reference_model(model_file, mode)
        oracle = mg.oracle_for(model, mode, raw_inputs)
        bad_img = {key: 0 for key in ('raw_model', 'convolved_model', 'residual',
                                      'composite_ivm', 'point_source_subtracted')}
        bad_lnl = finite = 0
        worst = 0.0
        for theta in thetas:
            lnl, _, imgs = mg.ref_images_and_lnl(model, theta)
            with np.errstate(all='ignore'):      # (param_values are set: models.py:296-306)
                imgs['point_source_subtracted'] = model.point_source_subtracted()
                o_imgs = oracle.images(theta)
            for key in bad_img:
                if not np.array_equal(imgs[key], o_imgs[key], equal_nan=True):
                    bad_img[key] += 1
            with np.errstate(all='ignore'):
                o_lnl = oracle.lnlike(theta)
            o_lnl = o_lnl if np.isfinite(o_lnl) else float('-inf')
            finite += int(np.isfinite(lnl))
            if o_lnl != lnl:
                bad_lnl += 1
                if np.isfinite(lnl) and np.isfinite(o_lnl):
                    worst = max(worst, abs(o_lnl - lnl) / abs(lnl))
                else:
                    worst = float('inf')
                if bad_lnl <= 3:
                    print('   lnL differs: reference {!r} oracle {!r}\n      theta {}'.format(
                        lnl, o_lnl, np.array2string(theta, precision=6, max_line_width=200)))
        print('{} {} {}: {} thetas ({} finite in the reference): lnL differing {} (worst '
              'relative {:.3g}); images differing {}'.format(
                  which, box, mode, count, finite, bad_lnl, worst, bad_img))


if __name__ == '__main__':
    main()
