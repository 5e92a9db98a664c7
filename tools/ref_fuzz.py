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

    python tools/ref_fuzz.py [n_thetas] [seed] [wide|typical] [c1|c1_2psf]
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
    model_file = os.path.join(ROOT, 'tests', 'golden', 'j0005',
                              'model_c1_2psf.py' if two_psf else 'model_c1.py')
    layout = refshim.build_reference_model(model_file, 'M1')
    names, lens = [], []        # (the reference's param_names sums ragged lists: numpy 1 only)
    for comp in layout.components:
        names += list(comp.stochastic_names())
        lens += list(comp.stochastic_lens())
    thetas = draw(np.random.RandomState(seed), count, box, None, names, lens, (128, 128),
                  2 if two_psf else 1)
    assert thetas.shape[1] == layout.num_params
    if two_psf:
        # PSF index last: both PSFs, half-integers included (rint is half-to-even). An index
        # that rounds outside the list raises IndexError in the reference -- or wraps around
        # for -1; its prior is -inf there and the likelihood is never reached
        index = np.random.RandomState(seed + 1).uniform(-0.49, 1.49, count)
        index[:4] = (0.5, 1.4999, -0.4999, 0.49999)
        thetas[:, -1] = index
    raw_inputs = mg.raw_inputs_j0005(two_psf)
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
