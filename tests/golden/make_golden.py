"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference, imported through oracle/refshim.py) in this
container, and pin the oracle restatement (oracle/psfmc_oracle.py) against it.

    python tests/golden/make_golden.py

Outputs (all committed):
  j0005/sci_psf_b.fits, j0005/ivm_psf_b.fits  synthetic second PSF (for the K=2 case)
  galfit/ivm_const.fits, galfit/psf_delta.fits, galfit/psfivm_delta.fits,
  galfit/model_n*.py                          C2 single-Sersic models on the
                                              reference's GALFIT fixtures
  j0005/{sci,ivm,mask}_crop*.fits, j0005/model_c1_crop*.py
                                              C1 cropped to 100 x 100 and 75 x 100
  c1_golden.json, c1_2psf_golden.json, c1_cropped_golden.json, c2_golden.json
      theta vectors, the reference's lnL in precision modes M1/M2/M3 (SURVEY.md
      section 8c), its lnprior, setup checksums and sampled image pixels
  pointsource_golden.json                     the reference's own known-answer
                                              (tests/test_components.py:121-144)
  c1_pssub_golden.json                        the reference's point_source_subtracted
                                              image (psfMC/models.py:296-306), sampled
                                              pixels + sums; `--pssub` regenerates only this

The data files copied verbatim from the reference are inputs, not source code:
j0005/{sci,ivm}_J0005-0006.fits, j0005/{sci,ivm}_psf.fits,
j0005/mask_J0005-0006.reg (from /root/reference/examples/) and
galfit/gfsim_n*.fits.gz (from /root/reference/tests/).

The script asserts that the oracle reproduces the reference bit-for-bit (lnL and
all four lnL-path images) for every theta in every mode before writing anything.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim                      # noqa: E402
from oracle import psfmc_oracle as orc          # noqa: E402
from psfmc_b200 import fitsio                   # noqa: E402

MODES = ('M1', 'M2', 'M3')


def ref_images_and_lnl(model, theta):
    """Drive the reference's own methods (psfMC/models.py:206-236) without the
    prior early-out, so that theta outside the prior support can be pinned too."""
    theta = np.asarray(theta, dtype=np.float64)
    with np.errstate(all='ignore'):
        model.param_values = theta
        lnprior = float(model.log_priors())
        raw_px = model.raw_model()
        conv_px = model.convolved_model(raw_px)
        resid_px = model.residual(conv_px)
        ivm_px = model.composite_ivm(raw_px)
        good = ~model.config.bad_px
        ivm_flat = ivm_px[good]
        resid_flat = resid_px[good]
        lnl = -0.5 * np.sum(resid_flat ** 2 * ivm_flat
                            - np.log(0.5 / np.pi * ivm_flat))
    lnl = float(lnl)
    if not np.isfinite(lnl):
        lnl = float('-inf')
    imgs = {'raw_model': raw_px, 'convolved_model': conv_px,
            'residual': resid_px, 'composite_ivm': ivm_px}
    return lnl, lnprior, imgs


def program_from_reference_model(model):
    """Translate the reference model's component list into the oracle's neutral
    program (the product's own translation lives in psfmc_b200/program.py and is
    tested against this one)."""
    program, offset = [], 0
    psf_slot = ('const', 0)
    for comp in model.components:
        kind = {'Sky': orc.SKY, 'PointSource': orc.POINT, 'Sersic': orc.SERSIC,
                'PSFSelector': 'psf'}[type(comp).__name__]
        names = sorted(comp._priors.keys())
        lens = comp.stochastic_lens()
        where = {}
        for name, length in zip(names, lens):
            where[name] = (offset, length)
            offset += length

        def slot(attr, sub=None):
            if attr in where:
                return ('theta', where[attr][0] + (sub or 0))
            const = comp._constants[attr]
            return ('const', float(np.ravel(const)[sub or 0]))

        if kind == 'psf':
            psf_slot = slot('psf_index')
            continue
        slots = {}
        for pname in orc.PARAM_NAMES[kind]:
            if pname in ('x', 'y'):
                slots[pname] = slot('xy', 0 if pname == 'x' else 1)
            else:
                slots[pname] = slot(pname)
        flags = {}
        if kind == orc.SERSIC:
            flags['angle_degrees'] = bool(comp.angle_degrees)
        if kind == orc.POINT:
            flags['shift_method'] = comp.shift_method
        program.append((kind, flags, slots))
    return program, psf_slot, offset


def oracle_for(model, mode, raw_inputs):
    program, psf_slot, _ = program_from_reference_model(model)
    obs, ivm, exclude, psfs, psfivms, zp = raw_inputs
    if mode == 'M3':
        obs, ivm = obs.astype(np.float64), ivm.astype(np.float64)
    return orc.build_from_raw_inputs(obs, ivm, exclude, psfs, psfivms, zp,
                                     program, psf_slot,
                                     fft_upcast=(mode != 'M1'))


def pin_and_collect(model_file, raw_inputs, thetas, sample_px):
    """Reference lnL per mode for each theta + bitwise check of the oracle."""
    out = {'lnl': {}, 'lnprior': None, 'pixels': {}}
    for mode in MODES:
        model = refshim.build_reference_model(model_file, mode)
        oracle = oracle_for(model, mode, raw_inputs)
        assert np.array_equal(oracle.bad_px, model.config.bad_px)
        assert np.array_equal(oracle.obs_var, model.config.obs_var)
        for kpsf in range(len(oracle.f_psf)):
            assert np.array_equal(oracle.f_psf[kpsf],
                                  model.config.psf_selector.psf_list[kpsf])
            assert np.array_equal(oracle.f_var[kpsf],
                                  model.config.psf_selector.var_list[kpsf])
        lnls, lnpriors, pixels = [], [], []
        for theta in thetas:
            lnl, lnprior, imgs = ref_images_and_lnl(model, theta)
            o_imgs = oracle.images(theta, with_point_source_subtracted=False)
            for key in imgs:
                assert np.array_equal(imgs[key], o_imgs[key], equal_nan=True), \
                    (mode, key, list(theta))
            o_lnl = oracle.lnlike(theta)
            assert (o_lnl == lnl) or (np.isnan(o_lnl) and np.isnan(lnl)), \
                (mode, o_lnl, lnl)
            lnls.append(lnl)
            lnpriors.append(lnprior)
            pixels.append({key: [float(v) for v in
                                 np.asarray(imgs[key], dtype=np.float64).flat[sample_px]]
                           for key in imgs})
        out['lnl'][mode] = lnls
        out['lnprior'] = lnpriors
        out['pixels'][mode] = pixels
        good = ~model.config.bad_px
        out['setup'] = {
            'n_good': int(good.sum()),
            'good_index_sum': int(np.flatnonzero(good).sum()),
            'first_good': int(np.flatnonzero(good)[0]),
            'last_good': int(np.flatnonzero(good)[-1]),
            'shape': list(good.shape),
            'num_params': int(model.num_params),
        }
        print('  {} pinned: {} thetas, oracle == reference bitwise'.format(
            mode, len(thetas)))
    return out


def prior_draws(model_file, count, seed):
    model = refshim.build_reference_model(model_file, 'M1')
    np.random.seed(seed)
    return model.init_params_from_priors(count)


def dump(name, payload):
    path = os.path.join(HERE, name)
    with open(path, 'w') as fobj:
        json.dump(payload, fobj, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')


def make_second_psf():
    """A second PSF for the K=2 case: the first one shifted by one pixel
    diagonally blended with itself, re-noised weights. Deterministic."""
    psf = fitsio.getdata(os.path.join(HERE, 'j0005', 'sci_psf.fits'))
    ivm = fitsio.getdata(os.path.join(HERE, 'j0005', 'ivm_psf.fits'))
    rng = np.random.RandomState(7)
    psf_b = (0.8 * psf + 0.1 * np.roll(psf, 1, axis=0)
             + 0.1 * np.roll(psf, -1, axis=1)).astype(np.float32)
    psf_b *= (1 + 0.01 * rng.standard_normal(psf.shape)).astype(np.float32)
    ivm_b = (ivm * (0.9 + 0.2 * rng.random_sample(ivm.shape))).astype(np.float32)
    ivm_b[3, 5] = 0.0          # a zero-weight PSF pixel (psfMC/utils.py:115-117)
    psf_b[60, 2] = np.nan      # and a non-finite one
    fitsio.writeto(os.path.join(HERE, 'j0005', 'sci_psf_b.fits'), psf_b)
    fitsio.writeto(os.path.join(HERE, 'j0005', 'ivm_psf_b.fits'), ivm_b)


def raw_inputs_j0005(two_psf):
    jdir = os.path.join(HERE, 'j0005')
    obs = fitsio.getdata(os.path.join(jdir, 'sci_J0005-0006.fits'))
    ivm = fitsio.getdata(os.path.join(jdir, 'ivm_J0005-0006.fits'))
    from psfmc_b200 import regions
    exclude = ~regions.region_mask_from_file(
        os.path.join(jdir, 'mask_J0005-0006.reg'), obs.shape)
    psfs = [fitsio.getdata(os.path.join(jdir, 'sci_psf.fits'))]
    ivms = [fitsio.getdata(os.path.join(jdir, 'ivm_psf.fits'))]
    if two_psf:
        psfs.append(fitsio.getdata(os.path.join(jdir, 'sci_psf_b.fits')))
        ivms.append(fitsio.getdata(os.path.join(jdir, 'ivm_psf_b.fits')))
    return obs, ivm, exclude, psfs, ivms, 25.9463


def main():
    if not refshim.reference_available():
        raise SystemExit('the reference is not present; cannot regenerate')
    sample_px = [0, 127, 1080, 5000, 8256, 8257, 8320, 8384, 10925, 11070,
                 15176, 16383]

    # ---- C1 -----------------------------------------------------------------
    c1_model = os.path.join(HERE, 'j0005', 'model_c1.py')
    named = {
        'A': [0.001, 21.0, 64.3, 64.1, 30, 2.5, 22.5, 6.0, 4.0, 64.8, 63.9,
              120, 1.2, 24.5, 4.0, 3.0, 46.2, 85.4],
        'B': [-0.004, 20.7, 65.02, 63.55, 95, 4.0, 21.5, 9.5, 3.1, 63.2, 65.7,
              10, 0.7, 25.0, 6.5, 2.2, 44.1, 87.9],
        'D': [0.012, 22.0, 60.0, 70.0, 170, 7.5, 26.5, 11.5, 11.0, 58.25,
              59.125, 60, 3.0, 23.6, 2.1, 2.05, 50.5, 81.0],
        # Sersic-1 centred exactly on a pixel centre: 0/0 -> NaN -> lnL = -inf
        'C_exact_centre': [0.012, 22.0, 60.0, 70.0, 170, 7.5, 26.5, 11.5, 11.0,
                           58.0, 59.0, 60, 3.0, 23.6, 2.1, 2.05, 50.5, 81.0],
        # point source clipped at the lower-left / upper-right frame edges,
        # half-integer positions (round-half-even in the stamp bounds)
        'E_edge_ll': [0.0, 21.0, 1.5, 0.2, 30, 2.5, 22.5, 6.0, 4.0, 64.8, 63.9,
                      120, 1.2, 24.5, 4.0, 3.0, 46.2, 85.4],
        'F_edge_ur': [0.0, 21.0, 126.5, 127.4, 30, 0.3, 22.5, 6.0, 4.0, 64.8,
                      63.9, 120, 9.7, 24.5, 4.0, 3.0, 46.2, 85.4],
        'G_half_int': [0.002, 20.5, 64.5, 63.5, 45, 1.0, 21.0, 5.0, 5.0, 64.25,
                       64.75, 0, 0.5, 25.5, 2.0, 2.0, 44.3, 88.0],
        # exactly on an integer position: sinc(0) branch of the Lanczos kernel
        'H_integer_ps': [0.0, 21.5, 64.0, 65.0, 200, 2.0, 22.0, 8.0, 3.0, 66.3,
                         62.2, -30, 5.0, 24.0, 7.9, 2.1, 41.0, 90.5],
    }
    draws = prior_draws(c1_model, 56, seed=20261018)
    thetas = [list(map(float, v)) for v in named.values()] + \
             [list(map(float, row)) for row in draws]
    print('C1:', len(thetas), 'thetas')
    c1 = pin_and_collect(c1_model, raw_inputs_j0005(False), thetas, sample_px)
    c1['theta'] = thetas
    c1['names'] = list(named.keys())
    c1['sample_px'] = sample_px
    c1['model_file'] = 'j0005/model_c1.py'
    dump('c1_golden.json', c1)

    # ---- C1 with two PSFs ---------------------------------------------------
    make_second_psf()
    c1b_model = os.path.join(HERE, 'j0005', 'model_c1_2psf.py')
    draws = prior_draws(c1b_model, 24, seed=77)
    thetas = [list(map(float, row)) for row in draws]
    # PSF index rounding: half to even, both PSFs exercised
    for num, val in enumerate((0.5, 1.49, 0.51, -0.4, 1.0, 0.0)):
        thetas[num][-1] = val
    print('C1/2psf:', len(thetas), 'thetas')
    c1b = pin_and_collect(c1b_model, raw_inputs_j0005(True), thetas, sample_px)
    c1b['theta'] = thetas
    c1b['sample_px'] = sample_px
    c1b['model_file'] = 'j0005/model_c1_2psf.py'
    dump('c1_2psf_golden.json', c1b)

    # ---- C1 cropped to frames that are not powers of two ------------------------
    # (the reference convolves circularly at the image size, whatever it is:
    # psfMC/utils.py:25-32; odd heights work, odd widths do not, models.py:276)
    crops = {'crop100': (slice(14, 114), slice(14, 114)),      # 100 x 100
             'crop75x100': (slice(30, 105), slice(10, 110))}   # 75 x 100 (odd height)
    cropped = {'cases': {}}
    for tag, (rows, cols) in crops.items():
        obs, ivm, exclude, psfs, ivms, zp = raw_inputs_j0005(False)
        y0, x0 = rows.start, cols.start
        fitsio.writeto(os.path.join(HERE, 'j0005', 'sci_{}.fits'.format(tag)),
                       np.ascontiguousarray(obs[rows, cols]))
        fitsio.writeto(os.path.join(HERE, 'j0005', 'ivm_{}.fits'.format(tag)),
                       np.ascontiguousarray(ivm[rows, cols]))
        fitsio.writeto(os.path.join(HERE, 'j0005', 'mask_{}.fits'.format(tag)),
                       np.ascontiguousarray(exclude[rows, cols]).astype(np.int16))
        mfile = os.path.join(HERE, 'j0005', 'model_c1_{}.py'.format(tag))
        with open(mfile, 'w') as fobj:
            fobj.write(
                "# C1 (model_c1.py) on the J0005-0006 frames cropped to rows {r0}:{r1},\n"
                "# columns {c0}:{c1} -- a frame that is not a power of two; the mask is\n"
                "# the cropped region mask as a FITS image (nonzero = excluded).\n"
                "from numpy import array\n\n"
                "qso_mag = 20.66\n"
                "qso_xy, qso_box = array((64.5 - {c0}, 64.5 - {r0})), array((8, 8))\n"
                "blob_xy, blob_box = array((46 - {c0}, 85.6 - {r0})), array((5, 5))\n\n"
                "Configuration(obs_file='sci_{t}.fits', obsivm_file='ivm_{t}.fits',\n"
                "              psf_files='sci_psf.fits', psfivm_files='ivm_psf.fits',\n"
                "              mask_file='mask_{t}.fits', mag_zeropoint=25.9463)\n"
                "Sky(adu=Normal(loc=0, scale=0.01))\n"
                "PointSource(xy=Uniform(loc=qso_xy - qso_box, scale=2 * qso_box),\n"
                "            mag=Uniform(loc=qso_mag - 0.2, scale=0.2 + 1.5))\n"
                "Sersic(xy=Uniform(loc=qso_xy - qso_box, scale=2 * qso_box),\n"
                "       mag=Uniform(loc=qso_mag, scale=27.5 - qso_mag),\n"
                "       reff=Uniform(loc=2.0, scale=10.0), reff_b=Uniform(loc=2.0, scale=10.0),\n"
                "       index=WeibullMinimum(c=1.5, scale=4),\n"
                "       angle=Uniform(loc=0, scale=180), angle_degrees=True)\n"
                "Sersic(xy=Uniform(loc=blob_xy - blob_box, scale=2 * blob_box),\n"
                "       mag=Uniform(loc=23.5, scale=2.0),\n"
                "       reff=Uniform(loc=2.0, scale=6.0), reff_b=Uniform(loc=2.0, scale=6.0),\n"
                "       index=WeibullMinimum(c=1.5, scale=4),\n"
                "       angle=Uniform(loc=0, scale=180), angle_degrees=True)\n".format(
                    r0=rows.start, r1=rows.stop, c0=cols.start, c1=cols.stop, t=tag))
        draws = prior_draws(mfile, 10, seed=4242 + y0)
        thetas = [list(map(float, row)) for row in draws]
        # a bright point source next to the frame corner: its PSF wings wrap around
        thetas[0][1:4] = [19.5, 1.25, 2.5]
        thetas[1][1:4] = [19.5, (cols.stop - cols.start) - 1.75,
                          (rows.stop - rows.start) - 2.25]
        npx = (rows.stop - rows.start) * (cols.stop - cols.start)
        crop_px = [0, 99, 100, 1234, npx // 2, npx // 2 + 37, npx - 101, npx - 1]
        raw = (np.ascontiguousarray(obs[rows, cols]), np.ascontiguousarray(ivm[rows, cols]),
               np.ascontiguousarray(exclude[rows, cols]), psfs, ivms, zp)
        print('C1/{}:'.format(tag), len(thetas), 'thetas')
        case = pin_and_collect(mfile, raw, thetas, crop_px)
        case['theta'] = thetas
        case['sample_px'] = crop_px
        case['model_file'] = 'j0005/model_c1_{}.py'.format(tag)
        cropped['cases'][tag] = case
    dump('c1_cropped_golden.json', cropped)

    # ---- C2: GALFIT single-Sersic sweep ------------------------------------------
    gdir = os.path.join(HERE, 'galfit')
    fitsio.writeto(os.path.join(gdir, 'ivm_const.fits'),
                   np.full((128, 128), 4.0e6, dtype=np.float32))
    delta = np.zeros((8, 8), dtype=np.float32)
    delta[4, 4] = 1.0
    fitsio.writeto(os.path.join(gdir, 'psf_delta.fits'), delta)
    fitsio.writeto(os.path.join(gdir, 'psfivm_delta.fits'),
                   np.full((8, 8), 1.0e12, dtype=np.float32))
    c2 = {'cases': {}, 'sample_px': sample_px}
    for index in (0.5, 1.0, 3.1, 4.0, 6.5):
        gfile = 'gfsim_n{:0.1f}.fits.gz'.format(index)
        hdr = fitsio.getheader(os.path.join(gdir, gfile))
        par = {key: float(str(hdr[key]).split('+/-')[0])
               for key in ('1_XC', '1_YC', '1_MAG', '1_RE', '1_N', '1_AR', '1_PA')}
        zp = float(hdr['MAGZPT'])
        mfile = os.path.join(gdir, 'model_n{:0.1f}.py'.format(index))
        with open(mfile, 'w') as fobj:
            fobj.write(
                "# C2: single Sersic on the reference's GALFIT fixture {g}\n"
                "# (parameters from its header as in tests/test_components.py:63-74:\n"
                "# xy = GALFIT - 1, reff_b = RE * AR, PA in degrees); delta PSF,\n"
                "# constant weight, no mask. reff_b and angle are fixed constants.\n"
                "from numpy import array\n"
                "Configuration(obs_file='{g}', obsivm_file='ivm_const.fits',\n"
                "              psf_files='psf_delta.fits',\n"
                "              psfivm_files='psfivm_delta.fits',\n"
                "              mag_zeropoint={zp!r})\n"
                "Sersic(xy=Uniform(loc=array(({x!r}, {y!r})) - 2, scale=array((4, 4))),\n"
                "       mag=Uniform(loc={m!r} - 1, scale=2),\n"
                "       reff=Uniform(loc={re!r} - 2, scale=4), reff_b={reb!r},\n"
                "       index=Uniform(loc=0.3, scale=8), angle={pa!r},\n"
                "       angle_degrees=True)\n".format(
                    g=gfile, zp=zp, x=par['1_XC'] - 1, y=par['1_YC'] - 1,
                    m=par['1_MAG'], re=par['1_RE'],
                    reb=par['1_RE'] * par['1_AR'], pa=par['1_PA']))
        # theta order: index, mag, reff, x, y (sorted prior names; xy -> 2 slots)
        header_theta = [par['1_N'], par['1_MAG'], par['1_RE'],
                        par['1_XC'] - 1, par['1_YC'] - 1]
        draws = prior_draws(mfile, 3, seed=int(index * 10))
        thetas = [header_theta] + [list(map(float, row)) for row in draws]
        obs = fitsio.getdata(os.path.join(gdir, gfile))
        raw_inputs = (obs, fitsio.getdata(os.path.join(gdir, 'ivm_const.fits')),
                      None, [fitsio.getdata(os.path.join(gdir, 'psf_delta.fits'))],
                      [fitsio.getdata(os.path.join(gdir, 'psfivm_delta.fits'))], zp)
        print('C2 n={}:'.format(index), len(thetas), 'thetas')
        case = pin_and_collect(mfile, raw_inputs, thetas, sample_px)
        case['theta'] = thetas
        case['model_file'] = 'galfit/model_n{:0.1f}.py'.format(index)
        # the reference test's own (un-asserted) sanity figure: psfMC vs GALFIT
        model = refshim.build_reference_model(mfile, 'M3')
        _, _, imgs = ref_images_and_lnl(model, header_theta)
        frac = np.abs(imgs['raw_model'] - obs) / obs
        case['galfit_max_frac_err'] = float(frac.max())
        case['galfit_median_frac_err'] = float(np.median(frac))
        case['galfit_flux_ratio'] = float(imgs['raw_model'].sum() / obs.sum())
        c2['cases']['{:0.1f}'.format(index)] = case
    dump('c2_golden.json', c2)

    # ---- the reference's own PointSource known answer --------------------------
    refshim.load_reference()
    from psfMC.ModelComponents import PointSource
    from scipy.ndimage import shift
    refarr = np.zeros((5, 5))
    refarr[1, 1] = 1.0
    refarr = shift(refarr, np.array((2.2, 2.7))[::-1] - 1, order=1)
    testarr = np.zeros((5, 5))
    PointSource(xy=np.array((2.2, 2.7)), mag=0,
                shift_method='bilinear').add_to_array(testarr, mag_zp=0)
    assert np.allclose(refarr, testarr)
    lanc = np.zeros((16, 16))
    PointSource(xy=np.array((7.3, 8.6)), mag=0,
                shift_method='lanczos3').add_to_array(lanc, mag_zp=0)
    dump('pointsource_golden.json', {
        'bilinear_xy': [2.2, 2.7], 'bilinear_5x5': testarr.tolist(),
        'scipy_shift_5x5': refarr.tolist(),
        'lanczos_xy': [7.3, 8.6], 'lanczos_16x16': lanc.tolist()})


def make_pssub_golden():
    """c1_pssub_golden.json: the reference's point_source_subtracted image
    (psfMC/models.py:296-306) for 12 C1 and 8 two-PSF parameter vectors in all three
    precision modes, after asserting that the oracle reproduces it BIT FOR BIT."""
    if not refshim.reference_available():
        raise SystemExit('the reference is not present; cannot regenerate')
    with open(os.path.join(HERE, 'c1_golden.json')) as fobj:
        c1 = json.load(fobj)
    with open(os.path.join(HERE, 'c1_2psf_golden.json')) as fobj:
        c1b = json.load(fobj)
    rng = np.random.RandomState(99)
    px = sorted(set(rng.randint(0, 128 * 128, size=96).tolist()
                    + [0, 127, 8256, 8257, 8320, 8384, 16383]))
    out = {'sample_px': px, 'cases': {}}
    for tag, golden, two_psf, count in (('c1', c1, False, 12), ('c1_2psf', c1b, True, 8)):
        model_file = os.path.join(HERE, golden['model_file'])
        thetas = golden['theta'][:count]
        case = {'model_file': golden['model_file'], 'theta': thetas, 'pixels': {},
                'sum': {}, 'abs_sum': {}}
        for mode in MODES:
            model = refshim.build_reference_model(model_file, mode)
            oracle = oracle_for(model, mode, raw_inputs_j0005(two_psf))
            pixels, sums, abs_sums = [], [], []
            for theta in thetas:
                with np.errstate(all='ignore'):
                    model.param_values = np.asarray(theta, dtype=np.float64)
                    ref_px = model.point_source_subtracted()
                mine = oracle.images(theta)['point_source_subtracted']
                assert mine.dtype == ref_px.dtype, (mode, mine.dtype, ref_px.dtype)
                assert np.array_equal(mine, ref_px, equal_nan=True), (tag, mode, theta)
                flat = np.asarray(ref_px, dtype=np.float64).ravel()
                pixels.append([float(v) for v in flat[px]])
                sums.append(float(flat.sum()))
                abs_sums.append(float(np.abs(flat).sum()))
            case['pixels'][mode], case['sum'][mode] = pixels, sums
            case['abs_sum'][mode] = abs_sums
            print('  {} {}: point_source_subtracted pinned for {} thetas, oracle == '
                  'reference bitwise'.format(tag, mode, len(thetas)))
        out['cases'][tag] = case
    dump('c1_pssub_golden.json', out)


if __name__ == '__main__':
    if '--pssub' in sys.argv:
        make_pssub_golden()
    else:
        main()
        make_pssub_golden()
