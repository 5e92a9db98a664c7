"""
Synthetic workloads of the named sizes (SURVEY.md section 8d "Concrete inputs"):
C3 = 256^2 quasar + two-Sersic host with a non-trivial PSF-variance term,
C4 = 512^2 point source + three Sersic components; also a 128^2 variant with the
C1 component structure for ensemble-size sweeps. Fixed seeds throughout.
"""
import numpy as np

from .components import Configuration, PointSource, Sersic, Sky
from .distributions import Normal, Uniform


def synthetic_psf(size=64):
    """Gaussian core (sigma 1.5 px) + 5 % halo (sigma 5 px), x1000, peak at
    (size/2, size/2); weight 1/(psf/50 + 0.01) so the variance term matters."""
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    r2 = (xx - size // 2) ** 2 + (yy - size // 2) ** 2
    psf = np.exp(-0.5 * r2 / 1.5 ** 2) + 0.05 * np.exp(-0.5 * r2 / 5.0 ** 2)
    psf *= 1000.0 / psf.max()
    ivm = 1.0 / (psf / 50.0 + 0.01)
    return psf.astype(np.float32), ivm.astype(np.float32)


def synthetic_components(size, n_sersic, seed=None, mask_disc=True, psf_size=64,
                         dtype=np.float32):
    """Component list (Configuration first) for an size x size synthetic frame."""
    seed = 1234 + size if seed is None else seed
    rng = np.random.RandomState(seed)
    obs = (0.02 * rng.standard_normal((size, size))).astype(dtype)
    ivm = np.full((size, size), 2500.0, dtype=dtype)
    mask = None
    if mask_disc:
        yy, xx = np.mgrid[0:size, 0:size]
        mask = ((xx - size / 2.0) ** 2 + (yy - size / 2.0) ** 2) > (0.43 * size) ** 2
    psf, psf_ivm = synthetic_psf(min(psf_size, size))
    centre = np.array((size / 2.0, size / 2.0))
    box = np.array((4.0, 4.0))
    comps = [
        Configuration(obs_file=obs, obsivm_file=ivm, psf_files=psf,
                      psfivm_files=psf_ivm, mask_file=mask, mag_zeropoint=25.9463),
        Sky(adu=Normal(loc=0, scale=0.01)),
        PointSource(xy=Uniform(loc=centre - box, scale=2 * box),
                    mag=Uniform(loc=20, scale=2)),
    ]
    for _ in range(n_sersic):
        comps.append(Sersic(xy=Uniform(loc=centre - box, scale=2 * box),
                            mag=Uniform(loc=21, scale=4),
                            reff=Uniform(loc=6, scale=10),
                            reff_b=Uniform(loc=2, scale=4),
                            index=Uniform(loc=0.5, scale=6.0),
                            angle=Uniform(loc=0, scale=180), angle_degrees=True))
    return comps


WORKLOADS = {
    # name: (frame size, number of Sersic components, walkers per ensemble)
    'c3': (256, 2, 1024),
    'c4': (512, 3, 4096),
    's128': (128, 2, 4096),
}


def draw_walkers(model, nwalkers, seed=0):
    """Seeded walker positions from the model's priors (valid: finite prior)."""
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        return model.init_params_from_priors(nwalkers)
    finally:
        np.random.set_state(state)


def draw_walkers_fast(model, nwalkers, seed=0):
    """Vectorised prior draws (one rvs call per prior); rows violating a prior
    (reff_b > reff) are redrawn. Same distribution as init_params_from_priors,
    different random stream -- used for large synthetic ensembles."""
    rng = np.random.RandomState(seed)
    out = np.empty((0, model.num_params))
    while out.shape[0] < nwalkers:
        need = max(64, 2 * (nwalkers - out.shape[0]))
        cols = []
        for comp in model.components:
            for _, prior, length in comp.free_parameters():
                draw = prior.rv_frozen.rvs(size=(need, length) if length > 1
                                           else need, random_state=rng)
                cols.append(np.asarray(draw, dtype=np.float64).reshape(need, length))
        block = np.concatenate(cols, axis=1) if cols else np.zeros((need, 0))
        ok = np.isfinite(model.log_priors_batch(block))
        out = np.concatenate([out, block[ok]], axis=0)
    return np.ascontiguousarray(out[:nwalkers])


def write_synthetic_files(size, n_sersic, outdir, psf_size=64):
    """The same synthetic workload as :func:`synthetic_components`, written out as
    FITS files plus a model file in psfMC's model-file syntax, so that a model built
    from files (e.g. the reference's own MultiComponentModel in the CPU arm of
    bench.py) sees bit-identical inputs. Returns the model file's path."""
    import os
    from . import fitsio
    comps = synthetic_components(size, n_sersic, psf_size=psf_size)
    config = comps[0]
    rng = np.random.RandomState(1234 + size)
    obs = (0.02 * rng.standard_normal((size, size))).astype(np.float32)
    ivm = np.full((size, size), 2500.0, dtype=np.float32)
    yy, xx = np.mgrid[0:size, 0:size]
    mask = ((xx - size / 2.0) ** 2 + (yy - size / 2.0) ** 2) > (0.43 * size) ** 2
    psf, psf_ivm = synthetic_psf(min(psf_size, size))
    assert np.array_equal(mask | ~np.isfinite(obs), config.bad_px)
    os.makedirs(outdir, exist_ok=True)
    fitsio.writeto(os.path.join(outdir, 'sci.fits'), obs)
    fitsio.writeto(os.path.join(outdir, 'ivm.fits'), ivm)
    fitsio.writeto(os.path.join(outdir, 'mask.fits'), mask.astype(np.int16))
    fitsio.writeto(os.path.join(outdir, 'psf.fits'), psf)
    fitsio.writeto(os.path.join(outdir, 'psf_ivm.fits'), psf_ivm)
    lines = [
        'from numpy import array',
        'centre, box = array(({0}, {0})), array((4.0, 4.0))'.format(size / 2.0),
        "Configuration(obs_file='sci.fits', obsivm_file='ivm.fits', psf_files='psf.fits',",
        "              psfivm_files='psf_ivm.fits', mask_file='mask.fits',",
        '              mag_zeropoint=25.9463)',
        'Sky(adu=Normal(loc=0, scale=0.01))',
        'PointSource(xy=Uniform(loc=centre - box, scale=2 * box),',
        '            mag=Uniform(loc=20, scale=2))',
    ]
    for _ in range(n_sersic):
        lines += [
            'Sersic(xy=Uniform(loc=centre - box, scale=2 * box),',
            '       mag=Uniform(loc=21, scale=4), reff=Uniform(loc=6, scale=10),',
            '       reff_b=Uniform(loc=2, scale=4), index=Uniform(loc=0.5, scale=6.0),',
            '       angle=Uniform(loc=0, scale=180), angle_degrees=True)',
        ]
    model_file = os.path.join(outdir, 'model_synthetic_{}.py'.format(size))
    with open(model_file, 'w') as fobj:
        fobj.write('\n'.join(lines) + '\n')
    return model_file
