"""
One-time setup of the constant arrays the engine replicates on every GPU
(SURVEY.md section 8a row a15). Host-side numpy; same results as the reference's
setup functions, which are cited per function (paths relative to /root/reference).
"""
from math import fsum

import numpy as np

from . import fitsio, regions


def load_image(source):
    """Accept a filename (FITS, optionally gzipped) or an array."""
    if isinstance(source, np.ndarray):
        return source
    return fitsio.getdata(source)


def mask_from_file(mask_file, shape):
    """
    Exclusion mask (True = excluded) from a FITS image (nonzero = excluded) or a
    ds9 region file (pixels outside the region filter are excluded).
    cf. psfMC/utils.py:82-103.
    """
    try:
        return fitsio.getdata(mask_file).astype(bool)
    except IOError:
        pass
    try:
        return ~regions.region_mask_from_file(mask_file, shape)
    except UnicodeDecodeError:
        return None


def preprocess_obs(obs_data, obs_ivm, mask_file=None):
    """
    Observation, variance map (inf at bad pixels) and bad-pixel mask.
    cf. psfMC/utils.py:54-79. Bad = non-finite data or weight, or weight <= 0;
    mask-file exclusions are OR-ed into the mask but leave the variance alone.
    """
    data = load_image(obs_data)
    ivm = load_image(obs_ivm)
    with np.errstate(divide='ignore', invalid='ignore'):
        bad = ~(np.isfinite(data) & np.isfinite(ivm)) | (ivm <= 0)
        var = np.where(bad, np.inf, 1 / ivm)
    if mask_file is not None:
        excluded = mask_file if isinstance(mask_file, np.ndarray) \
            else mask_from_file(mask_file, data.shape)
        if excluded is not None:
            bad = bad | np.asarray(excluded, dtype=bool)
    return data, var, bad


def preprocess_psf(psf_data, psf_ivm):
    """
    Normalised PSF and its variance map. cf. psfMC/utils.py:106-123 and :45-51:
    bad PSF pixels are zeroed in data and weight, the PSF is divided by its
    math.fsum, the weight multiplied by sum**2, variance = 1/weight (0 where the
    weight is 0). Arithmetic stays in the input dtype, like the reference.
    """
    data = np.array(load_image(psf_data))
    ivm = np.array(load_image(psf_ivm))
    bad = ~(np.isfinite(data) & np.isfinite(ivm)) | (ivm <= 0)
    data[bad] = 0
    ivm[bad] = 0
    total = fsum(data.flat)
    data = data / total
    ivm = ivm * total ** 2
    with np.errstate(divide='ignore'):
        var = np.where(ivm <= 0, 0, 1 / ivm)
    return data, var


def add_psf_variability(psfs, psf_vars):
    """With more than one PSF, the per-pixel variance across the (normalised)
    PSFs is added to every variance map. cf. psfMC/utils.py:136-157."""
    if len(psfs) == 1:
        return list(psfs), list(psf_vars)
    mismatch = np.var(psfs, axis=0)
    return list(psfs), [var + mismatch for var in psf_vars]
