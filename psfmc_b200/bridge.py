"""
Bridge for an UNMODIFIED psfMC installation: build the CUDA engine from a reference
``psfMC.MultiComponentModel`` object and hand emcee a pool-like map object, so that
``psfMC.fitting.model_galaxy_mcmc`` needs one extra keyword (``pool=``) and nothing
else.

Only duck-typed attributes of the reference objects are read:
``model.config.{obs_data, obs_var, bad_px, mag_zeropoint}``,
``model.config.psf_selector.{psf_list, var_list}`` (the pre-transformed PSFs,
/root/reference/psfMC/ModelComponents/PSFSelector.py:39-43) and every component's
``_priors`` / ``_constants`` (ComponentBase.py:26-35).
"""
import numpy as np

from .engine import LikelihoodEngine
from .program import compile_program


def engine_for_reference_model(model, precision='fp32', devices=None, library=None):
    """
    :param model: a reference ``psfMC.models.MultiComponentModel``
    :return: (LikelihoodEngine, num_params)

    The reference keeps only the spectra of its zero-padded, normalised PSFs and
    variance maps (utils.py:126-133); transforming them back gives the padded
    frames, which the engine takes as full-size "stamps" (pad offset 0).
    """
    cfg = model.config
    selector = cfg.psf_selector
    shape = cfg.obs_data.shape
    psfs = [np.fft.irfft2(np.asarray(spec, dtype=np.complex128), s=shape)
            for spec in selector.psf_list]
    variances = [np.fft.irfft2(np.asarray(spec, dtype=np.complex128), s=shape)
                 for spec in selector.var_list]
    program, psf_slot, ndim = compile_program(model.components)
    engine = LikelihoodEngine(cfg.obs_data, cfg.obs_var, cfg.bad_px, psfs, variances,
                              cfg.mag_zeropoint, program, psf_slot,
                              precision=precision, devices=devices, library=library)
    return engine, ndim


class ReferenceBatchPool(object):
    """``pool.map`` for emcee with a reference model: one GPU batch per call for the
    likelihood, the reference's own (per-walker) code for the priors."""

    def __init__(self, model, precision='fp32', devices=None, library=None):
        self.model = model
        self.engine, self.num_params = engine_for_reference_model(
            model, precision=precision, devices=devices, library=library)

    def map(self, func, iterable):
        thetas = [np.asarray(p, dtype=np.float64) for p in iterable]
        if not thetas:
            return []
        block = np.stack(thetas)
        lnprior = np.empty(len(block))
        for row, theta in enumerate(block):
            self.model.param_values = theta
            lnprior[row] = self.model.log_priors()
        out = [(-np.inf, {})] * len(block)
        alive = np.isfinite(lnprior)
        if alive.any():
            lnl = self.engine.lnlike(block[alive])
            for row, value in zip(np.flatnonzero(alive), lnl):
                if np.isfinite(value):
                    out[row] = (float(value + lnprior[row]), {})
        return out

    def close(self):
        self.engine.close()

    def join(self):
        pass


def pool_for_reference_model(model, **kwargs):
    return ReferenceBatchPool(model, **kwargs)
