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


class ReferencePriors(object):
    """The reference model's priors as a ``psfmc_prior_plan`` (include/psfmc_b200.h): the
    reference spends ~1 ms per walker in scalar scipy calls (``ComponentBase.log_priors``,
    ComponentBase.py:121-129) -- a 2048-walker batch waits seconds for what the GPU answers
    in 0.3 ms. Its priors are ``psfMC.distributions`` objects wrapping a frozen scipy
    distribution (``prior.rv_frozen``, distributions.py:115-128): Uniform / Normal columns
    (and WeibullMinimum ones, if they agree to 4 ulps) are evaluated by the library from
    constants scipy computed, every other family by one vectorised ``prior.logp`` call per
    prior through the plan's callback; sums in the reference's order (insertion order of
    ``_priors`` per component, ``reff_b > reff`` rule of Sersic.py:41-45, components in
    model order). ``validate`` compares the plan with the reference's OWN scalar
    ``model.log_priors()`` on rows of the first batch: the plan is only used if it
    reproduces them (bit for bit; 4 ulps with Weibull columns in the library)."""

    def __init__(self, model, lib, weibull):
        import ctypes
        from . import _lib
        from scipy.stats import _continuous_distns as _cd
        from scipy import stats
        self.model, self.lib = model, lib
        comps = list(model.components)
        ndim = int(sum(np.size(p.value) for c in comps for p in c._priors.values()))
        self.ndim = ndim
        columns = (_lib.PriorColumn * max(ndim, 1))()
        families = {'uniform_gen': (_lib.PRIOR_UNIFORM, 0), 'norm_gen': (_lib.PRIOR_NORMAL, 0)}
        if weibull:
            families['weibull_min_gen'] = (_lib.PRIOR_WEIBULL_MIN, 1)
        self.weibull_native = False
        terms, rules, self.other, start = [], [], [], 0
        for num, comp in enumerate(comps):
            where = {}
            for attr in sorted(comp._priors):
                prior = comp._priors[attr]
                length = int(np.size(prior.value))
                where[attr] = (start, length)
                cols = list(range(start, start + length))
                start += length
                rv = getattr(prior, 'rv_frozen', None)
                native = False
                if rv is not None and isinstance(rv.dist, stats.rv_continuous):
                    family, nargs = families.get(type(rv.dist).__name__, (None, 0))
                    try:
                        args, loc, scale = rv.dist._parse_args(*rv.args, **rv.kwds)
                        cast = [np.broadcast_to(np.asarray(v, dtype=np.float64),
                                                (length,)).copy()
                                for v in tuple(args) + (loc, scale)]
                    except Exception:
                        family = None
                    if family is not None and len(cast) - 2 == nargs:
                        shape_args, loc, scale = cast[:-2], cast[-2], cast[-1]
                        with np.errstate(all='ignore'):
                            valid = np.broadcast_to(
                                rv.dist._argcheck(*shape_args) & (scale > 0), (length,))
                            log_scale = np.log(scale)
                            log_shape = np.log(shape_args[0]) if nargs else None
                        for k, col in enumerate(cols):
                            entry = columns[col]
                            entry.family, entry.theta_index = family, col
                            entry.valid = int(bool(valid[k]))
                            entry.loc, entry.scale = float(loc[k]), float(scale[k])
                            entry.log_scale = float(log_scale[k])
                            entry.log_norm = float(_cd._norm_pdf_logC)
                            if nargs:
                                entry.shape = float(shape_args[0][k])
                                entry.log_shape = float(log_shape[k])
                        native = True
                        self.weibull_native |= family == _lib.PRIOR_WEIBULL_MIN
                if not native:
                    discrete = rv is not None and isinstance(rv.dist, stats.rv_discrete)
                    self.other.append((prior, cols, discrete))
            for attr in comp._priors:                 # the order the reference adds them in
                terms.append((num,) + where[attr])
            if type(comp).__name__ == 'Sersic':       # Sersic.py:41-45
                rule = _lib.PriorRule()
                rule.component = num
                for tag, attr in (('a', 'reff'), ('b', 'reff_b')):
                    if attr in where:
                        setattr(rule, tag + '_index', where[attr][0])
                    else:
                        setattr(rule, tag + '_index', -1)
                        setattr(rule, tag + '_value',
                                float(np.ravel(comp._constants[attr])[0]))
                rules.append(rule)
        self.columns = columns
        self.terms = (_lib.PriorTerm * max(len(terms), 1))()
        for k, (num, first, length) in enumerate(terms):
            self.terms[k].component, self.terms[k].first_column = num, first
            self.terms[k].n_columns = length
        self.rules = (_lib.PriorRule * max(len(rules), 1))(*rules)
        self.n_terms, self.n_rules, self.n_components = len(terms), len(rules), len(comps)

        def other_columns(user, theta_p, n_batch, ld, logp_p, ld_logp):
            try:
                block = np.ctypeslib.as_array(theta_p, shape=(n_batch, ld))
                logp = np.ctypeslib.as_array(logp_p, shape=(n_batch, ld_logp))
                self._other(block, logp)
                return 0
            except Exception:                  # never unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1

        self._callback = _lib.OTHER_COLUMNS_FN(other_columns)
        plan = _lib.PriorPlan()
        plan.columns = ctypes.cast(columns, ctypes.POINTER(_lib.PriorColumn))
        plan.terms = ctypes.cast(self.terms, ctypes.POINTER(_lib.PriorTerm))
        plan.rules = ctypes.cast(self.rules, ctypes.POINTER(_lib.PriorRule))
        plan.n_columns, plan.n_terms = ndim, self.n_terms
        plan.n_rules, plan.n_components = self.n_rules, self.n_components
        if self.other:
            plan.other_columns = self._callback
        self.plan = plan

    def _other(self, block, logp):
        with np.errstate(all='ignore'):
            for prior, cols, discrete in self.other:
                values = block[:, cols]
                if discrete:
                    values = np.rint(values).astype(int)
                logp[:, cols] = np.asarray(prior.logp(values), dtype=np.float64)

    def lnprior(self, block):
        """Joint log-prior of every row through the plan's own entry points."""
        import ctypes
        from . import _lib
        block = np.ascontiguousarray(block, dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        logp = np.empty((len(block), max(self.ndim, 1)))
        _lib.check(self.lib, self.lib.psfmc_prior_columns(
            self.columns, self.ndim, block.ctypes.data_as(dbl_p), len(block), block.shape[1],
            logp.ctypes.data_as(dbl_p), logp.shape[1]))
        self._other(block, logp)
        out = np.empty(len(block))
        _lib.check(self.lib, self.lib.psfmc_prior_sum(
            logp.ctypes.data_as(dbl_p), len(block), logp.shape[1], block.ctypes.data_as(dbl_p),
            block.shape[1], self.terms, self.n_terms, self.rules, self.n_rules,
            self.n_components, out.ctypes.data_as(dbl_p)))
        return out

    def validate(self, block, rows=32):
        """The plan against the reference's own ``log_priors()`` on up to ``rows`` rows."""
        block = np.asarray(block, dtype=np.float64)[:rows]
        want = np.empty(len(block))
        for row, theta in enumerate(block):
            self.model.param_values = theta
            want[row] = self.model.log_priors()
        got = self.lnprior(block)
        if np.array_equal(got, want, equal_nan=True):
            return True
        if not self.weibull_native:
            return False
        finite = np.isfinite(want)
        return bool(np.array_equal(np.isfinite(got), finite) and
                    np.array_equal(np.isnan(got), np.isnan(want)) and
                    np.all(np.abs(got[finite] - want[finite]) <=
                           4 * np.spacing(np.abs(want[finite]))))


class ReferenceBatchPool(object):
    """``pool.map`` for emcee with a reference model: one GPU batch per call for the
    likelihood; the priors batched through :class:`ReferencePriors` once that has
    reproduced the reference's own per-walker ``log_priors()`` on rows of the first batch
    (``PSFMC_BRIDGE_PRIORS=reference`` keeps the per-walker code)."""

    def __init__(self, model, precision='fp32', devices=None, library=None):
        self.model = model
        self.engine, self.num_params = engine_for_reference_model(
            model, precision=precision, devices=devices, library=library)
        self.priors = None          # None: undecided, False: the reference's own code

    def _decide(self, block):
        import os
        self.priors = False
        if os.environ.get('PSFMC_BRIDGE_PRIORS', 'plan') == 'reference':
            return
        strict = os.environ.get('PSFMC_PRIORS_STRICT', '0') == '1'
        for weibull in ((False,) if strict else (True, False)):
            try:
                candidate = ReferencePriors(self.model, self.engine._lib, weibull)
                if candidate.validate(block):
                    self.priors = candidate
                    return
            except Exception:
                continue

    def map(self, func, iterable):
        thetas = [np.asarray(p, dtype=np.float64) for p in iterable]
        if not thetas:
            return []
        block = np.stack(thetas)
        if self.priors is None:
            self._decide(block)
        if self.priors:
            from itertools import repeat
            lnpost = self.engine.lnpost(self.priors.plan, block)
            return list(zip(lnpost.tolist(), repeat({})))
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
