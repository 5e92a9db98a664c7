"""
MultiComponentModel: the host-side model object with the reference's interface
(/root/reference/psfMC/models.py:9-306), backed by the CUDA engine.

``log_posterior(theta, model=...)`` keeps the reference's static signature and
return value ``(lnL + lnprior, blobs)`` so it can still be handed to emcee as is;
the fast path is :meth:`log_posterior_batch`, which :class:`psfmc_b200.pool.BatchPool`
calls once per emcee ``map`` with the whole (half-)ensemble.
"""
import os

import numpy as np

from .components import Configuration, PointSource, Sersic
from .engine import LikelihoodEngine
from .model_parser import component_list_from_file
from .program import compile_program

IMAGE_TYPES = ('raw_model', 'convolved_model', 'residual', 'composite_ivm',
               'point_source_subtracted')


class MultiComponentModel(object):
    """
    :param components: list of components (one Configuration among them) or the
        name of a model file
    :param precision: engine precision ('fp32' default, 'fp64', 'fp64_rawf32')
    :param devices: CUDA ordinals to shard walker batches over
    :param fp64_rescue: 'fp32' only: repeat non-finite float32 evaluations in float64
        on the GPU (default True; see include/psfmc_b200.h PSFMC_DESC_NO_FP64_RESCUE)
    :param with_blobs: if True, ``log_posterior`` returns the five images per
        evaluation like the reference (slow: they travel to the host); the default
        returns empty blobs and posterior images are re-rendered from the chain
    """

    def __init__(self, components, precision='fp32', devices=None,
                 with_blobs=False, library=None, fp64_rescue=True):
        if isinstance(components, str):
            components = component_list_from_file(components)
        components = list(components)
        configs = [comp for comp in components
                   if isinstance(comp, Configuration)
                   or type(comp).__name__ == 'Configuration']
        if not configs:
            raise ValueError('Unable to find the Configuration component, required '
                             'for setting up input images.')
        config = configs[-1]
        components.remove(config)
        components.append(config.psf_selector)      # PSF selector goes last
        for count, comp in enumerate(components):
            comp.update_stochastic_names(count=count)

        self.config = config
        self.components = components
        self.psf_comps = [comp for comp in components
                          if isinstance(comp, PointSource)]
        self.obs_header = config.obs_header
        self.with_blobs = with_blobs
        self.precision = precision

        self.program, self.psf_index_slot, self._num_params = \
            compile_program(components)
        selector = config.psf_selector
        # batches of at least this many rows overlap the priors with the GPU call
        self.overlap_min_batch = 512
        self.engine = LikelihoodEngine(
            config.obs_data, config.obs_var, config.bad_px,
            selector.psf_images, selector.var_images, config.mag_zeropoint,
            self.program, self.psf_index_slot, precision=precision,
            devices=devices, library=library, fp64_rescue=fp64_rescue)

        self._param_vector = np.zeros(self.num_params)
        self.posterior_images = {}
        self.accumulated_samples = 0
        self.reset_images()
        if precision == 'fp32':
            self._warn_if_float32_is_short()

    def float32_dynamic_range(self):
        """Brightest model pixel the priors allow (the 0.1 % quantile of every magnitude /
        radius prior, or the constant) over the median noise of an unmasked pixel. The
        float32 transform carries errors relative to the LARGEST values on the frame
        ('outside the priors'): up to ~1e5 they stay below the stated
        tolerance, beyond it ``precision='fp64'`` is the mode to use."""
        good = ~np.asarray(self.config.bad_px, dtype=bool)
        if not good.any():
            return 0.0
        sigma = float(np.sqrt(np.median(np.asarray(self.config.obs_var,
                                                   dtype=np.float64)[good])))
        def low_end(comp, attr):
            prior = getattr(comp, '_priors', {}).get(attr)
            if prior is not None:
                frozen = getattr(prior, 'rv_frozen', None)
                return float(np.min(frozen.ppf(1e-3) if frozen is not None
                                    else prior.median()))
            if attr in getattr(comp, '_constants', {}):
                return float(np.min(comp._constants[attr]))
            return None
        peak = 0.0
        for comp in self.components:
            mag = low_end(comp, 'mag')
            if mag is None:
                continue
            flux = 10.0 ** (0.4 * (float(self.config.mag_zeropoint) - mag))
            # a point source puts its flux into a pixel or two; a Sersic profile about
            # 1 / (reff reff_b) of it into its brightest pixel
            reff, reff_b = low_end(comp, 'reff'), low_end(comp, 'reff_b')
            if reff is not None and reff_b is not None and reff * reff_b > 1.0:
                flux /= reff * reff_b
            peak = max(peak, flux)
        if not np.isfinite(peak) or not sigma > 0:
            return 0.0
        return peak / sigma

    def _warn_if_float32_is_short(self):
        try:
            ratio = self.float32_dynamic_range()
        except Exception:                 # a prior without a quantile function: no advice
            return
        if ratio > 1.0e5:
            import warnings
            warnings.warn(
                'psfmc_b200: the priors allow a component {:.1e} times brighter than the '
                'median pixel noise; the float32 likelihood is stated for up to ~1e5 (host '
                'calls repeat what comes out non-finite in float64, device loops do not). '
                'Consider precision=\'fp64\'.'.format(ratio))

    # -- parameter bookkeeping -------------------------------------------------
    @property
    def num_params(self):
        return int(self._num_params)

    @property
    def param_names(self):
        return [name for comp in self.components for name in comp.stochastic_names()]

    @property
    def param_fits_abbrs(self):
        return [name for comp in self.components
                for name in comp.stochastic_names(name_attr='fitsname')]

    @property
    def param_lens(self):
        return [length for comp in self.components for length in comp.stochastic_lens()]

    @property
    def param_values(self):
        splits = np.cumsum(self.param_lens)[:-1]
        return dict(zip(self.param_names, np.split(self._param_vector, splits)))

    @param_values.setter
    def param_values(self, value_vector):
        value_vector = np.asarray(value_vector, dtype=np.float64)
        self._param_vector = value_vector
        start = 0
        for comp in self.components:
            count = comp.num_stochastics()
            comp.set_stochastic_values(value_vector[start:start + count])
            start += count

    def get_distribution(self, param_name):
        for comp in self.components:
            try:
                return comp.get_distribution(param_name)
            except KeyError:
                continue
        return None

    def init_params_from_priors(self, nwalkers):
        """Starting positions drawn from the priors, redrawing per component until
        its joint prior is finite (cf. models.py:108-130). Same distribution as the
        reference's per-walker loop, but every prior is drawn for all walkers in one
        ``rvs`` call and the component's joint prior is evaluated once per block
        (the scalar loop cost 0.4 s for 250 walkers -- half of the example run); the
        draws come from numpy's global random state like the reference's."""
        nwalkers = int(nwalkers)
        start_positions = np.zeros((nwalkers, self.num_params))
        column = 0
        for comp in self.components:
            free = comp.free_parameters()
            width = comp.num_stochastics()
            if width == 0:
                continue
            block = np.empty((nwalkers, width))
            todo = np.arange(nwalkers)
            for _ in range(1000):
                cols = []
                for _, prior, length in free:
                    size = (len(todo), length) if length > 1 else len(todo)
                    draw = np.asarray(prior.random(size=size), dtype=np.float64)
                    cols.append(draw.reshape(len(todo), length))
                drawn = np.concatenate(cols, axis=1)
                block[todo] = drawn
                with np.errstate(all='ignore'):
                    ok = np.isfinite(comp.log_priors_batch(drawn))
                todo = todo[~ok]
                if len(todo) == 0:
                    break
            else:
                raise RuntimeError('could not draw starting positions with a finite '
                                   'prior for component {}'.format(type(comp).__name__))
            start_positions[:, column:column + width] = block
            column += width
        # leave the components at the last walker's values, like the scalar loop did
        if nwalkers:
            self.param_values = start_positions[-1]
        return start_positions

    # -- priors ------------------------------------------------------------------
    def log_priors(self):
        return np.sum([comp.log_priors() for comp in self.components])

    def _log_priors_per_component(self, thetas, column_logp=None):
        total = np.zeros(thetas.shape[0])
        start = 0
        for comp in self.components:
            count = comp.num_stochastics()
            if count or hasattr(comp, 'log_priors_batch'):
                cols = None if column_logp is None else column_logp[:, start:start + count]
                total = total + comp.log_priors_batch(thetas[:, start:start + count], cols)
            start += count
        return total

    def _prior_groups(self):
        """Priors of one scipy family share one array evaluation: per family the
        theta columns and their (shape arguments, loc, scale), stacked."""
        groups, others, start = {}, [], 0
        for comp in self.components:
            for _, prior, length in comp.free_parameters():
                cols = list(range(start, start + length))
                start += length
                spec = prior.direct_spec(length) if hasattr(prior, 'direct_spec') else None
                if spec is None:
                    others.append((prior, cols))
                    continue
                dist, args, loc, scale = spec
                # (a frozen distribution carries its own copy of the generator object:
                # the family is its class and support)
                key = (type(dist), repr(dist.a), repr(dist.b), len(args))
                group = groups.setdefault(key, {
                    'dist': dist, 'cols': [], 'loc': [], 'scale': [],
                    'args': [[] for _ in args]})
                group['cols'] += cols
                group['loc'].append(loc)
                group['scale'].append(scale)
                for store, arg in zip(group['args'], args):
                    store.append(arg)
        for group in groups.values():
            group['cols'] = np.array(group['cols'])
            group['loc'] = np.concatenate(group['loc'])
            group['scale'] = np.concatenate(group['scale'])
            group['args'] = tuple(np.concatenate(arg) for arg in group['args'])
        return list(groups.values()), others

    def _column_logp(self, thetas, groups=None, others=None, out=None):
        """Per-column log-densities (B, D): rv_continuous.logpdf's own operations
        (cf. ScipyDistribution._logpdf_direct), one pass per distribution family."""
        if groups is None:
            groups, others = self._prior_plan
        # parameter-major working layout: the columns of one family are contiguous rows
        out_t = np.empty((thetas.shape[1], thetas.shape[0]), dtype=np.float64)
        with np.errstate(all='ignore'):
            for group in groups:
                dist, scale = group['dist'], group['scale'][:, None]
                args = tuple(arg[:, None] for arg in group['args'])
                theta_g = np.ascontiguousarray(thetas[:, group['cols']].T)
                std = (theta_g - group['loc'][:, None]) / scale
                cond0 = dist._argcheck(*args) & (scale > 0)
                cond = cond0 & dist._support_mask(std, *args) & (scale > 0)
                values = dist._logpdf(std, *args) - np.log(scale)
                logp = np.where(cond, values, -np.inf)
                bad = (1 - cond0) + np.isnan(std)
                if np.any(bad):
                    logp = np.where(bad, dist.badvalue, logp)
                out_t[group['cols']] = logp
            for prior, cols in others:
                block = thetas[:, cols]
                if getattr(prior, 'discrete', False):
                    block = np.rint(block).astype(int)
                out_t[cols] = np.asarray(getattr(prior, 'logp_batch', prior.logp)(block)).T
        if out is None:
            return out_t.T
        for group in groups:
            out[:, group['cols']] = out_t[group['cols']].T
        for _, cols in others:
            out[:, cols] = out_t[cols].T
        return out

    # -- native evaluation (include/psfmc_b200.h: psfmc_prior_columns / _sum) ------
    def _native_prior_plan(self, weibull=False):
        """Tables for the library's host-side prior evaluation: Uniform and Normal
        columns are evaluated there (closed forms whose constants -- log(scale), the
        normal's log sqrt(2 pi) -- are computed HERE by numpy/scipy, so only IEEE-exact
        operations run in C), every other family stays with ``_column_logp``; the
        summation over priors and components runs in C in the reference's order."""
        import ctypes
        from . import _lib
        from scipy.stats import _continuous_distns as _cd
        groups, others = self._prior_plan
        ndim = self.num_params
        columns = (_lib.PriorColumn * max(ndim, 1))()
        families = {'uniform_gen': (_lib.PRIOR_UNIFORM, 0), 'norm_gen': (_lib.PRIOR_NORMAL, 0)}
        if weibull:
            # (libm's log / pow: what numpy 1.21 calls; see PSFMC_PRIOR_WEIBULL_MIN)
            families['weibull_min_gen'] = (_lib.PRIOR_WEIBULL_MIN, 1)
        rest = []
        for group in groups:
            family, nargs = families.get(type(group['dist']).__name__, (None, 0))
            if family is None or len(group['args']) != nargs:
                rest.append(group)
                continue
            dist, scale = group['dist'], group['scale'][:, None]
            args = tuple(arg[:, None] for arg in group['args'])
            with np.errstate(all='ignore'):
                cond0 = np.broadcast_to(dist._argcheck(*args) & (scale > 0), scale.shape)
                log_scale = np.log(scale)
                log_shape = np.log(args[0]) if nargs else None
            for k, col in enumerate(group['cols']):
                entry = columns[int(col)]
                entry.family = family
                entry.theta_index = int(col)
                entry.valid = int(bool(cond0[k, 0]))
                entry.loc = float(group['loc'][k])
                entry.scale = float(group['scale'][k])
                entry.log_scale = float(log_scale[k, 0])
                entry.log_norm = float(_cd._norm_pdf_logC)
                if nargs:
                    entry.shape = float(group['args'][0][k])
                    entry.log_shape = float(log_shape[k, 0])
        terms, rules, start = [], [], 0
        for num, comp in enumerate(self.components):
            where = {}
            for name, _, length in comp.free_parameters():
                where[name] = start
                start += length
            # terms in the order the reference adds them (ComponentBase.prior_terms)
            for name, _, _, length in comp.prior_terms():
                terms.append((num, where[name], length))
            if isinstance(comp, Sersic):
                rule = _lib.PriorRule()
                rule.component = num
                for tag, attr in (('a', 'reff'), ('b', 'reff_b')):
                    if attr in where:
                        setattr(rule, tag + '_index', where[attr])
                    else:
                        setattr(rule, tag + '_index', -1)
                        setattr(rule, tag + '_value',
                                float(np.ravel(comp._constants[attr])[0]))
                rules.append(rule)
        term_arr = (_lib.PriorTerm * max(len(terms), 1))()
        for k, (num, first, length) in enumerate(terms):
            term_arr[k].component, term_arr[k].first_column = num, first
            term_arr[k].n_columns = length
        rule_arr = (_lib.PriorRule * max(len(rules), 1))(*rules)
        return {'columns': columns, 'terms': term_arr, 'n_terms': len(terms),
                'rules': rule_arr, 'n_rules': len(rules), 'rest': rest, 'others': others,
                'n_components': len(self.components), 'ndim': ndim,
                'lib': self.engine._lib}

    def _log_priors_native(self, thetas):
        import ctypes
        from . import _lib
        plan = self._native_plan
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        n_batch, ld = thetas.shape
        ndim = plan['ndim']
        if ld < ndim:
            raise ValueError('theta has too few columns')
        dbl_p = ctypes.POINTER(ctypes.c_double)
        logp = np.empty((n_batch, max(ndim, 1)), dtype=np.float64)
        lib = plan['lib']
        _lib.check(lib, lib.psfmc_prior_columns(
            plan['columns'], ndim, thetas.ctypes.data_as(dbl_p), n_batch, ld,
            logp.ctypes.data_as(dbl_p), logp.shape[1]))
        if plan['rest'] or plan['others']:
            self._column_logp(thetas, plan['rest'], plan['others'], out=logp)
        lnprior = np.empty(n_batch, dtype=np.float64)
        _lib.check(lib, lib.psfmc_prior_sum(
            logp.ctypes.data_as(dbl_p), n_batch, logp.shape[1],
            thetas.ctypes.data_as(dbl_p), ld, plan['terms'], plan['n_terms'],
            plan['rules'], plan['n_rules'], plan['n_components'],
            lnprior.ctypes.data_as(dbl_p)))
        return lnprior

    def native_sampler_plan(self, thetas):
        """The ``psfmc_prior_plan`` for ``psfmc_ensemble_run`` / ``psfmc_lnpost_batch``
        (include/psfmc_b200.h): Uniform / Normal columns and every sum in the library's
        host code, bit-identical to scipy; WeibullMinimum columns there too (the C
        library's log / pow) if that reproduces ``log_priors_batch(thetas)`` to 4 ulps
        and ``PSFMC_PRIORS_STRICT`` is not 1; every other family, custom priors and
        discrete priors through a callback into ``_column_logp``. ``thetas``: the
        validation batch (the starting ensemble). Returns None when no plan reproduces
        the Python priors -- the caller then keeps the Python loop."""
        import ctypes
        from . import _lib
        cached = getattr(self, '_sampler_plan', None)
        if cached is not None:
            return cached or None
        self._sampler_plan = False
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        expect = self.log_priors_batch(thetas)
        if not hasattr(self, '_prior_plan'):
            return None
        strict = os.environ.get('PSFMC_PRIORS_STRICT', '0') == '1'
        dbl_p = ctypes.POINTER(ctypes.c_double)
        for weibull in ((False,) if strict else (True, False)):
            try:
                table = self._native_prior_plan(weibull=weibull)
            except Exception:
                continue
            rest, others = table['rest'], table['others']

            def other_columns(user, theta_p, n_batch, ld, logp_p, ld_logp,
                              rest=rest, others=others):
                try:
                    block = np.ctypeslib.as_array(theta_p, shape=(n_batch, ld))
                    logp = np.ctypeslib.as_array(logp_p, shape=(n_batch, ld_logp))
                    self._column_logp(block, rest, others, out=logp)
                    return 0
                except Exception:          # never unwind through the C frames
                    import traceback
                    traceback.print_exc()
                    return 1

            for col in [c for g in rest for c in g['cols']] + \
                    [c for _, cols in others for c in cols]:
                table['columns'][int(col)].family = _lib.PRIOR_OTHER
            plan = _lib.PriorPlan()
            plan.columns = ctypes.cast(table['columns'], ctypes.POINTER(_lib.PriorColumn))
            plan.terms = ctypes.cast(table['terms'], ctypes.POINTER(_lib.PriorTerm))
            plan.rules = ctypes.cast(table['rules'], ctypes.POINTER(_lib.PriorRule))
            plan.n_columns, plan.n_terms = table['ndim'], table['n_terms']
            plan.n_rules, plan.n_components = table['n_rules'], table['n_components']
            callback = _lib.OTHER_COLUMNS_FN(other_columns)
            if rest or others:
                plan.other_columns = callback
            holder = {'plan': plan, 'table': table, 'callback': callback,
                      'weibull_native': weibull, 'python_columns': bool(rest or others)}
            # validate: lnpost - lnL on the validation batch == the Python priors
            lib = self.engine._lib
            got = np.empty(thetas.shape[0])
            logp = np.empty((thetas.shape[0], max(table['ndim'], 1)))
            try:
                _lib.check(lib, lib.psfmc_prior_columns(
                    table['columns'], table['ndim'], thetas.ctypes.data_as(dbl_p),
                    thetas.shape[0], thetas.shape[1], logp.ctypes.data_as(dbl_p),
                    logp.shape[1]))
                if rest or others:
                    self._column_logp(thetas, rest, others, out=logp)
                _lib.check(lib, lib.psfmc_prior_sum(
                    logp.ctypes.data_as(dbl_p), thetas.shape[0], logp.shape[1],
                    thetas.ctypes.data_as(dbl_p), thetas.shape[1], table['terms'],
                    table['n_terms'], table['rules'], table['n_rules'],
                    table['n_components'], got.ctypes.data_as(dbl_p)))
            except Exception:
                continue
            same = np.array_equal(got, expect, equal_nan=True)
            if not same and weibull:
                finite = np.isfinite(expect)
                same = np.array_equal(np.isfinite(got), finite) and \
                    np.array_equal(np.isnan(got), np.isnan(expect)) and \
                    np.all(np.abs(got[finite] - expect[finite]) <=
                           4 * np.spacing(np.abs(expect[finite])))
            if same:
                self._sampler_plan = holder
                return holder
        return None

    def log_priors_batch(self, thetas):
        """Joint log-prior of every row of ``thetas`` (B, D). Three implementations,
        tried in this order on the first batch and kept only if they reproduce the
        per-prior ``rv_frozen.logpdf`` sums BIT FOR BIT on it: 'native' (Uniform and
        Normal columns and all sums in the library's host code, other families as in
        'grouped'), 'grouped' (priors of one scipy family in one array operation: the
        reference's 11 scalar scipy calls per walker, SURVEY.md section 0.7, become ~3
        array operations per BATCH), 'scipy' (one ``logpdf`` call per prior)."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        mode = getattr(self, '_prior_mode', None)
        if mode == 'scipy' or thetas.shape[0] == 0:
            return self._log_priors_per_component(thetas)
        if mode == 'native':
            return self._log_priors_native(thetas)
        if mode == 'grouped':
            return self._log_priors_per_component(thetas, self._column_logp(thetas))
        slow = self._log_priors_per_component(thetas)
        self._prior_mode = 'scipy'
        try:
            self._prior_plan = self._prior_groups()
        except Exception:
            return slow
        for candidate in ('native', 'grouped'):
            if os.environ.get('PSFMC_PRIORS', candidate) != candidate:
                continue
            try:
                if candidate == 'native':
                    self._native_plan = self._native_prior_plan()
                    fast = self._log_priors_native(thetas)
                else:
                    fast = self._log_priors_per_component(thetas, self._column_logp(thetas))
            except Exception:
                continue
            if np.array_equal(fast, slow, equal_nan=True):
                self._prior_mode = candidate
                break
        return slow

    # -- posterior ---------------------------------------------------------------
    def log_likelihood_batch(self, thetas):
        return self.engine.lnlike(thetas)

    def log_posterior_batch(self, thetas):
        """
        lnL + lnprior for every row. Rows whose prior is not finite get -inf and
        are not sent to the GPU (cf. models.py:209-211).
        """
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        # priors and lnL in ONE library call (psfmc_lnpost_batch: closed-form priors first,
        # dead rows left out, the rest of the priors while the GPU computes) once a prior
        # plan has reproduced the Python priors on a batch (native_sampler_plan);
        # PSFMC_NATIVE_SAMPLER=0 keeps the Python path below
        if os.environ.get('PSFMC_NATIVE_SAMPLER', '1') != '0' and \
                hasattr(self.engine, 'lnpost'):
            holder = getattr(self, '_sampler_plan', None)
            if holder is None and thetas.shape[0] >= 2:
                holder = self.native_sampler_plan(thetas)
            if holder:
                return self.engine.lnpost(holder['plan'], thetas)
        if thetas.shape[0] >= self.overlap_min_batch:
            # large batches: the GPU works on ALL rows (psfmc_lnlike_batch_begin returns
            # once they are enqueued) while the priors are evaluated here; rows with a
            # dead prior are evaluated for nothing and discarded
            self.engine.lnlike_begin(thetas)
            try:
                lnprior = self.log_priors_batch(thetas)
            finally:
                lnl = self.engine.lnlike_end()
            ok = np.isfinite(lnprior) & np.isfinite(lnl)
            with np.errstate(invalid='ignore'):
                return np.where(ok, lnl + lnprior, -np.inf)
        lnprior = self.log_priors_batch(thetas)
        lnpost = np.full(thetas.shape[0], -np.inf)
        alive = np.isfinite(lnprior)
        if alive.any():
            lnl = self.engine.lnlike(thetas[alive])
            lnpost[alive] = np.where(np.isfinite(lnl), lnl + lnprior[alive], -np.inf)
        return lnpost

    @staticmethod
    def log_posterior(param_values, **kwargs):
        """emcee-compatible single evaluation: ``(lnpost, blobs)``
        (cf. models.py:193-243)."""
        model = kwargs.pop('model')
        theta = np.asarray(param_values, dtype=np.float64)
        lnpost = float(model.log_posterior_batch(theta[None, :])[0])
        if not model.with_blobs:
            return lnpost, {}
        if not np.isfinite(model.log_priors_batch(theta[None, :])[0]):
            return -np.inf, {}
        return lnpost, model.sample_images(theta)

    # -- images ------------------------------------------------------------------
    def sample_images(self, theta=None, which=IMAGE_TYPES):
        theta = self._param_vector if theta is None else theta
        imgs = self.engine.render(np.asarray(theta, dtype=np.float64)[None, :], which)
        return {name: arr[0] for name, arr in imgs.items()}

    def raw_model(self):
        return self.sample_images(which=('raw_model',))['raw_model']

    def convolved_model(self, raw_px=None):
        return self.sample_images(which=('convolved_model',))['convolved_model']

    def composite_ivm(self, raw_px=None):
        return self.sample_images(which=('composite_ivm',))['composite_ivm']

    def residual(self, convolved_px=None, raw_px=None):
        return self.sample_images(which=('residual',))['residual']

    def point_source_subtracted(self):
        return self.sample_images(
            which=('point_source_subtracted',))['point_source_subtracted']

    def reset_images(self):
        shape = self.config.obs_data.shape
        self.accumulated_samples = 0
        for img_type in IMAGE_TYPES:
            # ones, not zeros: the IVM image is averaged in variance space and
            # 1/0 * 0 would poison it (weight of the initial value is 0 anyway)
            self.posterior_images[img_type] = np.ones(shape, dtype=np.float64)

    def accumulate_images(self, sample_images):
        """Running per-pixel mean over samples; the composite IVM is averaged as a
        variance (cf. models.py:74-97). ``sample_images``: list of blob dicts."""
        with np.errstate(divide='ignore'):
            self.posterior_images['composite_ivm'] = \
                1 / self.posterior_images['composite_ivm']
            for img_dict in sample_images:
                self.accumulated_samples += 1
                count = self.accumulated_samples
                for img_type, img in img_dict.items():
                    if img_type == 'composite_ivm':
                        img = 1 / img
                    mean = self.posterior_images[img_type]
                    mean *= count - 1
                    mean += img
                    mean /= count
            self.posterior_images['composite_ivm'] = \
                1 / self.posterior_images['composite_ivm']

    def accumulate_from_chain(self, thetas, which=IMAGE_TYPES, batch=4096):
        """Fold the images of every row of ``thetas`` into the running means (the
        reference's re-render path, analysis/images.py:74-83). The images are
        rendered AND summed on the GPU (psfmc_accumulate_batch); only per-pixel
        sums come back, so this costs a few seconds for a whole trace database."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        with np.errstate(divide='ignore', invalid='ignore'):
            for start in range(0, thetas.shape[0], batch):
                block = thetas[start:start + batch]
                sums = self.engine.accumulate(block, which)
                old, new = self.accumulated_samples, self.accumulated_samples + len(block)
                for name, total in sums.items():
                    mean = self.posterior_images[name]
                    if name == 'composite_ivm':      # averaged as a variance
                        mean = 1 / mean if old else np.zeros_like(mean)
                        mean = (mean * old + total) / new
                        self.posterior_images[name] = 1 / mean
                    else:
                        if not old:
                            mean = np.zeros_like(mean)
                        self.posterior_images[name] = (mean * old + total) / new
                self.accumulated_samples = new
