"""
Prior distributions with the reference's public names (psfMC.distributions):
every descriptive alias the reference defines for a scipy.stats distribution
(/root/reference/psfMC/distributions.py:9-63) exists here under the same name,
takes the same scipy arguments, and exposes the same members (``value``,
``logp``, ``random``, ``median``, ``name``, ``fitsname``, ``rv_frozen``).

Priors stay on the host (BASELINE.json north_star); what changes is that
``logp`` is used column-vectorised over a whole batch of walkers
(psfmc_b200/models.py) instead of once per walker per parameter.
"""
import numpy as np
import scipy.stats as _stats

# scipy.stats name -> public alias. Names scipy has renamed since the reference's
# pinned scipy 1.7.3 are resolved through _RENAMED.
_ALIASES = {
    'alpha': 'Alpha', 'anglit': 'Anglit', 'arcsine': 'Arcsine',
    'bernoulli': 'Bernoulli', 'beta': 'Beta', 'betaprime': 'BetaPrime',
    'binom': 'Binomial', 'boltzmann': 'Boltzmann', 'bradford': 'Bradford',
    'burr': 'Burr3', 'burr12': 'Burr12', 'cauchy': 'Cauchy', 'chi': 'Chi',
    'chi2': 'ChiSquared', 'cosine': 'Cosine', 'dgamma': 'DoubleGamma',
    'dlaplace': 'DiscreteLaplace', 'dweibull': 'DoubleWeibull',
    'erlang': 'Erlang', 'expon': 'Exponential', 'exponnorm': 'ExponentialNormal',
    'exponpow': 'ExponentialPower', 'exponweib': 'ExponentialWeibull', 'f': 'F',
    'fatiguelife': 'FatigueLife', 'fisk': 'Fisk', 'foldcauchy': 'FoldedCauchy',
    'foldnorm': 'FoldedNormal', 'gamma': 'Gamma', 'gausshyper': 'GaussHypergeometric',
    'genexpon': 'GeneralExponential', 'genextreme': 'GeneralExtreme',
    'gengamma': 'GeneralGamma', 'genhalflogistic': 'GeneralHalfLogistic',
    'genlogistic': 'GeneralLogistic', 'gennorm': 'GeneralNormal',
    'genpareto': 'GeneralPareto', 'geom': 'Geometric', 'gilbrat': 'Gilbrat',
    'gompertz': 'Gompertz', 'gumbel_l': 'GumbelLeft', 'gumbel_r': 'GumbelRight',
    'halfcauchy': 'HalfCauchy', 'halfgennorm': 'HalfGeneralNormal',
    'halflogistic': 'HalfLogistic', 'halfnorm': 'HalfNormal',
    'hypergeom': 'Hypergeometric', 'hypsecant': 'HyperbolicSecant',
    'invgamma': 'InverseGamma', 'invgauss': 'InverseGaussian',
    'invweibull': 'InverseWeibull', 'johnsonsb': 'JohnsonSB',
    'johnsonsu': 'JohnsonSU', 'kappa3': 'Kappa3', 'kappa4': 'Kappa4',
    'ksone': 'KSOneSided', 'kstwobign': 'KSTwoSided', 'laplace': 'Laplace',
    'levy': 'Levy', 'levy_l': 'LevyLeft', 'levy_stable': 'LevyStable',
    'loggamma': 'LogGamma', 'logistic': 'Logistic', 'loglaplace': 'LogLaplace',
    'lognorm': 'LogNormal', 'logser': 'LogSeries', 'lomax': 'Lomax',
    'maxwell': 'Maxwell', 'mielke': 'Mielke', 'nakagami': 'Nakagami',
    'nbinom': 'NegativeBinomial', 'ncf': 'NonCentralF', 'nct': 'NonCentralT',
    'ncx2': 'NonCentralChiSquared', 'norm': 'Normal', 'pareto': 'Pareto',
    'pearson3': 'PearsonType3', 'planck': 'Planck', 'poisson': 'Poisson',
    'powerlaw': 'PowerLaw', 'powerlognorm': 'PowerLogNormal',
    'powernorm': 'PowerNormal', 'randint': 'DiscreteUniform',
    'rayleigh': 'Rayleigh', 'rdist': 'RDistributed', 'recipinvgauss':
    'ReciprocalInverseGaussian', 'reciprocal': 'Reciprocal', 'rice': 'Rice',
    'semicircular': 'Semicircular', 'skellam': 'Skellam', 'skewnorm': 'SkewNormal',
    't': 'T', 'trapz': 'Trapezoidal', 'triang': 'Triangular',
    'truncexpon': 'TruncatedExponential', 'truncnorm': 'TruncatedNormal',
    'tukeylambda': 'TukeyLambda', 'uniform': 'Uniform', 'vonmises': 'VonMises',
    'vonmises_line': 'VonMisesLine', 'wald': 'Wald', 'weibull_max': 'WeibullMaximum',
    'weibull_min': 'WeibullMinimum', 'wrapcauchy': 'WrappedCauchy', 'zipf': 'Zipf',
}
_RENAMED = {'gilbrat': 'gibrat', 'trapz': 'trapezoid'}


class Distribution(object):
    """Base class of all priors (subclass it for a custom prior: provide
    ``random``, ``logp`` and ``median``)."""
    discrete = False

    def __init__(self):
        self.name = ''
        self.fitsname = ''
        self._value = self.random()

    def random(self):
        return 0

    def logp(self, x):
        return 0

    def median(self):
        return 0

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        self._value = val


class ScipyDistribution(Distribution):
    """A frozen scipy.stats distribution as a prior. Discrete distributions round
    the assigned value half-to-even to an integer; one-element arrays become
    Python scalars (cf. psfMC/distributions.py:130-138)."""
    scipy_name = None

    def __init__(self, *args, **kwargs):
        generator = getattr(_stats, self.scipy_name, None)
        if generator is None:
            generator = getattr(_stats, _RENAMED[self.scipy_name])
        self.rv_frozen = generator(*args, **kwargs)
        self.discrete = isinstance(self.rv_frozen.dist, _stats.rv_discrete)
        if not self.discrete and \
                not isinstance(self.rv_frozen.dist, _stats.rv_continuous):
            raise TypeError('Only rv_continuous and rv_discrete distributions '
                            'are supported')
        self.logp = self.rv_frozen.logpmf if self.discrete \
            else self.rv_frozen.logpdf
        self.random = self.rv_frozen.rvs
        self.median = self.rv_frozen.median
        super(ScipyDistribution, self).__init__()

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        if self.discrete:
            val = np.rint(val).astype(int)
        arr = np.asarray(val)
        self._value = arr.item() if arr.size == 1 else val

    # -- batched evaluation ----------------------------------------------------
    def _logpdf_direct(self, x):
        """rv_continuous.logpdf restated without its per-call bookkeeping
        (argsreduce / place / broadcast_arrays cost ~100 us per call, more than the
        arithmetic of a whole column of walkers): the same operations in the same
        order -- standardise, the distribution's own ``_logpdf``, minus log(scale),
        -inf outside the support, the bad value for invalid arguments or NaN."""
        rv = self.rv_frozen
        dist = rv.dist
        args, loc, scale = dist._parse_args(*rv.args, **rv.kwds)
        args = tuple(np.asarray(arg) for arg in args)
        loc, scale = np.asarray(loc), np.asarray(scale)
        dtyp = np.promote_types(np.asarray(x).dtype, np.float64)
        std = np.asarray((x - loc) / scale, dtype=dtyp)
        cond0 = dist._argcheck(*args) & (scale > 0)
        cond1 = dist._support_mask(std, *args) & (scale > 0)
        cond = cond0 & cond1
        values = dist._logpdf(std, *args) - np.log(scale)
        out = np.where(cond, values, -np.inf)
        bad = (1 - cond0) + np.isnan(std)
        if np.any(bad):
            out = np.where(bad, dist.badvalue, out)
        return out

    def direct_spec(self, length):
        """(scipy distribution, shape arguments, loc, scale), each broadcast to
        (length,), for evaluating several priors of one family in ONE array
        operation (psfmc_b200/models.py); None if this prior cannot take part."""
        if self.discrete or getattr(self, '_direct_state', None) is False:
            return None
        try:
            rv = self.rv_frozen
            args, loc, scale = rv.dist._parse_args(*rv.args, **rv.kwds)
            cast = [np.broadcast_to(np.asarray(v, dtype=np.float64), (length,)).copy()
                    for v in tuple(args) + (loc, scale)]
        except Exception:
            return None
        return rv.dist, cast[:-2], cast[-2], cast[-1]

    def logp_batch(self, x):
        """``logp`` of every element of the (B, length) array ``x``. Continuous
        distributions take the direct path above once it has reproduced
        ``rv_frozen.logpdf`` BIT FOR BIT on the first batch it sees (checked per
        prior object); anything else goes through scipy's generic entry point."""
        if self.discrete:
            return self.logp(x)
        state = getattr(self, '_direct_state', None)
        if state is False:
            return self.logp(x)
        try:
            with np.errstate(all='ignore'):
                fast = self._logpdf_direct(x)
        except Exception:
            self._direct_state = False
            return self.logp(x)
        if state is None:
            with np.errstate(all='ignore'):
                slow = np.asarray(self.logp(x))
            same = fast.shape == slow.shape and np.array_equal(fast, slow, equal_nan=True)
            self._direct_state = bool(same)
            return slow
        return fast


def _make_class(alias, scipy_name):
    doc = '{} prior: scipy.stats.{} with the same arguments.'.format(alias, scipy_name)
    return type(alias, (ScipyDistribution,),
                {'scipy_name': scipy_name, '__doc__': doc})


__all__ = ['Distribution']
for _scipy_name, _alias in _ALIASES.items():
    globals()[_alias] = _make_class(_alias, _scipy_name)
    __all__.append(_alias)
