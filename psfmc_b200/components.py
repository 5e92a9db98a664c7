"""
Model components with the reference's public interface (psfMC.ModelComponents):
``Configuration``, ``Sky``, ``PointSource``, ``Sersic`` (+ the internal
``PSFSelector``), so that psfMC model files run unchanged. Reference:
/root/reference/psfMC/ModelComponents/*.py.

What differs from the reference: components here do not render themselves with
numpy -- they only describe the model. :mod:`psfmc_b200.program` flattens the
component list into the engine's program (which parameter is a constant, which
is a slot of theta) and the CUDA engine renders whole walker batches. Priors are
evaluated column-vectorised over a batch (:meth:`ComponentBase.log_priors_batch`).
"""
import numpy as np

from . import preprocess
from .distributions import DiscreteUniform

__all__ = ['Configuration', 'Sky', 'PointSource', 'Sersic', 'PSFSelector',
           'ComponentBase', 'StochasticProperty']


class StochasticProperty(object):
    """Descriptor for a model parameter that is either a fixed value or a prior
    (cf. ComponentBase.py:132-153): reading gives the current value."""

    def __init__(self, key):
        self.key = key

    def __get__(self, instance, owner=None):
        if instance is None:
            return self
        return instance.get_stochastic_val(self.key)

    def __set__(self, instance, value):
        instance.assign_stochastic(self.key, value)

    def __delete__(self, instance):
        raise NotImplementedError('Cannot delete stochastics')


class ComponentBase(object):
    """
    Bookkeeping shared by all components. Free parameters are ordered
    alphabetically by attribute name inside a component, and a parameter with an
    n-vector value (``xy``) occupies n consecutive slots of theta
    (cf. ComponentBase.py:45-97).
    """
    _fits_abbrs = []

    def __init__(self):
        self._priors = {}
        self._constants = {}

    # -- storage ---------------------------------------------------------------
    def assign_stochastic(self, name, value):
        """A value with a ``.value`` attribute is a prior, anything else a constant."""
        if hasattr(value, 'value'):
            self._priors[name] = value
            self._constants.pop(name, None)
        else:
            self._constants[name] = value
            self._priors.pop(name, None)

    def get_stochastic_val(self, name):
        prior = self._priors.get(name)
        return prior.value if prior is not None else self._constants[name]

    def free_parameters(self):
        """[(attribute, prior, n_slots)] in canonical (sorted) order."""
        return [(name, self._priors[name], int(np.size(self._priors[name].value)))
                for name in sorted(self._priors)]

    # -- reference-compatible accessors ---------------------------------------
    def stochastic_lens(self):
        return [length for _, _, length in self.free_parameters()]

    def num_stochastics(self):
        return int(sum(self.stochastic_lens()))

    def stochastic_names(self, name_attr='name'):
        return [getattr(prior, name_attr) for _, prior, _ in self.free_parameters()]

    def get_distribution(self, stoch_name):
        found = [prior for prior in self._priors.values() if prior.name == stoch_name]
        if len(found) != 1:
            raise KeyError('Could not find unique prior with name: {}'
                           .format(stoch_name))
        return found[0]

    def set_stochastic_values(self, param_values='random'):
        """Assign all free parameters from a vector (canonical order), or draw
        them ('random') / take prior medians ('median'). Returns the vector."""
        free = self.free_parameters()
        if isinstance(param_values, str):
            drawn = [np.ravel(getattr(prior, param_values)()) for _, prior, _ in free]
            param_values = np.concatenate(drawn) if drawn else np.array([])
        start = 0
        for _, prior, length in free:
            prior.value = np.array(param_values[start:start + length])
            start += length
        return param_values

    def update_stochastic_names(self, count=None):
        """Trace names '<count>_<Class>_<attr>' and abbreviated FITS names
        (cf. ComponentBase.py:99-119)."""
        kind = type(self).__name__
        for attr, prior in self._priors.items():
            name = '{}_{}'.format(kind, attr)
            fitsname = name
            for longname, abbr in type(self)._fits_abbrs:
                fitsname = fitsname.replace(longname, abbr)
            if count is not None:
                name = '{:d}_{}'.format(count, name)
                fitsname = '{:d}{}'.format(count, fitsname)
            try:
                prior.name, prior.fitsname = name, fitsname
            except AttributeError:
                pass

    # -- priors ------------------------------------------------------------------
    def log_priors(self):
        """Joint log-prior at the current values (cf. ComponentBase.py:121-129)."""
        total = 0
        for prior in self._priors.values():
            total += np.sum(prior.logp(prior.value))
        return total

    def log_priors_batch(self, block, column_logp=None):
        """
        Joint log-prior for every row of ``block`` (B, num_stochastics()):
        one vectorised ``logp`` call per prior instead of B scalar calls.
        ``column_logp`` (same shape): per-column log-densities somebody already
        evaluated (the model groups priors of one family, models.py); only the
        summation -- in the same order -- is done here then.
        """
        block = np.asarray(block, dtype=np.float64)
        total = np.zeros(block.shape[0])
        for _, prior, start, length in self.prior_terms():
            if column_logp is not None:
                total = total + np.sum(column_logp[:, start:start + length], axis=1)
                continue
            cols = block[:, start:start + length]
            if getattr(prior, 'discrete', False):
                cols = np.rint(cols).astype(int)
            batched = getattr(prior, 'logp_batch', prior.logp)
            with np.errstate(all='ignore'):
                total = total + np.sum(batched(cols), axis=1)
        return total

    def prior_terms(self):
        """[(attribute, prior, first column, n_columns)] in the order the reference ADDS
        the priors up -- the insertion order of ``_priors`` (ComponentBase.py:121-129:
        ``for prior in self._priors.values()``) -- with the columns of the component's
        block, which is laid out in sorted attribute order (:meth:`free_parameters`)."""
        where, start = {}, 0
        for name, _, length in self.free_parameters():
            where[name] = (start, length)
            start += length
        return [(name, prior) + where[name] for name, prior in self._priors.items()]

    def column_of(self, attr, block, sub=0):
        """Per-walker values of parameter ``attr`` given the component's block."""
        start = 0
        for name, _, length in self.free_parameters():
            if name == attr:
                return block[:, start + sub]
            start += length
        return np.full(block.shape[0], np.ravel(self._constants[attr])[sub],
                       dtype=np.float64)


class Sky(ComponentBase):
    """Constant background, ADU per pixel (cf. Sky.py)."""
    adu = StochasticProperty('adu')

    def __init__(self, adu=None):
        super(Sky, self).__init__()
        self.adu = adu


class PointSource(ComponentBase):
    """Point source at 0-based pixel position ``xy`` with total magnitude ``mag``;
    ``shift_method`` is 'lanczos3' (default) or 'bilinear' (cf. PointSource.py)."""
    _fits_abbrs = [('PointSource', 'PS')]
    xy = StochasticProperty('xy')
    mag = StochasticProperty('mag')

    def __init__(self, xy=None, mag=None, shift_method='lanczos3'):
        super(PointSource, self).__init__()
        if shift_method not in ('lanczos3', 'bilinear'):
            raise ValueError('Unknown shift method: {}'.format(shift_method))
        self.xy = xy
        self.mag = mag
        self.shift_method = shift_method


class Sersic(ComponentBase):
    """Sersic profile (cf. Sersic.py): centre ``xy``, total ``mag``, semi-major /
    semi-minor effective radii ``reff`` / ``reff_b``, ``index`` n and position
    ``angle`` (CCW of up; radians unless ``angle_degrees``)."""
    _fits_abbrs = [('Sersic', 'SER'), ('reff_b', 'REB'), ('reff', 'RE'),
                   ('index', 'N'), ('angle', 'ANG')]
    xy = StochasticProperty('xy')
    mag = StochasticProperty('mag')
    reff = StochasticProperty('reff')
    reff_b = StochasticProperty('reff_b')
    index = StochasticProperty('index')
    angle = StochasticProperty('angle')

    def __init__(self, xy=None, mag=None, reff=None, reff_b=None, index=None,
                 angle=None, angle_degrees=False):
        super(Sersic, self).__init__()
        self.xy = xy
        self.mag = mag
        self.reff = reff
        self.reff_b = reff_b
        self.index = index
        self.angle = angle
        self.angle_degrees = angle_degrees

    def log_priors(self):
        # the semi-minor axis may not exceed the semi-major one (Sersic.py:41-45)
        logp = super(Sersic, self).log_priors()
        return logp + (-np.inf if self.reff_b > self.reff else 0)

    def log_priors_batch(self, block, column_logp=None):
        block = np.asarray(block, dtype=np.float64)
        logp = super(Sersic, self).log_priors_batch(block, column_logp)
        swapped = self.column_of('reff_b', block) > self.column_of('reff', block)
        return np.where(swapped, -np.inf, logp)


class PSFSelector(ComponentBase):
    """
    The PSFs of a model (cf. PSFSelector.py). With more than one PSF the index is
    a free parameter with a DiscreteUniform(0, K) prior, named 'PSF_Index'. The
    normalised real-space PSFs and variance maps are kept: the engine pads and
    transforms them on the device.
    """
    psf_index = StochasticProperty('psf_index')

    def __init__(self, psf_list, ivm_list, data_shape):
        super(PSFSelector, self).__init__()
        if isinstance(psf_list, (str, np.ndarray)):
            psf_list = [psf_list]
        if isinstance(ivm_list, (str, np.ndarray)):
            ivm_list = [ivm_list]
        if len(psf_list) != len(ivm_list):
            raise ValueError('PSF and IVM lists must be the same length')
        pairs = [preprocess.preprocess_psf(psf, ivm)
                 for psf, ivm in zip(psf_list, ivm_list)]
        psfs, variances = preprocess.add_psf_variability(*zip(*pairs))
        for psf in psfs:
            if psf.shape[0] > data_shape[0] or psf.shape[1] > data_shape[1]:
                raise NotImplementedError('PSF images larger than observation '
                                          'images are not yet supported')
        self.filenames = list(psf_list)
        self.psf_images = psfs
        self.var_images = variances
        self.data_shape = tuple(data_shape)
        self.psf_index = DiscreteUniform(low=0, high=len(psfs)) \
            if len(psfs) > 1 else 0

    def update_stochastic_names(self, count=None):
        if 'psf_index' in self._priors:
            self._priors['psf_index'].name = 'PSF_Index'
            self._priors['psf_index'].fitsname = 'PSF_IDX'

    @property
    def filename(self):
        return self.filenames[self.psf_index]


class Configuration(ComponentBase):
    """
    Input images and control parameters (cf. Configuration.py:10-52): observed
    image + inverse-variance map, PSF image(s) + their inverse-variance maps,
    optional mask (FITS, nonzero = excluded, or ds9 region file) and the magnitude
    zeropoint. Files may be names or arrays.
    """

    def __init__(self, obs_file, obsivm_file, psf_files, psfivm_files,
                 mask_file=None, mag_zeropoint=0):
        super(Configuration, self).__init__()
        self.mag_zeropoint = mag_zeropoint
        self.obs_header = None
        if isinstance(obs_file, str):
            from . import fitsio
            self.obs_header = fitsio.getheader(obs_file)
        self.obs_data, self.obs_var, self.bad_px = \
            preprocess.preprocess_obs(obs_file, obsivm_file, mask_file)
        if self.obs_header is None:
            from .fitsio import Header
            self.obs_header = Header()
        self.psf_selector = PSFSelector(psf_files, psfivm_files, self.obs_data.shape)
