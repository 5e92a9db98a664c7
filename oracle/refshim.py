"""
TEST INFRASTRUCTURE ONLY -- loader that imports the *unmodified* reference
(/root/reference/psfMC) in this container so that the oracle restatement can be
pinned against it and golden vectors can be generated (tests/golden/make_golden.py).

The reference does not import as-is here: astropy, emcee,
numexpr, pyregion, matplotlib and corner are not installed, and numpy 2 / scipy
1.18 removed four names it uses. This module installs
  * API aliases: np.product, np.asscalar, scipy.stats.gilbrat, scipy.stats.trapz
  * stub modules: matplotlib*, mpl_toolkits*, corner, emcee, astropy.table,
    astropy.wcs
  * ``astropy.io.fits`` -> psfmc_b200.fitsio (getdata/getheader/writeto; raises
    IOError on non-FITS input, which psfMC/utils.py:87-90 relies on)
  * ``pyregion`` -> a stand-in over psfmc_b200.regions (image-frame shapes)
and then imports ``psfMC`` from ``/root/reference`` -- or, where that does not exist
(the GPU box), from the verbatim copy ``oracle/make_ref.py`` placed under
``oracle/_ref`` (git-ignored). Nothing of the reference is modified. Importers: the CPU
test tier, tests/golden/make_golden.py, and the CPU arm of bench.py (``--impl
reference`` and the ``cpu_baseline`` leg) -- never the product, never the ``-m gpu``
tests or ``smoke()``.

Precision modes:
  M1  numpy-2 native: float32 storage, complex64 FFT, float32 reduce
  M2  numpy-1.x-faithful (the reference's pinned numpy 1.21.5): float32 raw-model
      storage, FFT / residual / IVM / sum in float64 (np.fft up-casts)
  M3  all-float64: obs_data / obs_var are float64 (=> float64 raw model);
      PSF normalisation stays in the file dtype, spectra are complex128
"""
import os
import sys
import types

import numpy as np

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    """Where the unmodified reference package lives: $PSFMC_REFERENCE_ROOT, else
    /root/reference (this container), else the verbatim copy oracle/make_ref.py
    placed under oracle/_ref (the GPU box; git-ignored)."""
    env = os.environ.get('PSFMC_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isdir('/root/reference/psfMC'):
        return '/root/reference'
    return os.path.join(_REPO_ROOT, 'oracle', '_ref')


REFERENCE_ROOT = _reference_root()

_state = {'mode': 'M1', 'loaded': False}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'psfMC'))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []
    sys.modules[name] = mod
    return mod


class _Anything(object):
    """Attribute sink for plotting libraries that are imported but never used."""

    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *args, **kwargs):
        return _Anything()


def _install_stubs():
    if _REPO_ROOT not in sys.path:
        sys.path.insert(0, _REPO_ROOT)
    from psfmc_b200 import fitsio, regions

    if not hasattr(np, 'product'):
        np.product = np.prod
    if not hasattr(np, 'asscalar'):
        np.asscalar = lambda arr: np.asarray(arr).item()
    import scipy.stats as stats
    if not hasattr(stats, 'gilbrat'):
        stats.gilbrat = stats.gibrat
    if not hasattr(stats, 'trapz'):
        stats.trapz = stats.trapezoid

    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker',
                 'matplotlib.transforms', 'matplotlib.patheffects',
                 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.gridspec',
                 'mpl_toolkits', 'mpl_toolkits.axes_grid1', 'corner'):
        if name not in sys.modules:
            mod = _stub(name)
            mod.__getattr__ = lambda attr: _Anything()
    if 'emcee' not in sys.modules:
        emcee = _stub('emcee', EnsembleSampler=_Anything)
        emcee.autocorr = _stub('emcee.autocorr', AutocorrError=Exception,
                               integrated_time=_Anything())

    def _getdata(source, *args, **kwargs):
        data = fitsio.getdata(source)
        if _state['mode'] == 'M3' and _state.get('upcast_next', 0) > 0:
            _state['upcast_next'] -= 1
            data = data.astype(np.float64)
        return data

    astropy = _stub('astropy')
    astropy.io = _stub('astropy.io')
    astropy.io.fits = _stub('astropy.io.fits', getdata=_getdata,
                            getheader=fitsio.getheader, writeto=fitsio.writeto,
                            Header=fitsio.Header)
    astropy.table = _stub('astropy.table', Table=_Anything)
    astropy.wcs = _stub('astropy.wcs', WCS=_Anything)
    astropy.wcs.utils = _stub('astropy.wcs.utils',
                              proj_plane_pixel_area=_Anything())

    class _Filter(object):
        def __init__(self, text):
            self.text = text

        def mask(self, shape):
            return regions.region_mask(self.text, shape)

    class _ShapeList(object):
        def __init__(self, text):
            self.text = text

        def as_imagecoord(self, header):
            return self

        def get_filter(self):
            return _Filter(self.text)

    def _region_open(filename):
        # a FITS file handed to pyregion fails to decode as text
        with open(filename, 'r') as fobj:
            return _ShapeList(fobj.read())

    _stub('pyregion', open=_region_open)


def load_reference():
    """Import and return the reference ``psfMC`` package."""
    if not reference_available():
        raise ImportError('reference not present at ' + REFERENCE_ROOT)
    if not _state['loaded']:
        _install_stubs()
        # psfmc_b200.model_parser registers a stand-in ``psfMC`` module (so that model
        # files importing from psfMC keep working where the package is not installed):
        # drop it, the real package is wanted here
        for name in [n for n in sys.modules if n == 'psfMC' or n.startswith('psfMC.')]:
            origin = getattr(sys.modules[name], '__file__', None) or ''
            if not origin.startswith(REFERENCE_ROOT):
                del sys.modules[name]
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            import psfMC  # noqa: F401
        _state['loaded'] = True
    return sys.modules['psfMC']


class _UpcastFFT(object):
    """np.fft look-alike whose rfft2/irfft2 compute in float64/complex128, the
    behaviour of numpy < 2 (pocketfft always double) that the reference's pinned
    environment had (environment.yml:71)."""

    @staticmethod
    def rfft2(arr, *args, **kwargs):
        return np.fft.rfft2(np.asarray(arr, dtype=np.float64), *args, **kwargs)

    @staticmethod
    def irfft2(arr, *args, **kwargs):
        return np.fft.irfft2(np.asarray(arr, dtype=np.complex128), *args, **kwargs)

    ifftshift = staticmethod(np.fft.ifftshift)
    fftshift = staticmethod(np.fft.fftshift)


class _NumpyWithFFT(object):
    """Proxy for the ``np`` name inside psfMC.utils with ``.fft`` replaced."""

    def __init__(self, fft):
        self.fft = fft

    def __getattr__(self, name):
        return getattr(np, name)


def build_reference_model(model_file, mode='M1'):
    """
    Build the reference's MultiComponentModel for ``model_file`` under precision
    mode M1/M2/M3 and return it. The mode stays in force (module-global patch of
    psfMC.utils.np) until the next call.
    """
    psfMC = load_reference()
    import psfMC.utils as ref_utils
    _state['mode'] = mode
    if mode in ('M2', 'M3'):
        ref_utils.np = _NumpyWithFFT(_UpcastFFT)
    else:
        ref_utils.np = np
    # In M3 the first two getdata calls of preprocess_obs (obs, ivm;
    # psfMC/utils.py:61-62) return float64; PSF reads stay in file dtype.
    _state['upcast_next'] = 2 if mode == 'M3' else 0
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = psfMC.MultiComponentModel(components=os.path.abspath(model_file))
    _state['upcast_next'] = 0
    return model


def reference_lnlike(model, theta):
    """
    (lnL, lnprior) from the reference's own log_posterior
    (psfMC/models.py:193-243) for one parameter vector.
    """
    theta = np.asarray(theta, dtype=np.float64)
    with np.errstate(all='ignore'):
        model.param_values = theta
        lnprior = float(model.log_priors())
        lnpost, blobs = type(model).log_posterior(theta, model=model)
    lnpost = float(lnpost)
    if not np.isfinite(lnprior):
        return float('-inf'), lnprior, blobs
    if not np.isfinite(lnpost):
        return float('-inf'), lnprior, blobs
    # lnpost = lnL + lnprior was formed in floating point; recompute lnL from
    # the blobs exactly as models.py:233-236 does so that no cancellation enters.
    good = ~model.config.bad_px
    ivm_flat = blobs['composite_ivm'][good]
    resid_flat = blobs['residual'][good]
    lnl = -0.5 * np.sum(resid_flat ** 2 * ivm_flat - np.log(0.5 / np.pi * ivm_flat))
    return float(lnl), lnprior, blobs
