"""
psfmc_b200 -- B200-native batched likelihood engine for psfMC models.

The hot path (render -> FFT convolution with the PSF and its variance map ->
masked chi-square -> lnL) runs in hand-written sm_100a CUDA kernels behind a C ABI
(include/psfmc_b200.h); this package is the host side that mirrors the reference's
Python interface for that path.
"""
__version__ = '0.1.0'

from .engine import LikelihoodEngine, fp32_peak_tflops  # noqa: F401
from .models import MultiComponentModel                 # noqa: F401
from .pool import BatchPool                             # noqa: F401
from .sampler import EnsembleSampler                    # noqa: F401
from .fitting import model_galaxy_mcmc                  # noqa: F401
