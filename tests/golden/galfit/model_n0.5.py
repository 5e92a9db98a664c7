# C2: single Sersic on the reference's GALFIT fixture gfsim_n0.5.fits.gz
# (parameters from its header as in tests/test_components.py:63-74:
# xy = GALFIT - 1, reff_b = RE * AR, PA in degrees); delta PSF,
# constant weight, no mask. reff_b and angle are fixed constants.
from numpy import array
Configuration(obs_file='gfsim_n0.5.fits.gz', obsivm_file='ivm_const.fits',
              psf_files='psf_delta.fits',
              psfivm_files='psfivm_delta.fits',
              mag_zeropoint=26.2303)
Sersic(xy=Uniform(loc=array((63.5, 63.5)) - 2, scale=array((4, 4))),
       mag=Uniform(loc=21.72 - 1, scale=2),
       reff=Uniform(loc=6.3 - 2, scale=4), reff_b=5.1659999999999995,
       index=Uniform(loc=0.3, scale=8), angle=25.35,
       angle_degrees=True)
