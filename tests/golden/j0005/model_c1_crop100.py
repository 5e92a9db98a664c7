# C1 (model_c1.py) on the J0005-0006 frames cropped to rows 14:114,
# columns 14:114 -- a frame that is not a power of two; the mask is
# the cropped region mask as a FITS image (nonzero = excluded).
from numpy import array

qso_mag = 20.66
qso_xy, qso_box = array((64.5 - 14, 64.5 - 14)), array((8, 8))
blob_xy, blob_box = array((46 - 14, 85.6 - 14)), array((5, 5))

Configuration(obs_file='sci_crop100.fits', obsivm_file='ivm_crop100.fits',
              psf_files='sci_psf.fits', psfivm_files='ivm_psf.fits',
              mask_file='mask_crop100.fits', mag_zeropoint=25.9463)
Sky(adu=Normal(loc=0, scale=0.01))
PointSource(xy=Uniform(loc=qso_xy - qso_box, scale=2 * qso_box),
            mag=Uniform(loc=qso_mag - 0.2, scale=0.2 + 1.5))
Sersic(xy=Uniform(loc=qso_xy - qso_box, scale=2 * qso_box),
       mag=Uniform(loc=qso_mag, scale=27.5 - qso_mag),
       reff=Uniform(loc=2.0, scale=10.0), reff_b=Uniform(loc=2.0, scale=10.0),
       index=WeibullMinimum(c=1.5, scale=4),
       angle=Uniform(loc=0, scale=180), angle_degrees=True)
Sersic(xy=Uniform(loc=blob_xy - blob_box, scale=2 * blob_box),
       mag=Uniform(loc=23.5, scale=2.0),
       reff=Uniform(loc=2.0, scale=6.0), reff_b=Uniform(loc=2.0, scale=6.0),
       index=WeibullMinimum(c=1.5, scale=4),
       angle=Uniform(loc=0, scale=180), angle_degrees=True)
