# Quasar J0005-0006 (point source) + host galaxy (Sersic) + faint companion (Sersic)
# + sky: the components, priors and input files of psfMC's own example, written in
# psfMC's model-file syntax (component and prior names are injected by the parser,
# relative file names are resolved against this file's directory). The data files
# are the fixtures under tests/golden/j0005/.
from numpy import array

data = '../tests/golden/j0005/'
qso_mag = 20.66
qso_xy, qso_box = array((64.5, 64.5)), array((8, 8))
blob_xy, blob_box = array((46, 85.6)), array((5, 5))

Configuration(obs_file=data + 'sci_J0005-0006.fits',
              obsivm_file=data + 'ivm_J0005-0006.fits',
              psf_files=data + 'sci_psf.fits', psfivm_files=data + 'ivm_psf.fits',
              mask_file=data + 'mask_J0005-0006.reg', mag_zeropoint=25.9463)
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
