# C1 with two PSFs (=> PSF_Index becomes the last free parameter, D = 19, and the
# inter-PSF variance enters both variance maps; psfMC/utils.py:136-157), a
# bilinear point source and a Sersic with one fixed parameter and radians.
from numpy import array

Configuration(obs_file='sci_J0005-0006.fits', obsivm_file='ivm_J0005-0006.fits',
              psf_files=['sci_psf.fits', 'sci_psf_b.fits'],
              psfivm_files=['ivm_psf.fits', 'ivm_psf_b.fits'],
              mask_file='mask_J0005-0006.reg', mag_zeropoint=25.9463)
Sky(adu=Normal(loc=0, scale=0.01))
PointSource(xy=Uniform(loc=array((56.5, 56.5)), scale=array((16, 16))),
            mag=Uniform(loc=20.4, scale=1.7), shift_method='bilinear')
Sersic(xy=Uniform(loc=array((56.5, 56.5)), scale=array((16, 16))),
       mag=Uniform(loc=20.66, scale=6.84),
       reff=Uniform(loc=2.0, scale=10.0), reff_b=Uniform(loc=2.0, scale=10.0),
       index=1.0, angle=Uniform(loc=0, scale=3.141592653589793))
Sersic(xy=Uniform(loc=array((41, 80.6)), scale=array((10, 10))),
       mag=Uniform(loc=23.5, scale=2.0),
       reff=Uniform(loc=2.0, scale=6.0), reff_b=Uniform(loc=2.0, scale=6.0),
       index=WeibullMinimum(c=1.5, scale=4),
       angle=Uniform(loc=0, scale=180), angle_degrees=True)
