"""
ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy) restatement of the reference's likelihood hot path
(SURVEY.md section 8a): render the parametric model -> FFT-convolve with the PSF
(+ PSF-variance term) -> masked, IVM-weighted chi-square -> lnL, plus the one-time
setup that produces the constant arrays. Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module; the product path (psfmc_b200) never does
and fails loudly when its CUDA library is missing.

Parity pinning: this restatement is checked bit-for-bit against the unmodified
reference imported from /root/reference (through oracle/refshim.py) by
tests/golden/make_golden.py, which also froze the golden vectors under
tests/golden/. What is *unpinned*: the reference has no tests for this path
(SURVEY.md section 4), and the third-party arithmetic it calls (numpy.fft pocketfft,
scipy.special.gammaincinv/gamma) runs here at numpy 2.3.5 / scipy 1.18.1 rather
than the reference's pinned numpy 1.21.5 / scipy 1.7.3 (environment.yml:71,93);
the ds9-region mask was evaluated by a pyregion stand-in.

Every function cites the reference lines it follows (paths relative to
/root/reference). Numpy calls are deliberately the same calls the reference makes
(np.dot for the 2x2 rotation, np.sum pairwise reduction, in-place ``+=`` into the
storage dtype) so that results agree to the last bit in every precision mode.

Precision modes, selected by the dtype of the arrays handed to
:class:`OracleModel` and by ``fft_upcast``:
  M1  float32 arrays, fft_upcast=False  (numpy >= 2: complex64 FFT, float32 tail)
  M2  float32 arrays, fft_upcast=True   (numpy 1.x: FFT and tail in float64)
  M3  float64 obs arrays, fft_upcast=True (what the reference computes from
      float64 inputs; the PSF stays in its file dtype until it is transformed)
"""
from math import fsum

import numpy as np
from scipy.special import gamma, gammaincinv

SKY, POINT, SERSIC = 'sky', 'point', 'sersic'

# parameter names per component kind, in the order the engine's C-ABI uses
PARAM_NAMES = {
    SKY: ('adu',),
    POINT: ('x', 'y', 'mag'),
    SERSIC: ('x', 'y', 'mag', 'reff', 'reff_b', 'index', 'angle'),
}


# ----------------------------------------------------------------- setup --

def preprocess_obs(obs_data, obs_ivm, exclude_px=None):
    """psfMC/utils.py:54-79 -- bad-pixel mask and variance map of the observation."""
    with np.errstate(divide='ignore', invalid='ignore'):
        badpx = ~np.isfinite(obs_data) | ~np.isfinite(obs_ivm) | (obs_ivm <= 0)
        obs_var = np.where(badpx, np.inf, 1 / obs_ivm)
    if exclude_px is not None:
        badpx = badpx | exclude_px.astype(bool)
    return obs_data, obs_var, badpx


def norm_psf(psf_data, psf_ivm):
    """psfMC/utils.py:45-51 -- fsum normalisation; IVM scaled by sum**2."""
    psf_sum = fsum(psf_data.flat)
    return psf_data / psf_sum, psf_ivm * psf_sum ** 2


def preprocess_psf(psf_data, psf_ivm):
    """psfMC/utils.py:106-123 -- zero bad pixels, normalise, variance map."""
    psf_data = np.array(psf_data)
    psf_ivm = np.array(psf_ivm)
    badpx = ~np.isfinite(psf_data) | ~np.isfinite(psf_ivm) | (psf_ivm <= 0)
    psf_data[badpx] = 0
    psf_ivm[badpx] = 0
    psf_data, psf_ivm = norm_psf(psf_data, psf_ivm)
    with np.errstate(divide='ignore'):
        psf_var = np.where(psf_ivm <= 0, 0, 1 / psf_ivm)
    return psf_data, psf_var


def calculate_psf_variability(psf_data, psf_vars):
    """psfMC/utils.py:136-157 -- add the inter-PSF variance to every map."""
    if len(psf_data) == 1:
        return list(psf_data), list(psf_vars)
    mismatch_var = np.var(psf_data, axis=0)
    return list(psf_data), [var + mismatch_var for var in psf_vars]


def pad_and_rfft_image(img, newshape, fft_upcast):
    """psfMC/utils.py:9-22 -- zero-pad at offset pad//2 and rfft2."""
    pad = np.asarray(newshape) - np.asarray(img.shape)
    if np.any(pad < 0):
        raise NotImplementedError('PSF images larger than observation images '
                                  'are not yet supported')
    img_pad = np.zeros(newshape, dtype=img.dtype)
    img_pad[pad[0] // 2:pad[0] // 2 + img.shape[0],
            pad[1] // 2:pad[1] // 2 + img.shape[1]] = img
    if fft_upcast:
        img_pad = img_pad.astype(np.float64)
    return np.fft.rfft2(img_pad)


def array_coords(shape):
    """psfMC/utils.py:35-42 -- (H*W, 2) float64 [x, y] per pixel, row-major."""
    indexes = np.arange(np.prod(shape))
    coords = [indexes % shape[1], indexes // shape[1]]
    return np.transpose(coords).astype('float64')


def mag_to_flux(mag, mag_zp):
    """psfMC/utils.py:160-164."""
    return 10 ** (-0.4 * (mag - mag_zp))


# ------------------------------------------------------------ components --

def sersic_kappa(index):
    """psfMC/ModelComponents/Sersic.py:47-53."""
    return gammaincinv(2 * index, 0.5)


def sersic_sb_eff(flux_tot, index, reff, reff_b, kappa):
    """psfMC/ModelComponents/Sersic.py:55-71."""
    return flux_tot / (np.pi * reff * reff_b * 2 * index *
                       np.exp(kappa + np.log(kappa) * -2 * index) *
                       gamma(2 * index))


def sersic_sq_radii(coords, xy, reff, reff_b, angle, angle_degrees):
    """psfMC/ModelComponents/Sersic.py:73-96 -- generalised-ellipse square radii
    and the per-pixel normalisation ``sq_radii / |offset|**2``."""
    angle = np.deg2rad(angle) if angle_degrees else angle
    angle += 0.5 * np.pi
    sin_ang, cos_ang = np.sin(angle), np.cos(angle)
    inv_xform = np.asarray((
        (cos_ang / reff, sin_ang / reff),
        (-sin_ang / reff_b, cos_ang / reff_b)
    ))
    coord_offsets = (coords - xy).T
    sq_radii = np.sum(np.dot(inv_xform, coord_offsets) ** 2, axis=0)
    sq_delta_r = sq_radii / np.sum(coord_offsets ** 2, axis=0)
    return sq_radii, sq_delta_r


def sersic_add_to_array(arr, mag_zp, coords, xy, mag, reff, reff_b, index,
                        angle, angle_degrees):
    """psfMC/ModelComponents/Sersic.py:98-134 and :136-153 (numpy branches; the
    numexpr branches evaluate the same two expressions)."""
    kappa = sersic_kappa(index)
    flux_tot = mag_to_flux(mag, mag_zp)
    sbeff = sersic_sb_eff(flux_tot, index, reff, reff_b, kappa)
    sq_radii, sq_delta_r = sersic_sq_radii(coords, xy, reff, reff_b, angle,
                                           angle_degrees)
    sq_radii = sq_radii.reshape(arr.shape)
    sq_delta_r = sq_delta_r.reshape(arr.shape)
    radius_pow = 0.5 / index
    sb = np.exp(-kappa * np.expm1(np.log(sq_radii) * radius_pow))
    normed_grad = -kappa * 2 * radius_pow * np.exp(
        np.log(sq_radii) * (radius_pow - 0.5))
    cent_offset = sq_delta_r / 12 * normed_grad
    arr += sbeff * sb * (1 + normed_grad * cent_offset)
    return arr


def _sinc(x):
    """psfMC/ModelComponents/PointSource.py:84-88."""
    return np.where(x != 0, np.sin(np.pi * x) / (np.pi * x), 1.0)


def _lanczos(x, a):
    """psfMC/ModelComponents/PointSource.py:91-97."""
    return np.where(np.abs(x) < a, _sinc(x) * _sinc(x / a), 0)


def minimal_slice(position, kern_radius, array_shape):
    """psfMC/ModelComponents/PointSource.py:60-81 -- clip, then round half to even."""
    kern_radius = np.array(kern_radius)
    array_shape = np.array(array_shape)
    clipped_pos = np.clip(position[::-1], kern_radius - 0.5,
                          array_shape - (kern_radius + 0.5))
    min_pos = np.round(clipped_pos - kern_radius).astype(int)
    max_pos = np.round(clipped_pos + kern_radius).astype(int)
    return (slice(min_pos[0], max_pos[0] + 1), slice(min_pos[1], max_pos[1] + 1))


def point_add_to_array(arr, mag_zp, coords, xy, mag, shift_method='lanczos3'):
    """psfMC/ModelComponents/PointSource.py:24-57."""
    coords_2d = coords.view()
    coords_2d.shape = arr.shape + (2,)
    xy = np.asarray(xy, dtype=np.float64)
    if shift_method == 'bilinear':
        kern_slice = minimal_slice(xy, 0.5, arr.shape)
        diffs = coords_2d[kern_slice] - xy
        kern = np.prod(1 - np.abs(diffs), axis=-1)
    elif shift_method == 'lanczos3':
        kern_slice = minimal_slice(xy, 3, arr.shape)
        diffs = coords_2d[kern_slice] - xy
        kern = np.prod(_lanczos(diffs, 3), axis=-1)
    else:
        raise ValueError('Unknown shift method: {}'.format(shift_method))
    flux = mag_to_flux(mag, mag_zp)
    arr[kern_slice] += kern * flux
    return arr


def sky_add_to_array(arr, adu):
    """psfMC/ModelComponents/Sky.py:14-16."""
    arr += adu
    return arr


# ---------------------------------------------------------------- model --

def discrete_value(val):
    """psfMC/distributions.py:130-138 -- discrete priors are rint-ed (half to
    even) to int before use."""
    return int(np.rint(val))


class OracleModel(object):
    """
    The reference's MultiComponentModel hot path over plain arrays.

    :param obs_data, obs_var, bad_px: arrays as produced by :func:`preprocess_obs`
    :param psf_list, psfvar_list: normalised PSFs and variance maps (after
        :func:`preprocess_psf` and :func:`calculate_psf_variability`)
    :param mag_zp: magnitude zeropoint
    :param program: list of ``(kind, flags, slots)``; ``slots`` maps each name of
        ``PARAM_NAMES[kind]`` to ``('theta', index)`` or ``('const', value)``;
        ``flags`` may hold ``angle_degrees`` / ``shift_method``
    :param psf_index_slot: slot for the PSF index (``('const', 0)`` for one PSF)
    :param fft_upcast: compute FFTs in float64 regardless of storage dtype
    """

    def __init__(self, obs_data, obs_var, bad_px, psf_list, psfvar_list, mag_zp,
                 program, psf_index_slot=('const', 0), fft_upcast=True):
        self.obs_data = obs_data
        self.obs_var = obs_var
        self.bad_px = np.asarray(bad_px, dtype=bool)
        self.mag_zp = mag_zp
        self.program = program
        self.psf_index_slot = psf_index_slot
        self.fft_upcast = fft_upcast
        shape = obs_data.shape
        # psfMC/ModelComponents/PSFSelector.py:39-43
        self.f_psf = [pad_and_rfft_image(psf, shape, fft_upcast)
                      for psf in psf_list]
        self.f_var = [pad_and_rfft_image(var, shape, fft_upcast)
                      for var in psfvar_list]
        # psfMC/ModelComponents/Configuration.py:52
        self.coords = array_coords(shape)

    @staticmethod
    def _slot(slot, theta):
        # the reference hands components Python scalars
        # (psfMC/distributions.py:135-138: asscalar)
        return float(theta[slot[1]]) if slot[0] == 'theta' else slot[1]

    def convolve(self, img, fourier_kernel):
        """psfMC/utils.py:25-32."""
        if self.fft_upcast:
            img = np.asarray(img, dtype=np.float64)
        return np.fft.ifftshift(np.fft.irfft2(np.fft.rfft2(img) * fourier_kernel))

    def raw_model(self, theta, only_point_sources=False):
        """psfMC/models.py:245-253 (and :301-305 for the point-source-only image)."""
        arr = np.zeros_like(self.obs_var)
        for kind, flags, slots in self.program:
            val = {name: self._slot(slots[name], theta)
                   for name in PARAM_NAMES[kind]}
            if only_point_sources and kind != POINT:
                continue
            if kind == SKY:
                sky_add_to_array(arr, val['adu'])
            elif kind == POINT:
                point_add_to_array(arr, self.mag_zp, self.coords,
                                   np.array([val['x'], val['y']]), val['mag'],
                                   flags.get('shift_method', 'lanczos3'))
            elif kind == SERSIC:
                sersic_add_to_array(arr, self.mag_zp, self.coords,
                                    np.array([val['x'], val['y']]), val['mag'],
                                    val['reff'], val['reff_b'], val['index'],
                                    val['angle'],
                                    flags.get('angle_degrees', False))
            else:
                raise ValueError('unknown component kind ' + str(kind))
        return arr

    def psf_index(self, theta):
        if self.psf_index_slot[0] == 'const':
            return int(self.psf_index_slot[1])
        return discrete_value(theta[self.psf_index_slot[1]])

    def images(self, theta, with_point_source_subtracted=True):
        """The five blob images of psfMC/models.py:213-226."""
        theta = np.asarray(theta, dtype=np.float64)
        kpsf = self.psf_index(theta)
        with np.errstate(all='ignore'):
            raw_px = self.raw_model(theta)
            conv_px = self.convolve(raw_px, self.f_psf[kpsf])          # :255-263
            resid_px = self.obs_data - conv_px                         # :282-294
            model_var = self.convolve(raw_px ** 2, self.f_var[kpsf])   # :265-280
            ivm_px = 1 / (model_var + self.obs_var)
            out = {'raw_model': raw_px, 'convolved_model': conv_px,
                   'residual': resid_px, 'composite_ivm': ivm_px}
            if with_point_source_subtracted:                           # :296-306
                ps_px = self.raw_model(theta, only_point_sources=True)
                ps_px = self.convolve(ps_px, self.f_psf[kpsf])
                out['point_source_subtracted'] = self.obs_data - ps_px
        return out

    def lnlike(self, theta):
        """psfMC/models.py:233-241 -- masked Normal lnL; non-finite => -inf."""
        imgs = self.images(theta, with_point_source_subtracted=False)
        good = ~self.bad_px
        with np.errstate(all='ignore'):
            ivm_flat = imgs['composite_ivm'][good]
            resid_flat = imgs['residual'][good]
            lnl = -0.5 * np.sum(resid_flat ** 2 * ivm_flat
                                - np.log(0.5 / np.pi * ivm_flat))
        lnl = float(lnl)
        return lnl if np.isfinite(lnl) else float('-inf')

    def lnlike_batch(self, thetas):
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        return np.array([self.lnlike(row) for row in thetas])


def build_from_raw_inputs(obs_data, obs_ivm, exclude_px, psfs, psf_ivms, mag_zp,
                          program, psf_index_slot=None, fft_upcast=True):
    """Run the reference's whole setup (Configuration.py:38-52,
    PSFSelector.py:16-43) on raw arrays and return an OracleModel."""
    obs_data, obs_var, bad_px = preprocess_obs(obs_data, obs_ivm, exclude_px)
    pairs = [preprocess_psf(psf, ivm) for psf, ivm in zip(psfs, psf_ivms)]
    psf_list, var_list = calculate_psf_variability(*zip(*pairs))
    if psf_index_slot is None:
        psf_index_slot = ('const', 0)
    return OracleModel(obs_data, obs_var, bad_px, psf_list, var_list, mag_zp,
                       program, psf_index_slot, fft_upcast)
