"""
CPU tier: the oracle (oracle/psfmc_oracle.py, numpy restatement of the reference's
lnL hot path) against the committed golden vectors, which were produced by the
UNMODIFIED reference (tests/golden/make_golden.py imports /root/reference through
oracle/refshim.py and asserts bitwise equality before writing). When the reference
is present (this container; never the GPU box) a slice is re-checked live.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, j0005_arrays, load_golden

from oracle import psfmc_oracle as orc
from psfmc_b200 import fitsio
from psfmc_b200.components import Configuration
from psfmc_b200.model_parser import component_list_from_file
from psfmc_b200.program import compile_program

MODES = ('M1', 'M2', 'M3')


def _program(model_file):
    comps = component_list_from_file(os.path.join(GOLDEN, model_file))
    config = [c for c in comps if isinstance(c, Configuration)][0]
    rest = [c for c in comps if c is not config] + [config.psf_selector]
    program, psf_slot, ndim = compile_program(rest)
    return program, psf_slot, ndim, config


def _oracle_j0005(model_file, mode, two_psf=False):
    program, psf_slot, ndim, _ = _program(model_file)
    obs, ivm, mask, psfs, ivms = j0005_arrays(np.float32, two_psf)
    if mode == 'M3':
        obs, ivm = obs.astype(np.float64), ivm.astype(np.float64)
    psfs = [fitsio.getdata(p) for p in psfs]
    ivms = [fitsio.getdata(p) for p in ivms]
    return orc.build_from_raw_inputs(obs, ivm, mask, psfs, ivms, 25.9463, program,
                                     psf_slot, fft_upcast=(mode != 'M1')), ndim


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('mode', MODES)
def test_c1_oracle_reproduces_reference_bitwise(mode):
    golden = load_golden('c1_golden.json')
    oracle, ndim = _oracle_j0005('j0005/model_c1.py', mode)
    assert ndim == golden['setup']['num_params'] == 18
    # mask / pixel-index logic is bit-exact (north_star)
    good = np.flatnonzero(~oracle.bad_px)
    assert len(good) == golden['setup']['n_good'] == 9176
    assert int(good.sum()) == golden['setup']['good_index_sum']
    assert good[0] == golden['setup']['first_good']
    assert good[-1] == golden['setup']['last_good']
    thetas = np.array(golden['theta'])
    rows = range(len(thetas)) if mode == 'M3' else range(0, len(thetas), 3)
    px = np.array(golden['sample_px'])
    for row in rows:
        lnl = oracle.lnlike(thetas[row])
        assert lnl == golden['lnl'][mode][row], (mode, row)
        if row < 12:
            imgs = oracle.images(thetas[row], with_point_source_subtracted=False)
            for key, want in golden['pixels'][mode][row].items():
                assert _same(imgs[key].ravel()[px], want), (mode, row, key)


@pytest.mark.parametrize('mode', MODES)
def test_point_source_subtracted_oracle_reproduces_reference_bitwise(mode):
    """psfMC/models.py:296-306 (row a13): sampled pixels and sums of the reference's
    image, frozen by make_golden.py --pssub."""
    golden = load_golden('c1_pssub_golden.json')
    px = np.array(golden['sample_px'])
    for tag, two_psf in (('c1', False), ('c1_2psf', True)):
        case = golden['cases'][tag]
        oracle, _ = _oracle_j0005(case['model_file'], mode, two_psf)
        for row, theta in enumerate(case['theta']):
            img = oracle.images(np.array(theta))['point_source_subtracted']
            flat = np.asarray(img, dtype=np.float64).ravel()
            assert _same(flat[px], case['pixels'][mode][row]), (tag, mode, row)
            assert float(flat.sum()) == case['sum'][mode][row], (tag, mode, row)


def test_c1_named_edge_cases():
    golden = load_golden('c1_golden.json')
    names = golden['names']
    lnl = golden['lnl']
    centre = names.index('C_exact_centre')
    for mode in MODES:
        # Sersic centred exactly on a pixel: 0/0 -> NaN -> lnL = -inf
        assert lnl[mode][centre] == -np.inf
    # SURVEY.md section 8(c) quoted figures for vectors A and B (M3, 1e-9 relative:
    # they were probed in a separate session)
    assert abs(lnl['M3'][0] / -4162.983738741816 - 1) < 1e-7
    assert abs(lnl['M3'][1] / -23910.64617480046 - 1) < 1e-7
    assert abs(golden['lnprior'][0] - -37.19252878610715) < 1e-12


@pytest.mark.parametrize('mode', ('M2', 'M3'))
def test_c1_two_psf_oracle(mode):
    golden = load_golden('c1_2psf_golden.json')
    oracle, ndim = _oracle_j0005('j0005/model_c1_2psf.py', mode, two_psf=True)
    assert ndim == golden['setup']['num_params']
    thetas = np.array(golden['theta'])
    for row in range(0, len(thetas), 2 if mode == 'M2' else 1):
        assert oracle.lnlike(thetas[row]) == golden['lnl'][mode][row]
    # half-to-even rounding of the PSF index (psfMC/distributions.py:130-138)
    assert [orc.discrete_value(v) for v in (0.5, 1.5, -0.5, 1.49, 0.51)] == \
        [0, 2, 0, 1, 1]


@pytest.mark.parametrize('index', ['0.5', '1.0', '3.1', '4.0', '6.5'])
def test_c2_galfit_sweep_oracle(index):
    case = load_golden('c2_golden.json')['cases'][index]
    program, psf_slot, ndim, config = _program(case['model_file'])
    gdir = os.path.join(GOLDEN, 'galfit')
    obs = fitsio.getdata(os.path.join(
        gdir, 'gfsim_n{}.fits.gz'.format(index))).astype(np.float64)
    ivm = fitsio.getdata(os.path.join(gdir, 'ivm_const.fits')).astype(np.float64)
    oracle = orc.build_from_raw_inputs(
        obs, ivm, None, [fitsio.getdata(os.path.join(gdir, 'psf_delta.fits'))],
        [fitsio.getdata(os.path.join(gdir, 'psfivm_delta.fits'))],
        config.mag_zeropoint, program, psf_slot, fft_upcast=True)
    thetas = np.array(case['theta'])
    assert thetas.shape[1] == ndim == 5
    for row, theta in enumerate(thetas):
        assert oracle.lnlike(theta) == case['lnl']['M3'][row]
    # the reference test's own sanity bound: psfMC vs GALFIT agree to a few per cent
    # (tests/test_components.py:49-118 prints, never asserts)
    assert case['galfit_median_frac_err'] < 0.01 or index == '0.5'
    assert abs(case['galfit_flux_ratio'] - 1) < 0.02


def test_pointsource_known_answer():
    """The one assertion the reference's tests hold for this path: bilinear shift
    of a unit source equals scipy.ndimage.shift(order=1)
    (/root/reference/tests/test_components.py:121-144)."""
    golden = load_golden('pointsource_golden.json')
    arr = np.zeros((5, 5))
    orc.point_add_to_array(arr, 0.0, orc.array_coords((5, 5)),
                           np.array(golden['bilinear_xy']), 0.0, 'bilinear')
    assert np.allclose(arr, np.array(golden['scipy_shift_5x5']))
    assert _same(arr, golden['bilinear_5x5'])
    arr = np.zeros((16, 16))
    orc.point_add_to_array(arr, 0.0, orc.array_coords((16, 16)),
                           np.array(golden['lanczos_xy']), 0.0, 'lanczos3')
    assert _same(arr, golden['lanczos_16x16'])
    # not renormalised: the 6x6 taps sum to slightly less than one
    assert 0.98 < arr.sum() < 1.0 + 1e-12


def test_stamp_bounds_round_half_even_and_clip():
    # psfMC/ModelComponents/PointSource.py:60-81
    sl = orc.minimal_slice(np.array([64.5, 63.5]), 3, (128, 128))
    assert (sl[0].start, sl[0].stop, sl[1].start, sl[1].stop) == (60, 67, 62, 69)
    sl = orc.minimal_slice(np.array([1.5, 0.2]), 3, (128, 128))      # clipped low
    assert (sl[0].start, sl[1].start) == (0, 0)
    # clipped high: 124.5 + 3 = 127.5 rounds (half to even) to 128, one past the
    # last row; numpy truncates the slice, so the reference still works
    sl = orc.minimal_slice(np.array([126.5, 127.4]), 3, (128, 128))
    assert (sl[0].stop, sl[1].stop) == (129, 129)
    assert np.zeros((128, 128))[sl].shape == (6, 6)


_LIVE_CHECK = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import refshim
golden = json.load(open(os.path.join(sys.argv[1], 'tests/golden/c1_golden.json')))
model = refshim.build_reference_model(
    os.path.join(sys.argv[1], 'tests/golden/j0005/model_c1.py'), 'M3')
for row in (0, 1, 3, 9):
    theta = np.array(golden['theta'][row])
    lnpost, _ = type(model).log_posterior(theta, model=model)
    want = golden['lnl']['M3'][row] + golden['lnprior'][row]
    assert lnpost == (want if np.isfinite(want) else -np.inf), (row, lnpost, want)
print('LIVE-OK')
"""


def test_live_against_reference_when_present():
    """Re-run the unmodified reference's own log_posterior on golden vectors (in a
    fresh interpreter: the shim loader and this package's model-file parser both
    register a ``psfMC`` module)."""
    import subprocess
    import sys
    from conftest import ROOT
    if not os.path.isdir('/root/reference/psfMC'):
        pytest.skip('/root/reference not present (GPU box)')
    proc = subprocess.run([sys.executable, '-c', _LIVE_CHECK, ROOT],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          universal_newlines=True, timeout=300)
    assert proc.returncode == 0 and 'LIVE-OK' in proc.stdout, proc.stdout


@pytest.mark.parametrize('tag', ['crop100', 'crop75x100'])
@pytest.mark.parametrize('mode', MODES)
def test_cropped_frames_oracle_reproduces_reference_bitwise(tag, mode):
    """Frames that are not powers of two (100 x 100, 75 x 100): the reference convolves
    circularly at the image size; golden vectors from the unmodified reference."""
    case = load_golden('c1_cropped_golden.json')['cases'][tag]
    program, psf_slot, ndim, _ = _program(case['model_file'])
    jdir = os.path.join(GOLDEN, 'j0005')
    obs = fitsio.getdata(os.path.join(jdir, 'sci_{}.fits'.format(tag)))
    ivm = fitsio.getdata(os.path.join(jdir, 'ivm_{}.fits'.format(tag)))
    mask = fitsio.getdata(os.path.join(jdir, 'mask_{}.fits'.format(tag))) != 0
    assert list(obs.shape) == case['setup']['shape']
    if mode == 'M3':
        obs, ivm = obs.astype(np.float64), ivm.astype(np.float64)
    psfs = [fitsio.getdata(os.path.join(jdir, 'sci_psf.fits'))]
    ivms = [fitsio.getdata(os.path.join(jdir, 'ivm_psf.fits'))]
    oracle = orc.build_from_raw_inputs(obs, ivm, mask, psfs, ivms, 25.9463, program,
                                       psf_slot, fft_upcast=(mode != 'M1'))
    good = np.flatnonzero(~oracle.bad_px)
    assert len(good) == case['setup']['n_good']
    assert int(good.sum()) == case['setup']['good_index_sum']
    thetas = np.array(case['theta'])
    px = np.array(case['sample_px'])
    for row in range(len(thetas)):
        assert oracle.lnlike(thetas[row]) == case['lnl'][mode][row], (mode, row)
        if row < 3:
            imgs = oracle.images(thetas[row], with_point_source_subtracted=False)
            for key, want in case['pixels'][mode][row].items():
                assert _same(imgs[key].ravel()[px], want), (mode, row, key)


@pytest.mark.parametrize('mode', ['M2', 'M3'])
def test_oracle_tiny_sersic_index_vectors(mode):
    """Indices below 0.01: the reference's own -inf (overflow of its gradient term,
    Sersic.py:129-133), frozen by tests/golden/make_tiny_index_golden.py; the oracle
    reproduces them and the finite neighbours bit for bit."""
    data = load_golden('c1_tiny_index.json')
    thetas = np.array(data['theta'])
    want = np.array([-np.inf if v is None else v for v in data['lnl_' + mode]])
    assert np.sum(~np.isfinite(want)) == 4
    oracle, _ = _oracle_j0005('j0005/model_c1.py', mode)
    with np.errstate(all='ignore'):
        got = oracle.lnlike_batch(thetas)
    assert np.array_equal(got, want)


def test_oracle_equals_reference_bitwise_outside_the_priors():
    """tools/ref_fuzz.py: the oracle against the unmodified reference's own raw_model /
    convolved_model / residual / composite_ivm on parameter vectors far outside the priors
    (centres off the frame, reff 0.05 ... 300 px, index 0.05 ... 12, 60 000 ADU components,
    both PSFs, a cropped frame, a GALFIT fixture, Sersic centres on and next to pixel
    centres where a third of the walkers is NaN) -- images and lnL bit for bit in the three
    precision modes. The golden vectors pin prior draws and named edge cases;
    profiles/r2e_ref_fuzz.txt holds the 17408-vector run of this check."""
    import subprocess
    import sys
    from conftest import ROOT
    if not os.path.isdir('/root/reference/psfMC'):
        pytest.skip('/root/reference not present (GPU box)')
    for which, box in (('c1', 'wide'), ('c1_2psf', 'wide'), ('crop75x100', 'wide'),
                       ('c2_n4.0', 'wide'), ('c1', 'hot')):
        proc = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ref_fuzz.py'),
                               '64', '31', box, which], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
        lines = [line for line in proc.stdout.splitlines() if line.startswith(which)]
        assert proc.returncode == 0 and len(lines) == 3, proc.stdout
        for line in lines:
            assert 'lnL differing 0 ' in line, line
            assert line.count(': 0') == 5, line        # the five images
