"""
CPU tier: the CUDA kernel SOURCES (psfmc_b200/csrc/*.cu*) compiled by g++ against
the SIMT emulator (tests/emu/cuda_emu.h) and driven through the same C ABI and the
same host code as on the GPU, compared with the oracle / golden vectors. This
checks the kernels' arithmetic, indexing, barriers and launch plans without a GPU;
it says nothing about speed. The product library is never built this way.
"""
import numpy as np
import pytest

from conftest import (ARBITRARY_FRAMES, assert_lnl_close, check_arbitrary_frame,
                      check_cluster_path_256, check_tiled_path_512, check_cropped_golden, check_nan_propagation,
                      check_fp64_rescue, check_fused_images, check_hot_pixel_walkers, check_near_centre_walkers, check_pssub_golden,
                      mixed_model_128, fp32_bounds, load_golden, model_from_file,
                      oracle_from_model)


@pytest.fixture(scope='module')
def c1_golden():
    return load_golden('c1_golden.json')


@pytest.mark.parametrize('path', ['fused', 'staged'])
def test_emu_c1_fp64_m3(emu_library, c1_golden, path, monkeypatch):
    monkeypatch.setenv('PSFMC_FORCE_STAGED', '1' if path == 'staged' else '0')
    model = model_from_file('j0005/model_c1.py', 'fp64', library=emu_library,
                            obs_dtype=np.float64)
    thetas = np.array(c1_golden['theta'][:10])
    got = model.log_likelihood_batch(thetas)
    assert_lnl_close(got, c1_golden['lnl']['M3'][:10], 'fp64')
    assert got[c1_golden['names'].index('C_exact_centre')] == -np.inf


@pytest.mark.parametrize('path', ['fused', 'staged'])
def test_emu_c1_fp32_tolerance(emu_library, c1_golden, path, monkeypatch):
    monkeypatch.setenv('PSFMC_FORCE_STAGED', '1' if path == 'staged' else '0')
    model = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library,
                            obs_dtype=np.float64)
    assert model.engine.info()['path'] == (1 if path == 'fused' else 0)
    thetas = np.array(c1_golden['theta'][:16])
    got = model.log_likelihood_batch(thetas)
    assert_lnl_close(got, c1_golden['lnl']['M3'][:16], 'fp32',
                     fp32_bounds(model, thetas))


def test_emu_fused_persistent_loop(emu_library, c1_golden, monkeypatch):
    """More walkers than CTAs: every CTA of the fused kernel walks over several
    walkers (next walker's render overlapped with this walker's inverse rows); the
    result must not depend on the grid size."""
    thetas = np.array(c1_golden['theta'][:11])
    monkeypatch.setenv('PSFMC_FUSED_CTAS', '148')
    wide = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library,
                           obs_dtype=np.float64).log_likelihood_batch(thetas)
    monkeypatch.setenv('PSFMC_FUSED_CTAS', '3')
    model = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library,
                            obs_dtype=np.float64)
    assert model.engine.info()['path'] == 1
    narrow = model.log_likelihood_batch(thetas)
    assert np.array_equal(wide, narrow)
    assert_lnl_close(narrow, c1_golden['lnl']['M3'][:11], 'fp32',
                     fp32_bounds(model, thetas))


def test_emu_rawf32_tracks_m2(emu_library, c1_golden):
    model = model_from_file('j0005/model_c1.py', 'fp64_rawf32', library=emu_library)
    thetas = np.array(c1_golden['theta'][:6])
    got = model.log_likelihood_batch(thetas)
    expect = np.array(c1_golden['lnl']['M2'][:6])
    finite = np.isfinite(expect)
    assert np.array_equal(np.isfinite(got), finite)
    rel = np.abs(got[finite] - expect[finite]) / np.abs(expect[finite])
    assert rel.max() < 1e-7


def test_emu_two_psf_selection(emu_library):
    golden = load_golden('c1_2psf_golden.json')
    thetas = np.array(golden['theta'][:8])
    for precision in ('fp64', 'fp32'):
        model = model_from_file('j0005/model_c1_2psf.py', precision,
                                library=emu_library, obs_dtype=np.float64,
                                two_psf=True)
        got = model.log_likelihood_batch(thetas)
        bounds = fp32_bounds(model, thetas) if precision == 'fp32' else None
        assert_lnl_close(got, golden['lnl']['M3'][:8], precision, bounds)
    bad = thetas[:2].copy()
    bad[:, -1] = (2.6, -0.7)
    assert np.all(model.log_likelihood_batch(bad) == -np.inf)


@pytest.mark.parametrize('index', ['0.5', '4.0'])
def test_emu_c2_galfit(emu_library, index):
    golden = load_golden('c2_golden.json')['cases'][index]
    thetas = np.array(golden['theta'])
    model = model_from_file(golden['model_file'], 'fp64', library=emu_library,
                            obs_dtype=np.float64)
    assert_lnl_close(model.log_likelihood_batch(thetas), golden['lnl']['M3'], 'fp64')
    model32 = model_from_file(golden['model_file'], 'fp32', library=emu_library,
                              obs_dtype=np.float64)
    assert_lnl_close(model32.log_likelihood_batch(thetas), golden['lnl']['M3'],
                     'fp32', fp32_bounds(model32, thetas))


@pytest.mark.parametrize('size,n_sersic', [(16, 1), (32, 1), (64, 1), (256, 2)])
def test_emu_synthetic_frames(emu_library, size, n_sersic):
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    nwalk = 2 if size >= 256 else 4
    comps = synthetic_components(size, n_sersic, dtype=np.float64,
                                 psf_size=min(64, size // 2))
    model = MultiComponentModel(comps, precision='fp64', library=emu_library)
    thetas = draw_walkers_fast(model, nwalk, seed=size)
    expect = oracle_from_model(model).lnlike_batch(thetas)
    assert_lnl_close(model.log_likelihood_batch(thetas), expect, 'fp64')
    comps = synthetic_components(size, n_sersic, dtype=np.float64,
                                 psf_size=min(64, size // 2))
    model32 = MultiComponentModel(comps, precision='fp32', library=emu_library)
    assert_lnl_close(model32.log_likelihood_batch(thetas), expect, 'fp32',
                     fp32_bounds(model32, thetas))


def test_emu_images_and_point_source_subtracted(emu_library, c1_golden):
    model = model_from_file('j0005/model_c1.py', 'fp64', library=emu_library,
                            obs_dtype=np.float64)
    thetas = np.array(c1_golden['theta'][:2])
    imgs = model.engine.render(thetas)
    px = np.array(c1_golden['sample_px'])
    for row in range(2):
        ref = c1_golden['pixels']['M3'][row]
        for key in ('raw_model', 'convolved_model', 'residual', 'composite_ivm'):
            got = imgs[key][row].ravel()[px]
            scale = np.abs(imgs[key][row]).max()
            assert np.allclose(got, ref[key], rtol=1e-9, atol=1e-12 * scale), key
    ref = oracle_from_model(model).images(thetas[0])['point_source_subtracted']
    assert np.allclose(imgs['point_source_subtracted'][0], ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_emu_point_source_subtracted_matches_the_reference(emu_library, precision):
    check_pssub_golden(emu_library, precision, 'c1', rows=[0, 4, 5])
    check_pssub_golden(emu_library, precision, 'c1_2psf', rows=[0, 1])


def test_emu_images_from_the_fused_kernel(emu_library, c1_golden, monkeypatch):
    check_fused_images(emu_library, c1_golden, monkeypatch)


def test_emu_short_theta_rows_are_rejected(emu_library, c1_golden):
    """ld smaller than the number of theta columns the program reads: every C-ABI
    entry point returns PSFMC_ERR_INVALID_ARG instead of reading past the rows."""
    import ctypes
    from psfmc_b200 import _lib
    model = model_from_file('j0005/model_c1.py', 'fp64', library=emu_library)
    eng = model.engine
    short = np.ascontiguousarray(np.array(c1_golden['theta'][:2])[:, :17])
    out = np.zeros(2)
    dbl = ctypes.POINTER(ctypes.c_double)
    args = (eng._handle, short.ctypes.data_as(dbl), 2, 17)
    assert eng._lib.psfmc_lnlike_batch(*args, out.ctypes.data_as(dbl)) == 1
    assert b'ld is smaller' in eng._lib.psfmc_last_error()
    assert eng._lib.psfmc_lnlike_batch_begin(*args, out.ctypes.data_as(dbl)) == 1
    assert eng._lib.psfmc_lnlike_batch_device(
        eng._handle, 0, short.ctypes.data_as(dbl), 2, 17, out.ctypes.data_as(dbl),
        None) == 1
    img = np.zeros((2, 128, 128))
    assert eng._lib.psfmc_render_batch(*args, 1, img.ctypes.data_as(dbl)) == 1
    assert eng._lib.psfmc_accumulate_batch(*args, 1, img.ctypes.data_as(dbl)) == 1
    with pytest.raises(ValueError):
        eng.render(short)
    with pytest.raises(ValueError):
        eng.accumulate(short)
    with pytest.raises(ValueError):
        eng.lnlike(short)
    # the engine is still usable
    assert np.isfinite(eng.lnlike(np.array(c1_golden['theta'][:1]))).all()


def test_emu_batch_edges_and_determinism(emu_library, c1_golden):
    model = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library)
    thetas = np.array(c1_golden['theta'][:5])
    assert model.log_likelihood_batch(thetas[:0].reshape(0, 18)).shape == (0,)
    full = model.log_likelihood_batch(thetas)
    assert np.array_equal(full, model.log_likelihood_batch(thetas))
    assert np.array_equal(full[::-1], model.log_likelihood_batch(thetas[::-1]))
    padded = np.concatenate([thetas, np.zeros((5, 3))], axis=1)
    assert np.array_equal(model.log_likelihood_batch(padded), full)
    pieces = np.concatenate([model.log_likelihood_batch(thetas[s:s + 2])
                             for s in range(0, 5, 2)])
    assert np.array_equal(pieces, full)


@pytest.mark.parametrize('group', ['8', '32'])
def test_emu_device_kappa_matches_scipy(emu_library, group, monkeypatch):
    """kappa = gammaincinv(2n, 0.5) on the device (float64 Halley iteration, 8 or 32
    cooperating lanes) against scipy over the whole prior range, through a one-Sersic
    model whose raw image at the effective radius equals sb_eff."""
    monkeypatch.setenv('PSFMC_PREPARE_GROUP', group)
    # group '8' also switches the Chebyshev table off: the iteration itself is tested
    monkeypatch.setenv('PSFMC_NO_KAPPA_TABLE', '1' if group == '8' else '0')
    from scipy.special import gammaincinv
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Configuration, Sersic
    from psfmc_b200.distributions import Uniform
    size = 32
    obs = np.zeros((size, size))
    ivm = np.ones((size, size))
    psf = np.zeros((8, 8))
    psf[4, 4] = 1.0
    comps = [Configuration(obs, ivm, psf, np.full((8, 8), 1e12), mag_zeropoint=25.0),
             Sersic(xy=(10.0, 16.0), mag=20.0, reff=6.0, reff_b=6.0,
                    index=Uniform(loc=0.01, scale=50), angle=0.0)]
    model = MultiComponentModel(comps, precision='fp64', library=emu_library)
    assert model.engine.info()['kappa_table'] == (0 if group == '8' else 1)
    # 0.03 and 40 lie outside the table (2n in [0.088, 64]): iteration fallback
    ns = np.array([0.03, 0.06, 0.13, 0.2, 0.36, 0.5, 0.75, 1.0, 1.7, 2.5, 4.0, 6.5, 9.9,
                   15.0, 31.9, 40.0])
    raw = model.engine.render(ns[:, None], which=('raw_model',))['raw_model']
    # pixel (x=16, y=16) is exactly at r = reff (circular): value = sbeff * (1 + corr)
    from oracle import psfmc_oracle as orc
    for n, img in zip(ns, raw):
        kappa = gammaincinv(2 * n, 0.5)
        sbeff = orc.sersic_sb_eff(orc.mag_to_flux(20.0, 25.0), n, 6.0, 6.0, kappa)
        grad = -kappa / n          # normed_grad at sq_radii = 1
        want = sbeff * (1 + grad * (1 / 36.0 / 12 * grad))
        assert abs(img[16, 16] / want - 1) < 1e-12, (n, img[16, 16], want)


def test_emu_accumulate_matches_rendered_images(emu_library, c1_golden):
    """psfmc_accumulate_batch (images summed on the device, the IVM in variance
    space) against the same images rendered one by one, and the running-mean
    bookkeeping of MultiComponentModel.accumulate_from_chain against the
    reference-style accumulate_images."""
    model = model_from_file('j0005/model_c1.py', 'fp64', library=emu_library,
                            obs_dtype=np.float64)
    thetas = np.array(c1_golden['theta'][:3] + c1_golden['theta'][6:8])
    imgs = model.engine.render(thetas)
    sums = model.engine.accumulate(thetas)
    for name in imgs:
        want = (1 / imgs[name]).sum(axis=0) if name == 'composite_ivm' \
            else imgs[name].sum(axis=0)
        assert np.allclose(sums[name], want, rtol=1e-12, atol=0), name
    model.reset_images()
    model.accumulate_from_chain(thetas[:2])
    model.accumulate_from_chain(thetas[2:])
    fast = {k: v.copy() for k, v in model.posterior_images.items()}
    model.reset_images()
    model.accumulate_images([{name: imgs[name][row] for name in imgs}
                             for row in range(len(thetas))])
    assert model.accumulated_samples == len(thetas)
    for name in imgs:
        assert np.allclose(fast[name], model.posterior_images[name], rtol=1e-11), name
    empty = model.engine.accumulate(thetas[:0].reshape(0, 18), ('residual',))
    assert np.all(empty['residual'] == 0)


def test_emu_mixed_components_fused_and_fp64(emu_library):
    """Bilinear + edge-clipped Lanczos point sources, fixed and free Sersic parameters,
    angle in radians, fixed sky, bad pixels: fused float32 kernel and float64 staged
    kernels against the oracle."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model64 = mixed_model_128('fp64', library=emu_library)
    thetas = draw_walkers_fast(model64, 5, seed=8)
    expect = oracle_from_model(model64).lnlike_batch(thetas)
    assert np.all(np.isfinite(expect))
    assert_lnl_close(model64.log_likelihood_batch(thetas), expect, 'fp64')
    model32 = mixed_model_128('fp32', library=emu_library)
    assert model32.engine.info()['path'] == 1
    assert_lnl_close(model32.log_likelihood_batch(thetas), expect, 'fp32',
                     fp32_bounds(model32, thetas))


def test_emu_fp64_rescue_of_high_dynamic_range_walkers(emu_library, c1_golden):
    check_fp64_rescue(emu_library, c1_golden)


def test_emu_hot_pixel_walkers(emu_library, c1_golden, monkeypatch):
    monkeypatch.setenv('PSFMC_FUSED_CTAS', '2')      # walker loops: both hot-state halves
    check_hot_pixel_walkers(emu_library, c1_golden)


def test_emu_fused_near_centre_walkers(emu_library):
    check_near_centre_walkers(emu_library)


def test_emu_tiny_sersic_index_is_minus_inf_like_the_reference(emu_library):
    from conftest import check_tiny_index_walkers
    check_tiny_index_walkers(emu_library)


def test_emu_cluster_kernel_256(emu_library, monkeypatch):
    """256 x 256 frame split over a four-CTA cluster (distributed shared memory,
    barrier.cluster emulated): two clusters walking over five walkers."""
    check_cluster_path_256(emu_library, 5, monkeypatch)


def test_emu_tiled_path_512(emu_library, monkeypatch):
    check_tiled_path_512(emu_library, 3, monkeypatch)


@pytest.mark.parametrize('dims', ARBITRARY_FRAMES[:3] + ARBITRARY_FRAMES[4:6])
def test_emu_arbitrary_frame_sizes(emu_library, dims):
    check_arbitrary_frame(emu_library, dims, n_walkers=2)


def test_emu_nan_parameters_give_minus_inf(emu_library, monkeypatch):
    check_nan_propagation(emu_library, monkeypatch)


@pytest.mark.parametrize('tag', ['crop100', 'crop75x100'])
def test_emu_cropped_frames_match_the_reference(emu_library, tag):
    check_cropped_golden(emu_library, tag)


def test_emu_wide_box_fuzz(emu_library):
    """Parameter vectors from boxes far wider than the example's priors (tools/emu_fuzz.py,
    profiles/r2e_emu_fuzz.txt): what the stated tolerances do NOT promise, pinned as what
    they do. 'typical' (reff >= 0.5 px, index >= 0.3, components up to 1500 ADU on 0.007 ADU
    of noise): float64 within FP64_RTOL, float32 within the bound extended by the
    variance-channel term and nearly always within the plain one. 'wide' (centres outside
    the frame, reff down to 0.05 px, index down to 0.05, single pixels of 10^7 ADU): the
    same walkers are finite in the engine and in the reference, float64 agrees to the
    rounding of ITS transform at that dynamic range, float32 to 5e-4 of |lnL|."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'tools'))
    from emu_fuzz import draw, extended_bounds
    from conftest import FP64_RTOL
    models = {prec: model_from_file('j0005/model_c1.py', prec, library=emu_library,
                                    obs_dtype=np.float64) for prec in ('fp64', 'fp32')}
    oracle = oracle_from_model(models['fp64'])
    for box, count in (('typical', 40), ('wide', 40)):
        thetas = draw(np.random.RandomState(17), count, box)
        with np.errstate(all='ignore'):
            expect = oracle.lnlike_batch(thetas)
        finite = np.isfinite(expect)
        assert finite.sum() >= count - 4
        got = {prec: models[prec].log_likelihood_batch(thetas) for prec in models}
        for prec in got:
            assert np.array_equal(np.isfinite(got[prec]), finite), (box, prec)
            assert np.all(got[prec][~finite] == -np.inf)
        rel = {prec: np.abs(got[prec][finite] - expect[finite]) / np.abs(expect[finite])
               for prec in got}
        if box == 'typical':
            assert rel['fp64'].max() <= FP64_RTOL
            err = np.abs(got['fp32'] - expect)[finite]
            assert np.all(err <= extended_bounds(models['fp32'], thetas, oracle)[finite])
            plain = fp32_bounds(models['fp32'], thetas, oracle)[finite]
            assert np.mean(err <= plain) >= 0.9
        else:
            assert rel['fp64'].max() <= 1e-7
            assert rel['fp32'].max() <= 5e-4
