"""
CPU tier: host-side logic that mirrors the reference's Python interface for the
path -- parameter-vector layout, batched priors, model-file parser, FITS / region
readers, the emcee-2.x-style sampler, the pool-like map object, the trace database
and model_galaxy_mcmc (driven end to end through the emulated kernels on a tiny
frame).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, model_from_file


# ------------------------------------------------------- layout and priors --

def test_theta_layout_matches_reference_rule(emu_library):
    """Components in file order, PSF selector last, sorted attribute names inside a
    component, xy = two slots."""
    model = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library)
    assert model.num_params == 18
    assert model.param_names == [
        '0_Sky_adu', '1_PointSource_mag', '1_PointSource_xy',
        '2_Sersic_angle', '2_Sersic_index', '2_Sersic_mag', '2_Sersic_reff',
        '2_Sersic_reff_b', '2_Sersic_xy',
        '3_Sersic_angle', '3_Sersic_index', '3_Sersic_mag', '3_Sersic_reff',
        '3_Sersic_reff_b', '3_Sersic_xy']
    assert model.param_lens == [1, 1, 2, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 2]
    kinds = [p[0] for p in model.program]
    assert kinds == ['sky', 'point', 'sersic', 'sersic']
    ser = model.program[2][2]
    assert ser['angle'] == ('theta', 4) and ser['index'] == ('theta', 5)
    assert ser['x'] == ('theta', 9) and ser['y'] == ('theta', 10)
    assert model.psf_index_slot == ('const', 0.0)
    two = model_from_file('j0005/model_c1_2psf.py', 'fp32', library=emu_library)
    assert two.param_names[-1] == 'PSF_Index' and two.psf_index_slot == ('theta', two.num_params - 1)


def test_batched_priors_equal_reference_lnprior(emu_library):
    golden = load_golden('c1_golden.json')
    model = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library)
    thetas = np.array(golden['theta'])
    got = model.log_priors_batch(thetas)
    want = np.array(golden['lnprior'])
    finite = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), finite)
    assert np.allclose(got[finite], want[finite], rtol=1e-13, atol=0)
    # the scalar path (component objects: the reference's own loop over _priors.values(),
    # ComponentBase.py:121-129) and the batched one add the priors up in the same order:
    # same bits
    for row in np.flatnonzero(finite)[:12]:
        model.param_values = thetas[row]
        assert model.log_priors() == got[row], row
    # reff_b > reff is rejected (Sersic.py:41-45)
    bad = thetas[0].copy()
    bad[7], bad[8] = 3.0, 5.0
    assert model.log_priors_batch(bad[None, :])[0] == -np.inf


def test_native_priors_are_bit_identical_to_scipy(emu_library, monkeypatch):
    """Uniform / Normal columns and all sums in the library's host code
    (psfmc_prior_columns / psfmc_prior_sum), other families (Weibull) in numpy: the same
    bits as one rv_frozen.logpdf call per prior, including values outside the support,
    NaN, +-inf and the reff_b > reff rule; every slower mode gives the same numbers."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library)
    thetas = draw_walkers_fast(model, 700, seed=11)
    rng = np.random.RandomState(5)
    for k in range(120):                       # knock single values out of the support
        thetas[rng.randint(700), rng.randint(model.num_params)] += rng.choice([-1, 1]) * \
            10.0 ** rng.uniform(-1, 3)
    thetas[3, 2] = np.nan
    thetas[5, 0] = np.inf
    thetas[6, 0] = -np.inf
    thetas[8, 7], thetas[8, 8] = 3.0, 5.0      # reff_b > reff
    want = model._log_priors_per_component(thetas)
    assert np.isnan(want[3]) and want[5] == -np.inf and want[8] == -np.inf
    first = model.log_priors_batch(thetas)     # the verifying call returns scipy's
    assert model._prior_mode == 'native'
    assert np.array_equal(first, want, equal_nan=True)
    for rows in (slice(None), slice(0, 1), slice(10, 331)):
        assert np.array_equal(model.log_priors_batch(thetas[rows]), want[rows],
                              equal_nan=True)
    # wider theta rows than the program needs (ld > D)
    wide = np.concatenate([thetas, np.zeros((700, 3))], axis=1)
    assert np.array_equal(model.log_priors_batch(wide), want, equal_nan=True)
    for mode in ('grouped', 'scipy'):
        monkeypatch.setenv('PSFMC_PRIORS', mode)
        other = model_from_file('j0005/model_c1.py', 'fp32', library=emu_library)
        other.log_priors_batch(thetas)
        assert other._prior_mode == mode
        assert np.array_equal(other.log_priors_batch(thetas), want, equal_nan=True)


def test_native_priors_with_constants_and_two_psfs(emu_library):
    """Fixed reff (the rule compares against the constant), a discrete PSF index
    (stays with scipy's logpmf) and a Normal prior."""
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Configuration, Sersic, Sky
    from psfmc_b200.distributions import Normal, Uniform
    two = model_from_file('j0005/model_c1_2psf.py', 'fp32', library=emu_library)
    rng = np.random.RandomState(2)
    thetas = np.array([two.init_params_from_priors(1)[0] for _ in range(40)])
    thetas[:, -1] = rng.uniform(-0.7, 2.7, size=40)       # PSF index incl. out of range
    want = two._log_priors_per_component(thetas)
    two.log_priors_batch(thetas)
    assert two._prior_mode == 'native'
    assert np.array_equal(two.log_priors_batch(thetas), want, equal_nan=True)
    assert np.any(want == -np.inf) and np.any(np.isfinite(want))

    obs = np.zeros((16, 16))
    psf = np.zeros((5, 5))
    psf[2, 2] = 1.0
    config = Configuration(obs, np.ones_like(obs), psf, np.ones_like(psf),
                           mag_zeropoint=25.0)
    comps = [config, Sky(adu=Normal(loc=0.0, scale=0.1)),
             Sersic(xy=Normal(loc=(8.0, 8.0), scale=(1.0, 2.0)), mag=20.0, reff=4.0,
                    reff_b=Uniform(loc=1.0, scale=6.0), index=Uniform(loc=0.5, scale=4.0),
                    angle=0.3)]
    model = MultiComponentModel(comps, precision='fp32', library=emu_library)
    thetas = rng.uniform(-1.0, 9.0, size=(300, model.num_params))
    want = model._log_priors_per_component(thetas)
    model.log_priors_batch(thetas)
    assert model._prior_mode == 'native'
    got = model.log_priors_batch(thetas)
    assert np.array_equal(got, want, equal_nan=True)
    reff_b = thetas[:, model.param_names.index('1_Sersic_reff_b')]
    assert np.all(got[reff_b > 4.0] == -np.inf) and np.any(np.isfinite(got))


def test_prior_entry_points_reject_bad_tables(emu_library):
    import ctypes
    from psfmc_b200 import _lib
    lib = _lib.load(emu_library)
    dbl_p = ctypes.POINTER(ctypes.c_double)
    theta = np.zeros((4, 3))
    logp = np.zeros((4, 3))
    cols = (_lib.PriorColumn * 3)()
    cols[0].family = 7
    call = lambda: lib.psfmc_prior_columns(cols, 3, theta.ctypes.data_as(dbl_p), 4, 3,
                                           logp.ctypes.data_as(dbl_p), 3)
    assert call() != 0 and b'family' in lib.psfmc_last_error()
    cols[0].family, cols[0].theta_index = _lib.PRIOR_UNIFORM, 3
    assert call() != 0
    cols[0].theta_index, cols[0].scale, cols[0].valid = 0, 1.0, 1
    assert call() == 0 and logp[0, 0] == 0.0
    terms = (_lib.PriorTerm * 2)()
    terms[0].component, terms[0].first_column, terms[0].n_columns = 1, 0, 1
    terms[1].component, terms[1].first_column, terms[1].n_columns = 0, 1, 1
    out = np.zeros(4)
    rc = lib.psfmc_prior_sum(logp.ctypes.data_as(dbl_p), 4, 3, theta.ctypes.data_as(dbl_p), 3,
                             terms, 2, None, 0, 2, out.ctypes.data_as(dbl_p))
    assert rc != 0 and b'ascending' in lib.psfmc_last_error()


def test_discrete_prior_rounds_half_to_even():
    from psfmc_b200.distributions import DiscreteUniform, Normal
    prior = DiscreteUniform(low=0, high=2)
    seen = []
    for val in (0.5, 1.5, -0.5, 1.49):
        prior.value = val
        seen.append(prior.value)
    assert seen == [0, 2, 0, 1]
    norm = Normal(loc=0, scale=1)
    norm.value = np.array([0.25])
    assert isinstance(norm.value, float)
    assert np.isclose(norm.logp(0.0), -0.5 * np.log(2 * np.pi))


def test_all_reference_distribution_names_exist():
    import re
    from psfmc_b200 import distributions
    path = '/root/reference/psfMC/distributions.py'
    if not os.path.exists(path):
        pytest.skip('reference not present')
    names = re.findall(r"'(\w+)':\s*'\w+'", open(path).read())
    assert len(names) > 90
    for name in names:
        assert hasattr(distributions, name), name


def test_model_parser_syntax_and_relative_paths(tmp_path):
    from psfmc_b200.components import Configuration, Sersic, Sky
    from psfmc_b200.model_parser import component_list_from_file
    comps = component_list_from_file(os.path.join(GOLDEN, 'j0005', 'model_c1.py'))
    assert [type(c).__name__ for c in comps] == \
        ['Configuration', 'Sky', 'PointSource', 'Sersic', 'Sersic']
    assert comps[0].obs_data.shape == (128, 128)
    # explicit psfMC imports in old model files keep working
    src = open(os.path.join(GOLDEN, 'galfit', 'model_n1.0.py')).read()
    model = tmp_path / 'm.py'
    model.write_text('from psfMC.ModelComponents import *\n'
                     'from psfMC.distributions import *\n' +
                     src.replace("'gfsim", "'" + os.path.join(GOLDEN, 'galfit', 'gfsim')
                                 ).replace("'ivm_const", "'" + os.path.join(GOLDEN, 'galfit', 'ivm_const')
                                 ).replace("'psf", "'" + os.path.join(GOLDEN, 'galfit', 'psf')))
    comps = component_list_from_file(str(model))
    assert isinstance(comps[0], Configuration) and isinstance(comps[1], Sersic)
    assert Sky is not None


# ------------------------------------------------------------ file formats --

def test_fits_image_roundtrip_and_gzip(tmp_path):
    from psfmc_b200 import fitsio
    rng = np.random.RandomState(1)
    for dtype in (np.float32, np.float64):
        img = rng.standard_normal((7, 13)).astype(dtype)
        name = str(tmp_path / 'img_{}.fits'.format(np.dtype(dtype).name))
        fitsio.writeto(name, img)
        back = fitsio.getdata(name)
        assert back.dtype == dtype and np.array_equal(back, img)
    gal = fitsio.getdata(os.path.join(GOLDEN, 'galfit', 'gfsim_n1.0.fits.gz'))
    assert gal.shape == (128, 128) and gal.dtype == np.float32
    hdr = fitsio.getheader(os.path.join(GOLDEN, 'galfit', 'gfsim_n1.0.fits.gz'))
    assert abs(float(hdr['MAGZPT']) - 26.2303) < 1e-6
    with pytest.raises(IOError):
        fitsio.getdata(os.path.join(GOLDEN, 'j0005', 'mask_J0005-0006.reg'))


def test_region_mask_pixel_indices():
    from psfmc_b200 import preprocess
    golden = load_golden('c1_golden.json')['setup']
    bad = preprocess.mask_from_file(
        os.path.join(GOLDEN, 'j0005', 'mask_J0005-0006.reg'), (128, 128))
    good = np.flatnonzero(~bad)
    assert (len(good), int(good.sum()), int(good[0]), int(good[-1])) == \
        (golden['n_good'], golden['good_index_sum'], golden['first_good'],
         golden['last_good'])
    row64 = np.flatnonzero(~bad[64])
    assert row64[0] == 8 and row64[-1] == 100


def test_psf_normalisation_uses_fsum():
    from psfmc_b200 import fitsio, preprocess
    psf = fitsio.getdata(os.path.join(GOLDEN, 'j0005', 'sci_psf.fits'))
    ivm = fitsio.getdata(os.path.join(GOLDEN, 'j0005', 'ivm_psf.fits'))
    from math import fsum
    assert abs(fsum(psf.flat) - 1858.0593897353801) < 1e-9
    norm, var = preprocess.preprocess_psf(psf, ivm)
    assert norm.dtype == np.float32 and abs(norm.sum() - 1) < 1e-5
    assert np.all(var >= 0)


# -------------------------------------------------------------- the sampler --

def _gauss_lnpost(theta, mean, ivar):
    return -0.5 * float(np.sum((theta - mean) ** 2 * ivar))


def test_sampler_recovers_a_gaussian():
    from psfmc_b200.sampler import EnsembleSampler
    mean, ivar = np.array([1.0, -2.0, 0.5]), np.array([1.0, 4.0, 0.25])
    sampler = EnsembleSampler(20, 3, _gauss_lnpost, args=[mean, ivar])
    sampler._random.seed(42)
    p0 = mean + np.random.RandomState(0).standard_normal((20, 3))
    pos, lnp, _ = sampler.run_mcmc(p0, 300)
    sampler.reset()
    sampler.run_mcmc(pos, 1500)
    flat = sampler.flatchain
    assert sampler.chain.shape == (20, 1500, 3)
    assert np.allclose(flat.mean(axis=0), mean, atol=0.12)
    assert np.allclose(flat.var(axis=0), 1 / ivar, rtol=0.15)
    assert 0.3 < sampler.acceptance_fraction.mean() < 0.8
    tau = sampler.get_autocorr_time(c=1)
    assert tau.shape == (3,) and np.all(tau > 1)


def test_sampler_call_pattern_and_random_stream():
    """One map over k walkers, then two maps of k/2 per iteration; RNG call order
    rand(Ns), randint(Nc, size=Ns), rand(Ns) per half-step."""
    from psfmc_b200.sampler import EnsembleSampler
    sizes = []

    class Pool(object):
        def map(self, func, items):
            items = list(items)
            sizes.append(len(items))
            return [func(it) for it in items]

    sampler = EnsembleSampler(8, 2, lambda th: (-0.5 * float(th @ th), {'n': 1}),
                              pool=Pool())
    sampler._random.seed(5)
    p0 = np.random.RandomState(1).standard_normal((8, 2))
    results = list(sampler.sample(p0, iterations=3))
    assert sizes == [8, 4, 4, 4, 4, 4, 4]
    assert len(results[0]) == 4 and len(results[0][3]) == 8
    # replay the first half-step by hand from the same stream
    rng = np.random.mtrand.RandomState(5)
    zz = ((2.0 - 1.0) * rng.rand(4) + 1) ** 2.0 / 2.0
    partner = rng.randint(4, size=(4,))
    q = p0[4:][partner] - zz[:, None] * (p0[4:][partner] - p0[:4])
    lnp0 = np.array([-0.5 * float(t @ t) for t in p0[:4]])
    newlnp = np.array([-0.5 * float(t @ t) for t in q])
    accept = (2 - 1.0) * np.log(zz) + newlnp - lnp0 > np.log(rng.rand(4))
    again = EnsembleSampler(8, 2, lambda th: -0.5 * float(th @ th))
    again._random.seed(5)
    pos, _, _ = next(again.sample(p0, iterations=1))
    # rows of the first half that were accepted moved to q
    moved = np.any(pos[:4] != p0[:4], axis=1)
    assert np.array_equal(moved, accept)
    assert np.allclose(pos[:4][accept], q[accept])
    with pytest.raises(AssertionError):
        EnsembleSampler(7, 2, lambda th: 0.0)
    with pytest.raises(AssertionError):
        EnsembleSampler(2, 2, lambda th: 0.0)
    with pytest.raises(ValueError):
        list(EnsembleSampler(4, 2, lambda th: float('nan')).sample(p0[:4]))


# ---------------------------------------------------- pool, database, driver --

@pytest.fixture(scope='module')
def tiny_model(emu_library):
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import synthetic_components
    comps = synthetic_components(32, 1, psf_size=16)
    return MultiComponentModel(comps, precision='fp32', library=emu_library)


def test_split_host_call(tiny_model):
    """psfmc_lnlike_batch_begin / _end: same numbers as the blocking call, one batch
    in flight per engine, _end without _begin is an error."""
    from psfmc_b200._lib import EngineError
    engine = tiny_model.engine
    thetas = tiny_model.init_params_from_priors(6)
    want = engine.lnlike(thetas)
    engine.lnlike_begin(thetas)
    with pytest.raises(EngineError, match='in flight'):
        engine.lnlike_begin(thetas)
    assert np.array_equal(engine.lnlike_end(), want)
    with pytest.raises(EngineError, match='no batch in flight'):
        from psfmc_b200 import _lib
        _lib.check(engine._lib, engine._lib.psfmc_lnlike_batch_end(engine._handle))
    assert np.array_equal(engine.lnlike(thetas), want)
    assert np.array_equal(tiny_model.log_posterior_batch(thetas[:1]),
                          tiny_model.log_posterior_batch(thetas)[:1])
    tiny_model.overlap_min_batch = 2          # overlapped priors on a small batch
    lnpost = tiny_model.log_posterior_batch(thetas)
    assert np.array_equal(lnpost, want + tiny_model.log_priors_batch(thetas))


def test_batch_pool_is_a_drop_in_for_map(tiny_model):
    from psfmc_b200 import BatchPool
    from psfmc_b200.sampler import EnsembleSampler
    from psfmc_b200.synthetic import draw_walkers_fast
    model = tiny_model
    thetas = draw_walkers_fast(model, 6, seed=3)
    thetas[2, model.param_names.index('2_Sersic_reff_b') + 1] = 99.0   # prior -inf
    pool = BatchPool(model)
    sampler = EnsembleSampler(6, model.num_params, model.log_posterior,
                              kwargs={'model': model}, pool=pool,
                              live_dangerously=True)
    batched = pool.map(sampler.lnprobfn, list(thetas))
    single = [sampler.lnprobfn(row) for row in thetas]      # emcee without a pool
    assert [b[0] for b in batched] == [s[0] for s in single]
    assert all(isinstance(b[1], dict) for b in batched)
    assert batched[2][0] == -np.inf and pool.calls == 1 and pool.evaluations == 6
    with_blobs = BatchPool(model, with_blobs=True).map(sampler.lnprobfn, list(thetas[:3]))
    assert set(with_blobs[0][1]) == {'raw_model', 'convolved_model', 'residual',
                                     'composite_ivm', 'point_source_subtracted'}
    assert with_blobs[2][1] == {}
    assert with_blobs[0][1]['raw_model'].shape == (32, 32)
    # the array protocol of this package's own sampler: same numbers, no lists
    lnpost, blobs = pool.map_batch(sampler.lnprobfn, thetas)
    assert blobs is None and lnpost.tolist() == [b[0] for b in batched]
    lnpost, blobs = BatchPool(model, with_blobs=True).map_batch(sampler.lnprobfn, thetas[:3])
    assert lnpost.tolist() == [b[0] for b in batched[:3]] and blobs[2] == {}
    assert np.array_equal(sampler._get_lnprob(thetas)[0], lnpost_all(batched))
    # lists of Python lists / tuples still work (the slow stacking path)
    assert [b[0] for b in pool.map(sampler.lnprobfn, [tuple(r) for r in thetas])] == \
        [b[0] for b in batched]


def lnpost_all(results):
    return np.array([r[0] for r in results])


def test_database_roundtrip(tmp_path, tiny_model):
    from psfmc_b200.database import (filter_lowp_walkers, load_database,
                                     param_matrix, save_database)

    class FakeSampler(object):
        pass
    rng = np.random.RandomState(0)
    sam = FakeSampler()
    ndim = tiny_model.num_params
    sam.chain = rng.standard_normal((6, 5, ndim))
    sam.lnprobability = rng.standard_normal((6, 5))
    sam.lnprobability[3] = -1e9                      # a lost walker
    name = str(tmp_path / 'db.fits')
    db = save_database(sam, tiny_model, name, {'MCITER': 5, 'MCBURN': 2,
                                               'MCCHAINS': 6, 'MCCONVRG': False,
                                               'MCACCEPT': 0.25})
    assert db.colnames == tiny_model.param_names + ['lnprobability', 'walker', 'sample']
    assert len(db) == 30 and db['1_PointSource_xy'].shape == (30, 2)
    assert np.array_equal(param_matrix(db, tiny_model), sam.chain.reshape(30, ndim))
    assert db.meta['MCITER'] == 5 and db.meta['MCCONVRG'] is False
    best = np.argmax(sam.lnprobability.ravel())
    assert db.meta['MAPWLKR'] == best // 5
    again = load_database(name)
    assert np.array_equal(again['lnprobability'], db['lnprobability'])
    kept = filter_lowp_walkers(db, percentile=10)
    assert 3 not in set(kept['walker']) and len(kept) == 25


def test_model_galaxy_mcmc_end_to_end(tmp_path, tiny_model):
    from psfmc_b200 import fitsio, model_galaxy_mcmc
    out = str(tmp_path / 'run')
    nwalk = 2 * tiny_model.num_params + 2
    db = model_galaxy_mcmc(tiny_model, output_name=out, iterations=4, burn=2,
                           chains=nwalk, seed=7, verbose=False,
                           write_fits=('raw_model', 'residual', 'composite_ivm'))
    assert len(db) == nwalk * 4 and db.meta['MCBURN'] == 2
    assert db.meta['MCCHAINS'] == nwalk and np.all(np.isfinite(db['lnprobability']))
    for ftype in ('raw_model', 'residual', 'composite_ivm'):
        img = fitsio.getdata('{}_{}.fits'.format(out, ftype))
        assert img.shape == (32, 32) and np.all(np.isfinite(img))
    # fit summary in the image headers (cf. analysis/images.py:104-143): sampler metadata,
    # 'mean +/- std' per parameter under its FITS abbreviation, the PSF image
    hdr = fitsio.getheader('{}_raw_model.fits'.format(out))
    assert hdr['MCCHAINS'] == nwalk and hdr['MCITER'] == 4
    assert 'PSFIMG' in hdr
    from psfmc_b200.database import filter_lowp_walkers
    kept = filter_lowp_walkers(db, percentile=10)      # as save_posterior_images does
    for name, abbr in zip(tiny_model.param_names, tiny_model.param_fits_abbrs):
        column = np.asarray(kept[name], dtype=np.float64)
        mean, std = np.mean(column, axis=0), np.std(column, axis=0)
        key = abbr if len(abbr) > 8 else abbr.upper()     # long keys: HIERARCH cards
        assert ' +/- ' in hdr[key], (key, hdr[key])
        if np.ndim(mean) == 0:
            assert hdr[key] == '{:0.4g} +/- {:0.4g}'.format(float(mean), float(std))
        else:
            assert hdr[key].startswith('(')
    # resume rule of the reference: an existing database is loaded, not re-sampled
    again = model_galaxy_mcmc(tiny_model, output_name=out, iterations=4, burn=2,
                              chains=nwalk, write_fits=(), verbose=False)
    assert np.array_equal(again['lnprobability'], db['lnprobability'])


def test_accumulate_images_running_mean(tiny_model):
    model = tiny_model
    model.reset_images()
    a = {'raw_model': np.full((32, 32), 2.0), 'composite_ivm': np.full((32, 32), 4.0)}
    b = {'raw_model': np.full((32, 32), 4.0), 'composite_ivm': np.full((32, 32), 1.0)}
    model.accumulate_images([a, b])
    assert model.accumulated_samples == 2
    assert np.allclose(model.posterior_images['raw_model'], 3.0)
    # the IVM is averaged in variance space (models.py:81-82,96-97)
    assert np.allclose(model.posterior_images['composite_ivm'], 1 / ((0.25 + 1.0) / 2))
    model.reset_images()


# ------------------------------------------- drop-in for the unmodified reference --

_BRIDGE_CHECK = r"""
import json, os, sys
import numpy as np
root, emu_lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
from oracle import refshim
golden = json.load(open(os.path.join(root, 'tests/golden/c1_golden.json')))
model = refshim.build_reference_model(
    os.path.join(root, 'tests/golden/j0005/model_c1.py'), 'M3')
from psfmc_b200.bridge import pool_for_reference_model
pool = pool_for_reference_model(model, precision='fp64', library=emu_lib)
rows = [0, 1, 2, 3, 4, 8, 9, 10]
thetas = [np.array(golden['theta'][r]) for r in rows]
bad = np.array(golden['theta'][0]); bad[7], bad[8] = 3.0, 5.0      # reff_b > reff
out = np.array(golden['theta'][1]); out[-1] = 400.0                 # angle outside its prior
thetas += [bad, out]
batched = pool.map(None, thetas)
# the priors went through the library's plan, validated against the reference's own code
assert pool.priors and (pool.priors.weibull_native or pool.priors.other)
for r, theta, (lnpost, blob) in zip(rows + [-1, -2], thetas, batched):
    want, _ = type(model).log_posterior(theta, model=model)     # the reference itself
    assert blob == {}
    if np.isfinite(want):
        assert abs(lnpost - want) <= 1e-10 * abs(want), (r, lnpost, want)
    else:
        assert lnpost == -np.inf, (r, lnpost, want)
assert batched[-1][0] == -np.inf and batched[-2][0] == -np.inf
# the plan's priors against the reference's scalar code, row by row, and the per-walker
# path (PSFMC_BRIDGE_PRIORS=reference) against the batched one
for theta, got in zip(thetas, pool.priors.lnprior(np.stack(thetas))):
    model.param_values = theta
    want = model.log_priors()
    assert got == want or (np.isfinite(want) and abs(got - want) <= 4 * np.spacing(abs(want))) \
        or (np.isnan(got) and np.isnan(want)), (got, want)
os.environ['PSFMC_PRIORS_STRICT'] = '1'
strict = pool_for_reference_model(model, precision='fp64', library=emu_lib)
again = strict.map(None, thetas)
assert strict.priors and not strict.priors.weibull_native and strict.priors.other
for theta, got in zip(thetas, strict.priors.lnprior(np.stack(thetas))):
    model.param_values = theta
    want = model.log_priors()
    assert got == want or (np.isnan(got) and np.isnan(want)), (got, want)
os.environ['PSFMC_BRIDGE_PRIORS'] = 'reference'
slow = pool_for_reference_model(model, precision='fp64', library=emu_lib)
per_walker = slow.map(None, thetas)
assert slow.priors is False
assert [a for a, _ in per_walker] == [a for a, _ in again]
print('BRIDGE-OK')
"""


def test_bridge_is_a_drop_in_for_the_unmodified_reference(emu_library):
    """The reference's OWN MultiComponentModel + our pool object (kernels emulated on
    the CPU here) reproduce the reference's log_posterior to 1e-10."""
    import subprocess
    import sys
    from conftest import ROOT
    if not os.path.isdir('/root/reference/psfMC'):
        pytest.skip('/root/reference not present (GPU box)')
    proc = subprocess.run([sys.executable, '-c', _BRIDGE_CHECK, ROOT, emu_library],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          universal_newlines=True, timeout=600)
    assert proc.returncode == 0 and 'BRIDGE-OK' in proc.stdout, proc.stdout[-3000:]


def test_rows_as_block_reassembles_emcee_row_lists():
    """BatchPool.map gets emcee's ``[p[i] for i in range(len(p))]``: consecutive row
    views of one array are re-assembled without copying; anything else is stacked."""
    from psfmc_b200.pool import rows_as_block
    rng = np.random.RandomState(0)
    p = rng.rand(64, 18)
    half = p[32:]
    block = rows_as_block([half[i] for i in range(len(half))])
    assert np.shares_memory(block, p) and np.array_equal(block, half)
    taken = p[np.arange(0, 64, 2)]                       # emcee's fancy-indexed half
    assert np.array_equal(rows_as_block([taken[i] for i in range(32)]), taken)
    order = rng.permutation(32)
    assert np.array_equal(rows_as_block([half[i] for i in order]), half[order])
    order[0], order[16], order[31] = 0, 16, 31           # lined-up ends, shuffled middle
    order[1], order[2] = 2, 1
    got = rows_as_block([half[i] for i in order])
    assert np.array_equal(got, half[order]) or np.array_equal(got, half)  # documented limit
    wide = rng.rand(40, 20)[:, :18]                      # strided rows: not one block
    assert np.array_equal(rows_as_block([wide[i] for i in range(40)]), wide)
    assert np.array_equal(rows_as_block([np.array(r) for r in half[:5]]), half[:5])
    assert np.array_equal(rows_as_block([list(r) for r in half[:3]]), half[:3])
    assert np.array_equal(rows_as_block([half[0]]), half[:1])
