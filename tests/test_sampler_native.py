"""
The sampler's inner loop inside the library (psfmc_ensemble_run, psfmc_lnpost_batch,
psfmc_rng_fill; include/psfmc_b200.h) against this package's numpy loop (sampler.py, emcee
2.x semantics) and against numpy's own random stream. CPU tier: the library is the
emulator build of the same sources (tests/emu); the GPU tier repeats the chain comparison
on the real library (test_gpu_parity.py).
"""
import ctypes

import numpy as np
import pytest

from conftest import oracle_from_model


def _state_arrays(rng):
    state = rng.get_state()
    return np.array(state[1], dtype=np.uint32), ctypes.c_int32(int(state[2]))


def test_rng_fill_continues_numpys_stream(emu_library):
    """random_sample and randint(bound) of a RandomState, interleaved, from a handed-over
    MT19937 state; the state handed back continues numpy's stream."""
    from psfmc_b200 import _lib
    lib = _lib.load(emu_library)
    dbl_p = ctypes.POINTER(ctypes.c_double)
    u32_p = ctypes.POINTER(ctypes.c_uint32)
    for seed in (0, 5, 12345):
        ref = np.random.RandomState(seed)
        mine = np.random.RandomState(seed)
        key, pos = _state_arrays(mine)
        for kind, n, bound in ((0, 700, 0), (1, 333, 125), (1, 10, 1), (0, 5, 0),
                               (1, 400, 2), (1, 257, 1000), (1, 50, 2 ** 31 + 7),
                               (1, 64, 4096), (0, 1300, 0)):
            out = np.empty(n)
            _lib.check(lib, lib.psfmc_rng_fill(key.ctypes.data_as(u32_p), ctypes.byref(pos),
                                               kind, n, bound, out.ctypes.data_as(dbl_p)))
            expect = ref.rand(n) if kind == 0 else ref.randint(bound, size=(n,))
            assert np.array_equal(out, np.asarray(expect, dtype=np.float64)), (seed, kind, bound)
        state = mine.get_state()
        mine.set_state((state[0], key, int(pos.value), state[3], state[4]))
        assert np.array_equal(mine.rand(10), ref.rand(10))
        assert np.array_equal(mine.standard_normal(4), ref.standard_normal(4))


def _small_model(library, weibull):
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Sersic
    from psfmc_b200.distributions import Uniform, WeibullMinimum
    from psfmc_b200.synthetic import synthetic_components
    size = 16
    comps = synthetic_components(size, 0 if weibull else 1, dtype=np.float64, psf_size=8)
    if weibull:
        centre, box = np.array((size / 2.0, size / 2.0)), np.array((3.0, 3.0))
        comps.append(Sersic(xy=Uniform(loc=centre - box, scale=2 * box),
                            mag=Uniform(loc=21, scale=4), reff=Uniform(loc=3, scale=4),
                            reff_b=Uniform(loc=2, scale=3),
                            index=WeibullMinimum(c=1.5, scale=4),
                            angle=Uniform(loc=0, scale=180), angle_degrees=True))
    return MultiComponentModel(comps, precision='fp64', library=library)


def _run(model, start, native, monkeypatch, iterations=5, thin=1, stepwise=False,
         device=False):
    from psfmc_b200 import BatchPool
    from psfmc_b200.sampler import EnsembleSampler
    monkeypatch.setenv('PSFMC_NATIVE_SAMPLER', '1' if native else '0')
    monkeypatch.setenv('PSFMC_DEVICE_LOOP', '1' if device else '0')
    nwalk, ndim = start.shape
    sampler = EnsembleSampler(nwalk, ndim, model.log_posterior, kwargs={'model': model},
                              pool=BatchPool(model))
    sampler._random.seed(5)
    if stepwise:
        yielded = []
        for pos, lnprob, rstate in sampler.sample(start, iterations=iterations, thin=thin):
            yielded.append((pos.copy(), lnprob.copy(), rstate[2]))
        pos, lnprob = yielded[-1][:2]
    else:
        pos, lnprob, _ = sampler.run_mcmc(start, iterations, thin=thin)
        yielded = None
    return {'pos': pos.copy(), 'lnprob': lnprob.copy(), 'chain': sampler.chain.copy(),
            'lnprobability': sampler.lnprobability.copy(),
            'naccepted': sampler.naccepted.copy(), 'iterations': sampler.iterations,
            'next': sampler._random.rand(4), 'yielded': yielded}


@pytest.mark.parametrize('device', [False, True])
@pytest.mark.parametrize('case', ['uniform', 'weibull_strict', 'weibull_native'])
def test_library_loop_reproduces_the_numpy_loop(emu_library, monkeypatch, case, device):
    """Same seed, same start: chain, lnprobability, acceptance counts and the random
    stream after the run are those of the numpy loop. 'uniform': every prior evaluated in
    the library; 'weibull_strict': the Weibull column through the callback (bit for bit);
    'weibull_native': the Weibull column in the library with the C library's log / pow --
    positions identical, lnprob to the last digits (numpy's SVML pow on AVX-512 hosts)."""
    from psfmc_b200.synthetic import draw_walkers_fast
    weibull = case != 'uniform'
    monkeypatch.setenv('PSFMC_PRIORS_STRICT', '1' if case == 'weibull_strict' else '0')
    model = _small_model(emu_library, weibull)
    nwalk = 2 * model.num_params + 2
    start = draw_walkers_fast(model, nwalk, seed=3)
    # tighten the ensemble so that a good share of the proposals is accepted
    start = start[0] + 0.02 * (start - start[0])
    assert np.all(np.isfinite(model.log_posterior_batch(start)))
    # device: proposals, priors and acceptance in kernels around the lnL kernels
    # (PSFMC_ENS_DEVICE; with the Weibull column in Python the library falls back to its
    # host loop)
    ref = _run(model, start, False, monkeypatch)
    launches = model.engine.info()['launches_total']
    got = _run(model, start, True, monkeypatch, device=device)
    per_half = (model.engine.info()['launches_total'] - launches) / 10.0
    holder = model._sampler_plan
    if device and case != 'weibull_strict':
        assert per_half > 6.9, per_half       # propose + prepare + 3 staged + finalize + accept
    else:
        assert per_half < 6.9, per_half
    assert holder, 'the library loop was not used'
    assert holder['weibull_native'] == (case != 'weibull_strict')
    assert holder['python_columns'] == (case == 'weibull_strict')
    assert 0 < ref['naccepted'].sum() < ref['naccepted'].size * ref['iterations']
    assert np.array_equal(got['chain'], ref['chain'])
    assert np.array_equal(got['naccepted'], ref['naccepted'])
    assert np.array_equal(got['next'], ref['next'])
    assert got['iterations'] == ref['iterations']
    if case == 'weibull_native':
        np.testing.assert_allclose(got['lnprobability'], ref['lnprobability'], rtol=1e-13)
    else:
        assert np.array_equal(got['lnprobability'], ref['lnprobability'])
        assert np.array_equal(got['lnprob'], ref['lnprob'])
    # the stored lnprob is the posterior of the stored position
    np.testing.assert_allclose(got['lnprobability'][:, -1],
                               model.log_posterior_batch(got['chain'][:, -1]), rtol=1e-13)


def test_library_loop_stepwise_and_thinned(emu_library, monkeypatch):
    """sample() (one library call per iteration, a yield after each) and thin = 2."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model = _small_model(emu_library, False)
    nwalk = 2 * model.num_params + 2
    start = draw_walkers_fast(model, nwalk, seed=4)
    start = start[0] + 0.02 * (start - start[0])
    ref = _run(model, start, False, monkeypatch, iterations=6, thin=2, stepwise=True)
    for device in (True, False):
        got = _run(model, start, True, monkeypatch, iterations=6, thin=3, device=device)
        want = _run(model, start, False, monkeypatch, iterations=6, thin=3)
        assert got['chain'].shape == (nwalk, 2, model.num_params)
        assert np.array_equal(got['chain'], want['chain'])
        assert np.array_equal(got['lnprobability'], want['lnprobability'])
        assert np.array_equal(got['pos'], want['pos'])
        assert np.array_equal(got['next'], want['next'])
    got = _run(model, start, True, monkeypatch, iterations=6, thin=2, stepwise=True,
               device=True)
    assert np.array_equal(got['chain'], ref['chain'])
    got = _run(model, start, True, monkeypatch, iterations=6, thin=2, stepwise=True)
    assert got['chain'].shape == (nwalk, 3, model.num_params)
    assert np.array_equal(got['chain'], ref['chain'])
    assert np.array_equal(got['lnprobability'], ref['lnprobability'])
    assert len(got['yielded']) == 6
    for (p1, l1, s1), (p2, l2, s2) in zip(got['yielded'], ref['yielded']):
        assert np.array_equal(p1, p2) and np.array_equal(l1, l2) and s1 == s2
    one = _run(model, start, True, monkeypatch, iterations=6, thin=2)
    assert np.array_equal(one['chain'], ref['chain'])
    assert np.array_equal(one['next'], ref['next'])


def test_device_loop_chain_blocks(emu_library, monkeypatch):
    """The device loop hands its chain over in blocks (two buffers: one fills while the
    other travels): blocks of 3 iterations, 11 iterations, thin 1 and 2."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model = _small_model(emu_library, False)
    nwalk = 2 * model.num_params + 2
    start = draw_walkers_fast(model, nwalk, seed=6)
    start = start[0] + 0.02 * (start - start[0])
    monkeypatch.setenv('PSFMC_CHAIN_BLOCK_BYTES', str(3 * nwalk * model.num_params * 8))
    for thin, iters in ((1, 11), (2, 10)):
        want = _run(model, start, False, monkeypatch, iterations=iters, thin=thin)
        got = _run(model, start, True, monkeypatch, iterations=iters, thin=thin, device=True)
        assert got['chain'].shape[1] == iters // thin
        assert np.array_equal(got['chain'], want['chain'])
        assert np.array_equal(got['lnprobability'], want['lnprobability'])
        assert np.array_equal(got['naccepted'], want['naccepted'])
        assert np.array_equal(got['next'], want['next'])


def test_lnpost_batch_matches_log_posterior_batch(emu_library):
    """psfmc_lnpost_batch: priors in the library while the engine computes; rows with a
    dead prior (reff_b > reff, outside a Uniform) and a dead likelihood come out -inf."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model = _small_model(emu_library, True)
    thetas = draw_walkers_fast(model, 12, seed=9)
    names = model.param_names
    thetas[1, 0] = 1e3                   # far outside the sky prior's bulk, still finite
    thetas[2, -1] = 400.0                # angle outside Uniform(0, 180)
    reff = [i for i, n in enumerate(names) if n.endswith('reff')][0]
    reff_b = [i for i, n in enumerate(names) if n.endswith('reff_b')][0]
    thetas[3, reff_b] = thetas[3, reff] + 0.5
    expect = model.log_posterior_batch(thetas)
    holder = model.native_sampler_plan(draw_walkers_fast(model, 8, seed=1))
    assert holder is not None
    got = model.engine.lnpost(holder['plan'], thetas)
    assert np.array_equal(np.isfinite(got), np.isfinite(expect))
    assert not np.isfinite(got[2]) and not np.isfinite(got[3])
    finite = np.isfinite(expect)
    np.testing.assert_allclose(got[finite], expect[finite], rtol=1e-13)
    # against the oracle's lnL + scipy priors
    oracle = oracle_from_model(model)
    lnl = oracle.lnlike_batch(thetas[finite])
    np.testing.assert_allclose(got[finite], lnl + model.log_priors_batch(thetas[finite]),
                               rtol=1e-9)


def test_lnpost_batch_leaves_dead_rows_out(emu_library):
    """More rows than a fixed-size (graph) batch: rows whose closed-form priors are dead
    are not sent to the engine at all, the others come back in their places; an all-dead
    batch makes no engine call."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model = _small_model(emu_library, True)
    thetas = draw_walkers_fast(model, 1300, seed=11)   # (several host threads)
    rng = np.random.RandomState(2)
    kill = rng.rand(1300) < 0.3
    thetas[kill, -1] = -5.0 - rng.rand(kill.sum())        # angle below Uniform(0, 180)
    # a dead row with parameters the engine could not digest: never evaluated
    thetas[np.flatnonzero(kill)[0], :] = np.nan
    thetas[np.flatnonzero(kill)[0], -1] = -1.0
    holder = model.native_sampler_plan(draw_walkers_fast(model, 8, seed=1))
    before = model.engine.info()['launches_total']
    got = model.engine.lnpost(holder['plan'], thetas)
    assert np.all(np.isneginf(got[kill])) and np.all(np.isfinite(got[~kill]))
    expect = model.log_posterior_batch(thetas[~kill])
    np.testing.assert_allclose(got[~kill], expect, rtol=1e-13)
    launched = model.engine.info()['launches_total']
    assert launched > before
    # (a batch is screened before the launch when it is small or when the previous call
    # lost more than 2 % of its rows; the call above -- all rows alive -- switched that off)
    launched = model.engine.info()['launches_total']
    dead = thetas[kill]
    assert np.all(np.isneginf(model.engine.lnpost(holder['plan'], dead[1:])))     # unscreened
    assert model.engine.info()['launches_total'] > launched
    launched = model.engine.info()['launches_total']
    assert np.all(np.isneginf(model.engine.lnpost(holder['plan'], dead)))         # screened
    assert np.all(np.isneginf(model.engine.lnpost(holder['plan'], dead[:100])))   # small
    assert model.engine.info()['launches_total'] == launched


def test_host_threads_and_fork(emu_library, monkeypatch):
    """The library's host threads (per-row work of large batches) are not inherited by a
    forked child: the child gets a fresh pool and computes the same numbers."""
    import os
    import time
    from psfmc_b200.synthetic import draw_walkers_fast
    monkeypatch.setenv('PSFMC_HOST_THREADS', '4')
    model = _small_model(emu_library, True)
    thetas = draw_walkers_fast(model, 1100, seed=12)
    holder = model.native_sampler_plan(thetas[:16])
    want = model.engine.lnpost(holder['plan'], thetas)       # (threads exist from here on)
    pid = os.fork()
    if pid == 0:
        code = 1
        try:
            got = model.engine.lnpost(holder['plan'], thetas)
            code = 0 if np.array_equal(got, want) else 2
        finally:
            os._exit(code)
    deadline = time.time() + 120
    while time.time() < deadline:
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0, status
            break
        time.sleep(0.05)
    else:
        os.kill(pid, 9)
        os.waitpid(pid, 0)
        raise AssertionError('the forked child hung in the library')
    assert np.array_equal(model.engine.lnpost(holder['plan'], thetas), want)


def test_ensemble_run_rejects_bad_input(emu_library):
    from psfmc_b200 import _lib
    model = _small_model(emu_library, False)
    engine = model.engine
    ndim = model.num_params
    key, pos = _state_arrays(np.random.RandomState(1))
    p = np.zeros((4, ndim))
    lnp = np.zeros(4)
    with pytest.raises(_lib.EngineError):     # odd ensemble
        engine.ensemble_run(None, np.zeros((3, ndim)), np.zeros(3), key, pos, 1)
    with pytest.raises(_lib.EngineError):     # rows shorter than the program's theta
        engine.ensemble_run(None, np.zeros((4, ndim - 1)), lnp, key, pos, 1)
    with pytest.raises(_lib.EngineError):
        engine.ensemble_run(None, p, lnp, key, pos, 1, a=1.0)
    with pytest.raises(ValueError):           # not updated in place: refused
        engine.ensemble_run(None, p.astype(np.float32), lnp, key, pos, 1)
    # a proposal with an infinite coordinate is emcee's ValueError
    p = np.ones((4, ndim))
    p[0, 0] = 1e308
    p[2, 0] = -1e308
    p[3, 0] = -1e308
    with pytest.raises(ValueError, match='infinite'):
        engine.ensemble_run(None, p, lnp, key, pos, 3)


def test_device_loop_is_not_chosen_for_a_model_float32_cannot_hold(emu_library, monkeypatch):
    """The device loop has no float64 repeat. BatchPool probes the starting ensemble once:
    a model whose walkers need the repeat (a point source 10^7 times brighter than the
    pixel noise) keeps its loop on the host, with a warning; an ordinary model gets the
    device loop. PSFMC_DEVICE_LOOP=1 forces it either way."""
    import warnings

    from psfmc_b200 import BatchPool, MultiComponentModel
    from psfmc_b200.components import PointSource
    from psfmc_b200.distributions import Uniform
    from psfmc_b200.synthetic import synthetic_components
    monkeypatch.setenv('PSFMC_NATIVE_SAMPLER', '1')
    monkeypatch.delenv('PSFMC_DEVICE_LOOP', raising=False)
    size = 32
    centre, box = np.array((size / 2.0, size / 2.0)), np.array((3.0, 3.0))

    def model_with(point_mag):
        comps = synthetic_components(size, 1, dtype=np.float64, psf_size=16)
        comps = [c for c in comps if not isinstance(c, PointSource)]
        comps.append(PointSource(xy=Uniform(loc=centre - box, scale=2 * box),
                                 mag=Uniform(loc=point_mag, scale=0.5)))
        return MultiComponentModel(comps, precision='fp32', library=emu_library)

    ordinary = model_with(21.0)
    start = ordinary.init_params_from_priors(520)
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        pool = BatchPool(ordinary)
        engine, holder = pool.native_sampler(start)
        assert holder.get('device_loop') is True
        assert pool.native_sampler(start)[1].get('device_loop') is True     # cached
    assert ordinary.engine.info()['rescued_total'] == 0

    assert 1e3 < ordinary.float32_dynamic_range() < 1e5
    with pytest.warns(UserWarning, match='fp64'):      # the advice at construction
        bright = model_with(8.0)        # 10^(0.4 * 17.9) = 1.4e7 ADU on 0.02 ADU of noise
    assert bright.float32_dynamic_range() > 1e8
    start = bright.init_params_from_priors(520)
    oracle = oracle_from_model(bright)
    assert np.all(np.isfinite(oracle.lnlike_batch(start[:8])))
    pool = BatchPool(bright)
    with pytest.warns(UserWarning, match='float64 repeat'):
        engine, holder = pool.native_sampler(start)
    assert not holder.get('device_loop')
    assert bright.engine.info()['rescued_total'] > 5
    with warnings.catch_warnings():
        warnings.simplefilter('error')                                       # warned once
        assert not pool.native_sampler(start)[1].get('device_loop')
    monkeypatch.setenv('PSFMC_DEVICE_LOOP', '1')
    assert BatchPool(bright).native_sampler(start)[1].get('device_loop') is True
    # small ensembles never ask
    monkeypatch.delenv('PSFMC_DEVICE_LOOP')
    fresh = BatchPool(bright)
    assert not fresh.native_sampler(start[:100])[1].get('device_loop')
    assert fresh._fp32_enough is None


def _random_model(rng, library):
    """A random small model: constants and priors mixed at random over every parameter,
    Uniform / Normal columns (library), Gamma / WeibullMinimum ones (callback into scipy
    under PSFMC_PRIORS_STRICT), bilinear and Lanczos point sources, angles in degrees and
    radians -- everything the prior plan and the propose / accept kernels branch on."""
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Configuration, PointSource, Sersic, Sky
    from psfmc_b200.distributions import Gamma, Normal, Uniform, WeibullMinimum
    from psfmc_b200.synthetic import synthetic_psf
    size = 16
    obs = 0.02 * rng.standard_normal((size, size))
    ivm = np.full((size, size), 2500.0)
    psf, psf_ivm = synthetic_psf(8)
    comps = [Configuration(obs_file=obs, obsivm_file=ivm, psf_files=psf, psfivm_files=psf_ivm,
                           mag_zeropoint=25.0)]

    def maybe(prior, constant):
        return prior if rng.rand() < 0.7 else constant

    def position():
        lo = rng.uniform(3.0, 9.0, 2)
        return maybe(Uniform(loc=lo, scale=rng.uniform(2.0, 4.0, 2)), tuple(lo + 1.3))

    def magnitude():
        kind = rng.randint(3)
        if kind == 0:
            return Uniform(loc=rng.uniform(20, 22), scale=rng.uniform(1, 3))
        if kind == 1:
            return Normal(loc=rng.uniform(21, 23), scale=rng.uniform(0.2, 1.0))
        return float(rng.uniform(21, 23))
    comps.append(Sky(adu=maybe(Normal(loc=0, scale=0.01), 0.002)))
    for _ in range(rng.randint(0, 3)):
        comps.append(PointSource(xy=position(), mag=magnitude(),
                                 shift_method=('bilinear', 'lanczos3')[rng.randint(2)]))
    for _ in range(rng.randint(1, 3)):
        degrees = bool(rng.randint(2))
        span = 180.0 if degrees else np.pi
        index = (Uniform(loc=0.5, scale=5.0), WeibullMinimum(c=1.5, scale=4),
                 Gamma(a=2.0, scale=1.2), 1.7)[rng.randint(4)]
        comps.append(Sersic(xy=position(), mag=magnitude(),
                            reff=maybe(Uniform(loc=2.0, scale=rng.uniform(2, 5)), 3.9),
                            reff_b=maybe(Uniform(loc=1.0, scale=rng.uniform(1.5, 4)), 1.9),
                            index=index,
                            angle=maybe(Uniform(loc=0, scale=span), 0.3 * span),
                            angle_degrees=degrees))
    return MultiComponentModel(comps, precision='fp64', library=library)


@pytest.mark.parametrize('seed', range(8))
def test_library_loops_on_random_models(emu_library, monkeypatch, seed):
    """Randomised models (see _random_model): the host loop and the device loop of
    psfmc_ensemble_run continue the numpy loop's chain bit for bit -- positions,
    lnprobability, acceptance counts, the random stream afterwards. (Strict priors: a
    Weibull / Gamma column goes through scipy, and a plan with such a column keeps the loop
    on the host.)"""
    from psfmc_b200.synthetic import draw_walkers_fast
    monkeypatch.setenv('PSFMC_PRIORS_STRICT', '1')
    model = _random_model(np.random.RandomState(100 + seed), emu_library)
    if model.num_params == 0:
        pytest.skip('every parameter came out constant')
    nwalk = 2 * model.num_params + 2
    start = draw_walkers_fast(model, nwalk, seed=seed)
    start = start[0] + 0.05 * (start - start[0])
    assert np.any(np.isfinite(model.log_posterior_batch(start)))
    ref = _run(model, start, False, monkeypatch, iterations=6)
    for device in (False, True):
        got = _run(model, start, True, monkeypatch, iterations=6, device=device)
        assert model._sampler_plan, 'the library loop was not used'
        for key in ('chain', 'lnprobability', 'naccepted', 'next', 'pos', 'lnprob'):
            assert np.array_equal(got[key], ref[key]), (key, device)
    assert 0 < ref['naccepted'].sum()


def test_library_loops_follow_any_call_pattern(emu_library, monkeypatch):
    """Random sequences of run_mcmc / sample calls with random iteration counts (zero
    included), thin factors (iterations need not be a multiple: the last sample is kept,
    where emcee 2.x fails with an IndexError) and storechain on / off: the host and the
    device loop leave the chain, lnprobability, acceptance counts, the iteration counter
    and the random stream of the numpy loop."""
    from psfmc_b200 import BatchPool
    from psfmc_b200.sampler import EnsembleSampler
    from psfmc_b200.synthetic import draw_walkers_fast
    model = _small_model(emu_library, False)
    nwalk = 2 * model.num_params + 2
    start = draw_walkers_fast(model, nwalk, seed=3)
    start = start[0] + 0.02 * (start - start[0])

    def run(native, device, calls):
        monkeypatch.setenv('PSFMC_NATIVE_SAMPLER', '1' if native else '0')
        monkeypatch.setenv('PSFMC_DEVICE_LOOP', '1' if device else '0')
        smp = EnsembleSampler(nwalk, model.num_params, model.log_posterior,
                              kwargs={'model': model}, pool=BatchPool(model))
        smp._random.seed(9)
        pos, lnp = start, None
        for its, thin, store, how in calls:
            if how == 'run':
                res = smp.run_mcmc(pos, its, lnprob0=lnp, thin=thin, storechain=store)
            else:
                res = None
                for res in smp.sample(pos, lnprob0=lnp, iterations=its, thin=thin,
                                      storechain=store):
                    pass
            if res is not None:
                pos, lnp = res[0], res[1]
        return (smp.chain.copy(), smp.lnprobability.copy(), smp.naccepted.copy(),
                smp.iterations, smp._random.rand(3))

    rng = np.random.RandomState(0)
    stored = 0
    for trial in range(8):
        calls = [(int(rng.randint(0, 7)), int(rng.randint(1, 4)), bool(rng.rand() < 0.8),
                  ('run', 'sample')[rng.randint(2)]) for _ in range(rng.randint(1, 4))]
        ref = run(False, False, calls)
        assert ref[0].shape[1] == sum(-(-its // thin) for its, thin, store, _ in calls
                                      if store)
        stored += ref[0].shape[1]
        for device in (False, True):
            got = run(True, device, calls)
            for mine, want in zip(got, ref):
                assert np.array_equal(mine, want), (trial, calls, device)
    assert stored > 10
