"""
Parity tests proper: the CUDA path on a real B200, called through the C ABI
(ctypes -> libpsfmc_b200.so), against the golden vectors produced by the
reference (tests/golden/make_golden.py) and against the oracle on seeded inputs.
"""
import os

import numpy as np
import pytest

from conftest import (assert_lnl_close, check_fused_images, check_pssub_golden, fp32_bounds,
                      load_golden,
                      mixed_model_128, model_from_file, oracle_from_model)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c1_golden():
    return load_golden('c1_golden.json')


def test_library_is_the_cuda_one(cuda_library):
    from psfmc_b200 import _lib
    assert _lib.library_path() == cuda_library
    import torch
    assert torch.cuda.is_available()


def test_c1_fp64_matches_reference_m3(cuda_library, c1_golden):
    """FP64 mode vs the reference's own lnL on float64 inputs: 1e-10 relative."""
    model = model_from_file('j0005/model_c1.py', 'fp64', obs_dtype=np.float64)
    thetas = np.array(c1_golden['theta'])
    got = model.log_likelihood_batch(thetas)
    assert_lnl_close(got, c1_golden['lnl']['M3'], 'fp64')
    # the -inf cases are exactly the reference's
    assert got[c1_golden['names'].index('C_exact_centre')] == -np.inf


def test_c1_fp32_within_stated_tolerance(cuda_library, c1_golden):
    model = model_from_file('j0005/model_c1.py', 'fp32', obs_dtype=np.float64)
    thetas = np.array(c1_golden['theta'])
    got = model.log_likelihood_batch(thetas)
    bounds = fp32_bounds(model, thetas)
    assert_lnl_close(got, c1_golden['lnl']['M3'], 'fp32', bounds)
    # and against the numpy-1.x-faithful mode of the reference (float32 FITS)
    model32 = model_from_file('j0005/model_c1.py', 'fp32')
    got32 = model32.log_likelihood_batch(thetas)
    assert_lnl_close(got32, c1_golden['lnl']['M2'], 'fp32', bounds)


def test_c1_fp32_absolute_tolerance(cuda_library):
    """The absolute statement that goes with the per-walker float32 bound:
    on a seeded prior-drawn C1 ensemble (far from the posterior: lnL down to -2e6) the
    float32 engine agrees with the float64 engine -- itself gated at 1e-10 against the
    reference -- within  |dlnL| <= max(0.08, 8e-6 |lnL|),  with the same walkers
    non-finite. (65536-walker audit, profiles/r2_fp32_large_audit.json: worst 0.054 for
    |lnL| <= 5e4, worst relative 4.3e-6 beyond.)"""
    from psfmc_b200.synthetic import draw_walkers_fast
    m64 = model_from_file('j0005/model_c1.py', 'fp64')
    m32 = model_from_file('j0005/model_c1.py', 'fp32')
    thetas = draw_walkers_fast(m64, 4096, seed=314)
    l64, l32 = m64.log_likelihood_batch(thetas), m32.log_likelihood_batch(thetas)
    finite = np.isfinite(l64)
    assert np.array_equal(np.isfinite(l32), finite)
    err = np.abs(l32 - l64)[finite]
    bound = np.maximum(0.08, 8e-6 * np.abs(l64[finite]))
    worst = np.argmax(err / bound)
    assert np.all(err <= bound), (err[worst], l64[finite][worst])
    assert np.median(err) < 0.05


def test_c1_fused_and_staged_paths_agree(cuda_library, c1_golden, monkeypatch):
    """128^2 float32 runs the fused shared-memory kernel; forcing the staged
    row/column kernels must give the same lnL within the float32 tolerance."""
    thetas = np.array(c1_golden['theta'])
    fused = model_from_file('j0005/model_c1.py', 'fp32', obs_dtype=np.float64)
    assert fused.engine.info()['path'] == 1
    got_fused = fused.log_likelihood_batch(thetas)
    monkeypatch.setenv('PSFMC_FORCE_STAGED', '1')
    staged = model_from_file('j0005/model_c1.py', 'fp32', obs_dtype=np.float64)
    assert staged.engine.info()['path'] == 0
    got_staged = staged.log_likelihood_batch(thetas)
    bounds = fp32_bounds(fused, thetas)
    assert_lnl_close(got_fused, c1_golden['lnl']['M3'], 'fp32', bounds)
    assert_lnl_close(got_staged, c1_golden['lnl']['M3'], 'fp32', bounds)
    assert_lnl_close(got_fused, got_staged, 'fp32', 2 * bounds)


def test_c1_rawf32_mode_tracks_reference_m2(cuda_library, c1_golden):
    """float32 raw-model storage + float64 FFT/tail = the reference on numpy 1.x
    with float32 FITS inputs. A float32 rounding may flip where the float64 value
    differs in the last ulp, so the gate is 1e-7 relative, not 1e-10."""
    model = model_from_file('j0005/model_c1.py', 'fp64_rawf32')
    thetas = np.array(c1_golden['theta'])
    got = model.log_likelihood_batch(thetas)
    expect = np.array(c1_golden['lnl']['M2'])
    finite = np.isfinite(expect)
    assert np.array_equal(np.isfinite(got), finite)
    rel = np.abs(got[finite] - expect[finite]) / np.abs(expect[finite])
    assert rel.max() < 1e-7, rel.max()


def test_c1_two_psf_index_selection(cuda_library):
    golden = load_golden('c1_2psf_golden.json')
    thetas = np.array(golden['theta'])
    model = model_from_file('j0005/model_c1_2psf.py', 'fp64', obs_dtype=np.float64,
                            two_psf=True)
    assert model.num_params == golden['setup']['num_params']
    assert_lnl_close(model.log_likelihood_batch(thetas), golden['lnl']['M3'], 'fp64')
    model32 = model_from_file('j0005/model_c1_2psf.py', 'fp32', obs_dtype=np.float64,
                              two_psf=True)
    assert_lnl_close(model32.log_likelihood_batch(thetas), golden['lnl']['M3'], 'fp32',
                     fp32_bounds(model32, thetas))
    # an out-of-range PSF index is -inf (the prior is -inf there)
    bad = thetas[:2].copy()
    bad[:, -1] = (2.6, -0.7)
    assert np.all(model.log_likelihood_batch(bad) == -np.inf)


@pytest.mark.parametrize('index', ['0.5', '1.0', '3.1', '4.0', '6.5'])
def test_c2_galfit_sersic_sweep(cuda_library, index):
    golden = load_golden('c2_golden.json')['cases'][index]
    thetas = np.array(golden['theta'])
    model = model_from_file(golden['model_file'], 'fp64', obs_dtype=np.float64)
    assert_lnl_close(model.log_likelihood_batch(thetas), golden['lnl']['M3'], 'fp64')
    model32 = model_from_file(golden['model_file'], 'fp32', obs_dtype=np.float64)
    got = model32.log_likelihood_batch(thetas)
    expect = np.array(golden['lnl']['M3'])
    # noiseless fixture with weights 4e6: very high S/N, so the float32 bound
    # (which scales with S/N) is correspondingly larger here
    assert_lnl_close(got, expect, 'fp32', fp32_bounds(model32, thetas))


def test_c1_images_match_reference_pixels(cuda_library, c1_golden):
    """Blob images (raw, convolved, residual, composite IVM) at sampled pixels."""
    model = model_from_file('j0005/model_c1.py', 'fp64', obs_dtype=np.float64)
    thetas = np.array(c1_golden['theta'][:6])
    imgs = model.engine.render(thetas)
    px = np.array(c1_golden['sample_px'])
    for row in range(len(thetas)):
        ref = c1_golden['pixels']['M3'][row]
        for key in ('raw_model', 'convolved_model', 'residual', 'composite_ivm'):
            got = imgs[key][row].ravel()[px]
            want = np.array(ref[key])
            if not np.all(np.isfinite(want)):
                continue
            scale = np.abs(imgs[key][row][np.isfinite(imgs[key][row])]).max()
            assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * scale), (row, key)
    # point-source-subtracted image against the oracle
    oracle = oracle_from_model(model)
    ref = oracle.images(thetas[0])['point_source_subtracted']
    assert np.allclose(imgs['point_source_subtracted'][0], ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_point_source_subtracted_matches_the_reference(cuda_library, precision):
    """Row a13 (psfMC/models.py:296-306) against the unmodified reference's image:
    12 C1 + 8 two-PSF parameter vectors, fp64 (1e-9) and fp32 (2e-6 of the scale)."""
    check_pssub_golden(cuda_library, precision, 'c1')
    check_pssub_golden(cuda_library, precision, 'c1_2psf')


@pytest.mark.parametrize('size,n_sersic', [(256, 2), (512, 3), (64, 1), (32, 1), (16, 1),
                                           (1024, 1)])
def test_synthetic_frames_against_oracle(cuda_library, size, n_sersic):
    """C3 / C4 frame sizes (and small ones) on seeded synthetic inputs."""
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    comps = synthetic_components(size, n_sersic, dtype=np.float64,
                                 psf_size=min(64, size // 2))
    model = MultiComponentModel(comps, precision='fp64')
    thetas = draw_walkers_fast(model, 3 if size >= 1024 else 6, seed=size)
    expect = oracle_from_model(model).lnlike_batch(thetas)
    assert_lnl_close(model.log_likelihood_batch(thetas), expect, 'fp64')
    comps = synthetic_components(size, n_sersic, dtype=np.float64,
                                 psf_size=min(64, size // 2))
    model32 = MultiComponentModel(comps, precision='fp32')
    assert_lnl_close(model32.log_likelihood_batch(thetas), expect, 'fp32',
                     fp32_bounds(model32, thetas))


def test_nonsquare_frame(cuda_library):
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Configuration, PointSource, Sersic, Sky
    rng = np.random.RandomState(3)
    obs = 0.05 * rng.standard_normal((64, 128))
    ivm = np.full((64, 128), 400.0)
    ivm[5, 7] = 0.0
    obs[9, 100] = np.nan
    psf = np.zeros((16, 32))
    psf[6:11, 14:19] = rng.random_sample((5, 5)) + 0.1
    psf_ivm = np.full((16, 32), 1.0e4)
    comps = [Configuration(obs, ivm, psf, psf_ivm, mag_zeropoint=25.0),
             Sky(adu=0.003), PointSource(xy=(70.3, 30.6), mag=19.0),
             Sersic(xy=(60.2, 33.1), mag=20.0, reff=7.0, reff_b=3.0, index=2.2,
                    angle=0.4)]
    model = MultiComponentModel(comps, precision='fp64')
    assert model.num_params == 0
    theta = np.zeros((3, 1))
    expect = oracle_from_model(model).lnlike_batch(np.zeros((3, 0)))
    assert_lnl_close(model.log_likelihood_batch(theta), expect, 'fp64')


def test_full_size_properties(cuda_library):
    """Size-independent properties at the benchmark's full ensemble size:
    determinism, independence of batch composition / chunking, permutation
    equivariance, agreement of the host and device-pointer entry points."""
    import torch
    from psfmc_b200.synthetic import draw_walkers_fast
    model = model_from_file('j0005/model_c1.py', 'fp32')
    thetas = draw_walkers_fast(model, 4096, seed=11)
    full = model.log_likelihood_batch(thetas)
    again = model.log_likelihood_batch(thetas)
    assert np.array_equal(full, again)
    perm = np.random.RandomState(0).permutation(len(thetas))
    assert np.array_equal(model.log_likelihood_batch(thetas[perm]), full[perm])
    pieces = np.concatenate([model.log_likelihood_batch(thetas[s:s + 333])
                             for s in range(0, len(thetas), 333)])
    assert np.array_equal(pieces, full)
    assert np.isfinite(full).mean() > 0.9
    # device-pointer entry point on torch's current stream
    dev = torch.device('cuda:0')
    th_d = torch.from_numpy(thetas).to(dev)
    out_d = torch.empty(len(thetas), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    model.engine.lnlike_device(th_d.data_ptr(), len(thetas), thetas.shape[1],
                               out_d.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    # (the device-pointer call returns the raw float32 results: the few walkers the
    # float32 kernels give up on are repeated in float64 by the host call only)
    got = out_d.cpu().numpy()
    kept = np.isfinite(got)
    assert np.array_equal(got[kept], full[kept])
    assert np.all(got[~kept] == -np.inf) and (~kept).sum() <= 8
    assert (~kept).sum() - (~np.isfinite(full)).sum() <= model.engine.info()['rescued_total']
    # a sample of the full batch against the oracle
    oracle = oracle_from_model(model)
    rows = np.arange(0, 4096, 256)
    assert_lnl_close(full[rows], oracle.lnlike_batch(thetas[rows]), 'fp32',
                     fp32_bounds(model, thetas[rows], oracle))


def test_edge_batches(cuda_library, c1_golden):
    model = model_from_file('j0005/model_c1.py', 'fp32')
    thetas = np.array(c1_golden['theta'])
    assert model.log_likelihood_batch(thetas[:0].reshape(0, 18)).shape == (0,)
    one = model.log_likelihood_batch(thetas[:1])
    assert one.shape == (1,) and np.isfinite(one[0])
    # ld larger than D (padded rows) is allowed
    padded = np.concatenate([thetas, np.zeros((len(thetas), 3))], axis=1)
    assert np.array_equal(model.log_likelihood_batch(padded),
                          model.log_likelihood_batch(thetas))


def test_multi_device_sharding(cuda_library, c1_golden):
    """In-process device list (psfmc_desc.devices): a batch split over all GPUs of the
    box -- one host thread per device enqueues its shard -- gives bit-identical lnL,
    images and image sums (up to the order of the per-device partial sums)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    from psfmc_b200.synthetic import draw_walkers_fast
    ndev = torch.cuda.device_count()
    single = model_from_file('j0005/model_c1.py', 'fp32')
    multi = model_from_file('j0005/model_c1.py', 'fp32', devices=list(range(ndev)))
    assert multi.engine.info()['n_devices'] == ndev
    thetas = draw_walkers_fast(single, 1001, seed=5)
    want = single.log_likelihood_batch(thetas)
    for _ in range(3):      # repeated calls: the per-device worker threads are reused
        assert np.array_equal(want, multi.log_likelihood_batch(thetas))
    # fewer rows than devices, a single row
    assert np.array_equal(want[:1], multi.log_likelihood_batch(thetas[:1]))
    assert np.array_equal(want[:ndev - 1], multi.log_likelihood_batch(thetas[:ndev - 1]))
    # split call with the priors overlapped (what BatchPool / ShardedPool use)
    assert np.array_equal(single.log_posterior_batch(thetas),
                          multi.log_posterior_batch(thetas))
    # blob images and their sums, rows split over the devices
    rows = np.array(c1_golden['theta'])
    rows = rows[np.isfinite(np.array(c1_golden['lnl']['M3']))][:7]
    a, b = single.engine.render(rows), multi.engine.render(rows)
    for key in a:
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    sa, sb = single.engine.accumulate(rows), multi.engine.accumulate(rows)
    for key in sa:
        scale = np.abs(sa[key][np.isfinite(sa[key])]).max()
        assert np.allclose(sa[key], sb[key], rtol=1e-12, atol=1e-12 * scale,
                           equal_nan=True), key
    # a short theta is rejected on the multi-device engine too, and it stays usable
    with pytest.raises(ValueError):
        multi.engine.lnlike(thetas[:, :17])
    assert np.array_equal(want, multi.log_likelihood_batch(thetas))
    # the sampler loop inside the library (psfmc_ensemble_run) over the device list:
    # the chain of the one-device engine
    from psfmc_b200 import BatchPool
    from psfmc_b200.sampler import EnsembleSampler
    centre = np.array(c1_golden['theta'][0])
    start = centre + 1e-3 * np.random.RandomState(3).standard_normal((600, len(centre))) * \
        np.maximum(np.abs(centre), 1.0)
    chains = []
    for model in (single, multi):
        sampler = EnsembleSampler(600, len(centre), model.log_posterior,
                                  kwargs={'model': model}, pool=BatchPool(model))
        sampler._random.seed(4)
        sampler.run_mcmc(start, 10)
        assert model._sampler_plan
        chains.append((sampler.chain.copy(), sampler.lnprobability.copy()))
    assert np.array_equal(chains[0][0], chains[1][0])
    assert np.array_equal(chains[0][1], chains[1][1])


def test_multi_process_sharded_pool_nccl(cuda_library, tmp_path):
    """One process per GPU under torchrun (the deployment bench.py --gpus N times):
    ShardedPool over NCCL returns on every rank exactly what one engine returns for
    the whole batch, ragged shards included."""
    import subprocess
    import sys
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip('needs >= 2 GPUs')
    script = tmp_path / 'worker.py'
    script.write_text("""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import torch, torch.distributed as dist
from conftest import model_from_file
from psfmc_b200.distributed import ShardedPool
from psfmc_b200.synthetic import draw_walkers_fast
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
model = model_from_file('j0005/model_c1.py', 'fp32', devices=[local])
thetas = draw_walkers_fast(model, 1001, seed=5)
pool = ShardedPool(model)
got, _ = pool.map_batch(None, thetas)
rows = [thetas[i] for i in range(37)]
listed = np.array([r[0] for r in pool.map(None, rows)])
# lnL gather over peer memory (CUDA IPC mailboxes): three calls (both mailbox halves),
# ragged shards, against the same engine evaluating the whole batch alone
from psfmc_b200.distributed import PeerExchange
raw = model_from_file('j0005/model_c1.py', 'fp32', devices=[local], fp64_rescue=False)
xch = PeerExchange(raw.engine, 1001)
th_dev = torch.from_numpy(thetas).cuda()
stream = torch.cuda.current_stream()
peer = []
for count in (1001, 64, 999):
    gathered = torch.full((1001,), 7.0, dtype=torch.float64, device='cuda')
    xch.lnlike(th_dev, count, thetas.shape[1], gathered, stream)
    stream.synchronize()
    peer.append(gathered.cpu().numpy()[:count])
alone = raw.log_likelihood_batch(thetas)
# the same gather with host buffers in one library call (psfmc_lnpost_batch_sharded)
host_sharded = raw.engine.lnpost_sharded(None, thetas)
host_sharded_37 = raw.engine.lnpost_sharded(None, thetas[:37])
# the sampler's loop inside the library, sharded (psfmc_ensemble_run + PSFMC_ENS_SHARDED:
# every rank the same seeded loop, its share of every half-ensemble, lnL over peer
# memory) against the same loop on this rank's engine alone; 250 walkers (ragged shards
# of 125 rows) and 1000
from psfmc_b200 import BatchPool
from psfmc_b200.sampler import EnsembleSampler
import json
centre = np.array(json.load(open(os.path.join({root!r}, 'tests', 'golden',
                                              'c1_golden.json')))['theta'][0])
chains = {{}}
for nwalk in (250, 1000):
    start = centre + 1e-3 * np.random.RandomState(3).standard_normal((nwalk, len(centre))) * \
        np.maximum(np.abs(centre), 1.0)
    for name, pl in (('sharded', ShardedPool(model)), ('single', BatchPool(model))):
        smp = EnsembleSampler(nwalk, len(centre), model.log_posterior,
                              kwargs={{'model': model}}, pool=pl)
        smp._random.seed(4)
        smp.run_mcmc(start, 8)
        if name == 'sharded':
            assert pl.native_sampler(start) is not None, 'the sharded library loop was not used'
        chains['%s_%d' % (name, nwalk)] = smp.chain.copy()
        chains['%s_lnp_%d' % (name, nwalk)] = smp.lnprobability.copy()
np.savez(os.path.join({out!r}, 'rank%d.npz' % dist.get_rank()), got=got, listed=listed,
         want=model.log_posterior_batch(thetas), alone=alone, peer0=peer[0],
         peer1=peer[1], peer2=peer[2], host_sharded=host_sharded,
         host_sharded_37=host_sharded_37, **chains)
dist.barrier()
dist.destroy_process_group()
""".format(root=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
           out=str(tmp_path)))
    nproc = min(ndev, 8)
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                    '--nproc-per-node', str(nproc), '--master-addr', '127.0.0.1',
                    '--master-port', '29533', str(script)], check=True, timeout=600)
    first = np.load(str(tmp_path / 'rank0.npz'))
    for r in range(nproc):
        data = np.load(str(tmp_path / 'rank{}.npz'.format(r)))
        assert np.array_equal(data['got'], data['want'])
        assert np.array_equal(data['got'], first['got'])
        assert np.array_equal(data['listed'], data['want'][:37])
        assert np.array_equal(data['peer0'], data['alone'])
        assert np.array_equal(data['peer1'], data['alone'][:64])
        assert np.array_equal(data['peer2'], data['alone'][:999])
        assert np.array_equal(data['host_sharded'], data['alone'])
        assert np.array_equal(data['host_sharded_37'], data['alone'][:37])
        for nwalk in (250, 1000):
            # (positions bit for bit; lnprob to the last digits: the sharded loop runs on
            # the devices, whose log / pow serve the Weibull priors, the single-engine loop
            # of 250 walkers on the host)
            key = '%d' % nwalk
            assert np.array_equal(data['sharded_' + key], data['single_' + key]), key
            assert np.array_equal(data['sharded_' + key], first['sharded_' + key]), key
            key = 'lnp_%d' % nwalk
            np.testing.assert_allclose(data['sharded_' + key], data['single_' + key],
                                       rtol=1e-13)
            assert np.array_equal(data['sharded_' + key], first['sharded_' + key]), key


@pytest.mark.parametrize('table', ['1', '0'])
def test_device_kappa_matches_scipy(cuda_library, table, monkeypatch):
    """kappa = gammaincinv(2n, 0.5), Gamma(2n) and sb_eff computed on the device
    (Chebyshev table built at engine creation, or the group-cooperative Halley
    iteration with approximate float64 reciprocals) against scipy over the whole
    prior range: the raw model at r = reff is sb_eff * (1 + corr)."""
    monkeypatch.setenv('PSFMC_NO_KAPPA_TABLE', '0' if table == '1' else '1')
    from scipy.special import gammaincinv
    from oracle import psfmc_oracle as orc
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Configuration, Sersic
    from psfmc_b200.distributions import Uniform
    size = 32
    psf = np.zeros((8, 8))
    psf[4, 4] = 1.0
    comps = [Configuration(np.zeros((size, size)), np.ones((size, size)), psf,
                           np.full((8, 8), 1e12), mag_zeropoint=25.0),
             Sersic(xy=(10.0, 16.0), mag=20.0, reff=6.0, reff_b=6.0,
                    index=Uniform(loc=0.01, scale=50), angle=0.0)]
    model = MultiComponentModel(comps, precision='fp64')
    assert model.engine.info()['kappa_table'] == int(table)
    ns = np.concatenate([[0.03, 0.06, 0.13, 0.2, 0.36, 0.5, 0.75, 1.0, 1.7, 2.5, 4.0, 6.5,
                          9.9, 15.0, 19.5, 31.9, 40.0],
                         np.random.RandomState(4).uniform(0.1, 12, 500)])
    raw = model.engine.render(ns[:, None], which=('raw_model',))['raw_model']
    kappa = gammaincinv(2 * ns, 0.5)
    sbeff = orc.sersic_sb_eff(orc.mag_to_flux(20.0, 25.0), ns, 6.0, 6.0, kappa)
    grad = -kappa / ns
    want = sbeff * (1 + grad * (1 / 36.0 / 12 * grad))
    assert np.max(np.abs(raw[:, 16, 16] / want - 1)) < 1e-12


def test_seeded_run_posterior_medians_within_mc_error(cuda_library):
    """north_star: posterior medians of a seeded run agree within Monte-Carlo error.
    A synthetic observation is made from a known parameter vector plus noise; the
    same seeded stretch-move sampler is then driven (a) by the GPU engine through
    BatchPool (float32 mode) and (b) by the oracle (CPU restatement of the
    reference); the posterior medians of the two runs are compared in units of the
    Monte-Carlo error, and both must recover the truth."""
    from psfmc_b200 import BatchPool, MultiComponentModel
    from psfmc_b200.components import Configuration
    from psfmc_b200.sampler import EnsembleSampler
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    size = 64
    comps = synthetic_components(size, 1, psf_size=32, mask_disc=False)
    probe = MultiComponentModel(comps, precision='fp64')
    truth = draw_walkers_fast(probe, 1, seed=5)[0]
    image = probe.engine.render(truth[None, :], which=('convolved_model',))
    rng = np.random.RandomState(8)
    obs = (image['convolved_model'][0] + 0.02 * rng.standard_normal((size, size)))
    old = comps[0]
    comps = [Configuration(obs.astype(np.float32), np.full((size, size), 2500.0, np.float32),
                           old.psf_selector.psf_images[0], 1.0 / np.maximum(
                               old.psf_selector.var_images[0], 1e-30),
                           mag_zeropoint=old.mag_zeropoint)] + comps[1:]
    model = MultiComponentModel(comps, precision='fp32')
    ndim = model.num_params
    nwalk = 4 * ndim
    start = truth + 1e-3 * np.random.RandomState(3).standard_normal((nwalk, ndim)) * \
        np.maximum(np.abs(truth), 1.0)
    med_ref, spread = _seeded_runs_agree(model, start, nburn=100, nkeep=200)
    assert np.all(np.abs(med_ref - truth) < 6.0 * spread + 1e-9)


def _seeded_runs_agree(model, start, nburn, nkeep):
    """The same seeded stretch-move run driven by the GPU engine (BatchPool, float32
    mode) and by the oracle; returns the oracle run's medians and spreads after
    checking that the medians of the two runs agree within Monte-Carlo error."""
    from psfmc_b200 import BatchPool
    from psfmc_b200.sampler import EnsembleSampler
    oracle = oracle_from_model(model)
    nwalk, ndim = start.shape

    class OraclePool(object):
        def map(self, func, items):
            thetas = np.stack([np.asarray(p) for p in items])
            lnprior = model.log_priors_batch(thetas)
            out = []
            for theta, lp in zip(thetas, lnprior):
                lnl = oracle.lnlike(theta) if np.isfinite(lp) else -np.inf
                good = np.isfinite(lnl) and np.isfinite(lp)
                out.append((lnl + lp if good else -np.inf, {}))
            return out

    chains = {}
    for name, pool in (('gpu', BatchPool(model)), ('oracle', OraclePool())):
        sampler = EnsembleSampler(nwalk, ndim, model.log_posterior,
                                  kwargs={'model': model}, pool=pool)
        sampler._random.seed(99)
        pos = sampler.run_mcmc(start, nburn)[0]
        sampler.reset()
        sampler.run_mcmc(pos, nkeep)
        chains[name] = sampler.flatchain
        assert 0.05 < sampler.acceptance_fraction.mean() < 0.9
    med_gpu = np.median(chains['gpu'], axis=0)
    med_ref = np.median(chains['oracle'], axis=0)
    spread = np.std(chains['oracle'], axis=0)
    # The two runs share seed and proposals until a float32 rounding flips one
    # accept decision; afterwards they are independent draws from the same posterior,
    # so medians agree within a few Monte-Carlo errors ~ spread / sqrt(n_eff).
    n_eff = nwalk * nkeep / 50.0
    assert np.all(np.abs(med_gpu - med_ref) < 5.0 * spread / np.sqrt(n_eff) + 1e-9), \
        (np.abs(med_gpu - med_ref) / spread)
    return med_ref, spread


def test_seeded_run_on_the_j0005_model(cuda_library, c1_golden):
    """The same comparison on a BASELINE configuration: the J0005-0006 example model
    (C1, 128 x 128, D = 18), a short chain of 40 walkers started in a tight ball around
    golden vector A (oracle: ~5 ms per evaluation, 4800 evaluations)."""
    model = model_from_file('j0005/model_c1.py', 'fp32')
    centre = np.array(c1_golden['theta'][0])
    nwalk = 40
    rng = np.random.RandomState(17)
    start = centre + 1e-3 * rng.standard_normal((nwalk, len(centre))) * \
        np.maximum(np.abs(centre), 1.0)
    assert np.all(np.isfinite(model.log_priors_batch(start)))
    _seeded_runs_agree(model, start, nburn=40, nkeep=80)


@pytest.mark.parametrize('strict,device', [('1', '0'), ('0', '0'), ('0', '1')])
def test_library_sampler_loop_reproduces_the_numpy_loop(cuda_library, c1_golden, monkeypatch,
                                                        strict, device):
    """psfmc_ensemble_run (the sampler's iterations inside the library) on the J0005-0006
    model, float32 engine: same seed, same start -> the chain of the numpy loop
    (sampler.py), bit for bit with the Weibull priors through the callback (strict), with
    identical positions when they are evaluated in the library with the C library's pow."""
    from psfmc_b200 import BatchPool
    from psfmc_b200.sampler import EnsembleSampler
    # device = '1': proposals, priors and acceptance in kernels around the lnL kernels
    # (PSFMC_ENS_DEVICE): no host round trip per half-ensemble
    monkeypatch.setenv('PSFMC_PRIORS_STRICT', strict)
    monkeypatch.setenv('PSFMC_DEVICE_LOOP', device)
    model = model_from_file('j0005/model_c1.py', 'fp32')
    centre = np.array(c1_golden['theta'][0])
    nwalk = 250
    rng = np.random.RandomState(23)
    start = centre + 1e-3 * rng.standard_normal((nwalk, len(centre))) * \
        np.maximum(np.abs(centre), 1.0)
    runs = {}
    for native in ('0', '1'):
        monkeypatch.setenv('PSFMC_NATIVE_SAMPLER', native)
        sampler = EnsembleSampler(nwalk, len(centre), model.log_posterior,
                                  kwargs={'model': model}, pool=BatchPool(model))
        sampler._random.seed(11)
        pos, lnp, _ = sampler.run_mcmc(start, 12)
        sampler.run_mcmc(pos, 12, lnprob0=lnp, thin=3)
        runs[native] = (sampler.chain.copy(), sampler.lnprobability.copy(),
                        sampler.naccepted.copy(), sampler._random.rand(3))
    assert model._sampler_plan and model._sampler_plan['python_columns'] == (strict == '1')
    assert runs['1'][0].shape == (nwalk, 16, len(centre))
    assert 0.05 < runs['0'][2].mean() / 24 < 0.95
    assert np.array_equal(runs['0'][0], runs['1'][0])
    assert np.array_equal(runs['0'][2], runs['1'][2])
    assert np.array_equal(runs['0'][3], runs['1'][3])
    if strict == '1':
        assert np.array_equal(runs['0'][1], runs['1'][1])
    else:
        np.testing.assert_allclose(runs['1'][1], runs['0'][1], rtol=1e-13)


def test_images_from_the_fused_kernel(cuda_library, c1_golden, monkeypatch):
    """Blob images of the float32 engine through the IMAGES instance of the fused kernel,
    several walkers per CTA (700 walkers on 4 CTAs, then on all SMs)."""
    check_fused_images(None, c1_golden, monkeypatch, n_extra=700)
    model = model_from_file('j0005/model_c1.py', 'fp32')
    thetas = np.array(c1_golden['theta'][:8])
    big = np.tile(thetas, (100, 1))
    sums = model.engine.accumulate(big)
    one = model.engine.accumulate(thetas)
    for name in sums:
        ok = np.isfinite(one[name])
        assert np.allclose(sums[name][ok], 100 * one[name][ok], rtol=1e-10), name


def test_accumulate_on_device_matches_rendered_images(cuda_library, c1_golden):
    """Posterior-image sums accumulated on the device (float32 engine) against the
    individually rendered images; the IVM is summed in variance space."""
    model = model_from_file('j0005/model_c1.py', 'fp32')
    thetas = np.array(c1_golden['theta'][:3] + c1_golden['theta'][6:16])
    imgs = model.engine.render(thetas)
    sums = model.engine.accumulate(thetas)
    for name in imgs:
        want = (1 / imgs[name]).sum(axis=0) if name == 'composite_ivm' \
            else imgs[name].sum(axis=0)
        scale = np.abs(want[np.isfinite(want)]).max()
        assert np.allclose(sums[name], want, rtol=1e-6, atol=1e-9 * scale, equal_nan=True), name


def test_gpu_mixed_components_fused_and_fp64(cuda_library):
    """Bilinear + edge-clipped Lanczos point sources, fixed and free Sersic parameters,
    angle in radians, fixed sky, bad pixels: fused float32 kernel and float64 staged
    kernels against the oracle."""
    from psfmc_b200.synthetic import draw_walkers_fast
    model64 = mixed_model_128('fp64')
    thetas = draw_walkers_fast(model64, 64, seed=8)
    expect = oracle_from_model(model64).lnlike_batch(thetas)
    assert np.all(np.isfinite(expect))
    assert_lnl_close(model64.log_likelihood_batch(thetas), expect, 'fp64')
    model32 = mixed_model_128('fp32')
    assert model32.engine.info()['path'] == 1
    assert_lnl_close(model32.log_likelihood_batch(thetas), expect, 'fp32',
                     fp32_bounds(model32, thetas))


@pytest.mark.gpu
def test_gpu_fp64_rescue_of_high_dynamic_range_walkers(cuda_library, c1_golden):
    from conftest import check_fp64_rescue
    check_fp64_rescue(cuda_library, c1_golden)


@pytest.mark.gpu
def test_gpu_fp64_rescue_inside_the_graph(cuda_library, c1_golden, monkeypatch):
    from conftest import check_fp64_rescue_on_device
    check_fp64_rescue_on_device(cuda_library, c1_golden, monkeypatch)


@pytest.mark.gpu
def test_gpu_split_host_call_and_overlapped_priors(cuda_library, c1_golden, monkeypatch):
    """psfmc_lnlike_batch_begin / _end on the GPU (plain launches and graph replay):
    the numbers of the blocking call; log_posterior_batch with the priors evaluated
    between the two halves equals lnL + lnprior of the separate calls."""
    from conftest import HIGH_DYNAMIC_RANGE_THETAS, model_from_file
    from psfmc_b200.synthetic import draw_walkers_fast
    monkeypatch.setenv('PSFMC_NO_HOT_PIXEL', '1')    # (the walker below must need the repeat)
    model = model_from_file('j0005/model_c1.py', 'fp32', library=cuda_library,
                            obs_dtype=np.float64)
    engine = model.engine
    thetas = draw_walkers_fast(model, 700, seed=9)
    want = engine.lnlike(thetas)
    engine.lnlike_begin(thetas)
    lnprior = model.log_priors_batch(thetas)
    assert np.array_equal(engine.lnlike_end(), want)
    assert thetas.shape[0] >= model.overlap_min_batch
    lnpost = model.log_posterior_batch(thetas)
    ok = np.isfinite(lnprior) & np.isfinite(want)
    assert np.array_equal(lnpost[ok], (want + lnprior)[ok]) and np.all(lnpost[~ok] == -np.inf)
    # a walker that needs the float64 repeat arms the graph path for the next calls
    thetas[11] = HIGH_DYNAMIC_RANGE_THETAS[0]
    first = engine.lnlike(thetas)
    assert np.isfinite(first[11]) and np.array_equal(np.delete(first, 11), np.delete(want, 11))
    replays = engine.info()['graph_replays']
    engine.lnlike_begin(thetas)
    assert np.array_equal(engine.lnlike_end(), first)
    lnprior = model.log_priors_batch(thetas)
    ok = np.isfinite(lnprior) & np.isfinite(first)
    assert np.array_equal(model.log_posterior_batch(thetas)[ok], (first + lnprior)[ok])
    info = engine.info()
    assert info['graph_replays'] == replays + 2 and info['rescued_on_device'] == 2


@pytest.mark.gpu
def test_gpu_masked_row_groups_are_skipped_exactly(cuda_library, monkeypatch):
    """Fused 128 kernel: four-row groups without a good pixel skip their inverse
    transform and epilogue -- bit-identical to transforming them (their pixels never
    enter the sum), on the example's disc mask and on a mask that leaves one row group."""
    from conftest import model_from_file
    from psfmc_b200.synthetic import draw_walkers_fast
    skipping = model_from_file('j0005/model_c1.py', 'fp32', library=cuda_library,
                               fp64_rescue=False)
    assert skipping.engine.info()['path'] == 1
    thetas = draw_walkers_fast(skipping, 600, seed=4)
    got = skipping.log_likelihood_batch(thetas)
    monkeypatch.setenv('PSFMC_NO_ROW_SKIP', '1')
    full = model_from_file('j0005/model_c1.py', 'fp32', library=cuda_library,
                           fp64_rescue=False)
    assert np.array_equal(full.log_likelihood_batch(thetas), got)
    monkeypatch.delenv('PSFMC_NO_ROW_SKIP')
    # everything masked except rows 60..63: 31 of the 32 groups drop out
    from conftest import GOLDEN, j0005_arrays
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.components import Configuration
    from psfmc_b200.model_parser import component_list_from_file
    obs, ivm, _, psfs, ivms = j0005_arrays(np.float64)
    mask = np.ones((128, 128), dtype=bool)
    mask[60:64, 5:120] = False
    comps = [c for c in component_list_from_file(os.path.join(GOLDEN, 'j0005/model_c1.py'))
             if not isinstance(c, Configuration)]
    config = Configuration(obs, ivm, psfs, ivms, mask_file=mask, mag_zeropoint=25.9463)
    narrow = MultiComponentModel([config] + comps, precision='fp32', library=cuda_library,
                                 fp64_rescue=False)
    ref64 = MultiComponentModel([config] + comps, precision='fp64', library=cuda_library)
    assert narrow.engine.info()['path'] == 1
    assert int(np.sum(~np.asarray(narrow.config.bad_px, dtype=bool))) == 4 * 115
    l32, l64 = narrow.log_likelihood_batch(thetas[:64]), ref64.log_likelihood_batch(thetas[:64])
    assert_lnl_close(l32, l64, 'fp32', fp32_bounds(narrow, thetas[:64]))


@pytest.mark.gpu
def test_gpu_tiny_sersic_index_is_minus_inf_like_the_reference(cuda_library):
    from conftest import check_tiny_index_walkers
    check_tiny_index_walkers(cuda_library)


@pytest.mark.gpu
def test_gpu_hot_pixel_walkers(cuda_library, c1_golden):
    from conftest import check_hot_pixel_walkers
    check_hot_pixel_walkers(cuda_library, c1_golden)


def test_gpu_fused_near_centre_walkers(cuda_library):
    from conftest import check_near_centre_walkers
    check_near_centre_walkers(cuda_library)


@pytest.mark.gpu
def test_gpu_cluster_kernel_256(cuda_library, monkeypatch):
    from conftest import check_cluster_path_256
    check_cluster_path_256(cuda_library, 300, monkeypatch)


@pytest.mark.gpu
def test_gpu_tiled_path_512(cuda_library, monkeypatch):
    from conftest import check_tiled_path_512
    check_tiled_path_512(cuda_library, 40, monkeypatch)


def test_gpu_tiled_path_512_properties(cuda_library):
    """C4 at full ensemble size: determinism, permutation equivariance, independence of the
    batch composition, agreement with the float64 engine within the absolute statement."""
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    model = MultiComponentModel(synthetic_components(512, 3), precision='fp32',
                                fp64_rescue=False)
    assert model.engine.info()['path'] == 3
    thetas = draw_walkers_fast(model, 600, seed=11)
    full = model.log_likelihood_batch(thetas)
    assert np.array_equal(full, model.log_likelihood_batch(thetas))
    perm = np.random.RandomState(1).permutation(len(thetas))
    assert np.array_equal(full[perm], model.log_likelihood_batch(thetas[perm]))
    assert np.array_equal(full[100:137], model.log_likelihood_batch(thetas[100:137]))
    m64 = MultiComponentModel(synthetic_components(512, 3), precision='fp64')
    l64 = m64.log_likelihood_batch(thetas[:64])
    finite = np.isfinite(l64)
    assert np.array_equal(np.isfinite(full[:64]), finite)
    err = np.abs(full[:64] - l64)[finite]
    assert np.all(err <= np.maximum(0.3, 8e-6 * np.abs(l64[finite]))), err.max()


def test_gpu_cluster_kernel_256_properties(cuda_library):
    """Full C3 ensemble through the cluster kernel: determinism, permutation
    equivariance, batch-composition independence."""
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast, synthetic_components
    model = MultiComponentModel(synthetic_components(256, 2), precision='fp32',
                                fp64_rescue=False)
    assert model.engine.info()['path'] == 2
    thetas = draw_walkers_fast(model, 1024, seed=3)
    first = model.log_likelihood_batch(thetas)
    assert np.array_equal(first, model.log_likelihood_batch(thetas))
    perm = np.random.RandomState(0).permutation(len(thetas))
    assert np.array_equal(model.log_likelihood_batch(thetas[perm]), first[perm])
    assert np.array_equal(model.log_likelihood_batch(thetas[:37]), first[:37])
    assert np.all(np.isfinite(first))


@pytest.mark.gpu
@pytest.mark.parametrize('dims', [(100, 100, 64, 64), (75, 100, 31, 17), (50, 36, 21, 36),
                                  (128, 100, 32, 32), (33, 64, 8, 9), (100, 100, 25, 25),
                                  (65, 64, 64, 64), (97, 120, 32, 9), (150, 150, 64, 64),
                                  (150, 130, 31, 64), (193, 192, 64, 64),
                                  (301, 300, 65, 64), (600, 500, 101, 99)])
def test_gpu_arbitrary_frame_sizes(cuda_library, dims):
    from conftest import check_arbitrary_frame
    check_arbitrary_frame(cuda_library, dims, n_walkers=5 if dims[0] < 400 else 2)


@pytest.mark.gpu
def test_gpu_nan_parameters_give_minus_inf(cuda_library, monkeypatch):
    from conftest import check_nan_propagation
    check_nan_propagation(cuda_library, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['crop100', 'crop75x100'])
def test_gpu_cropped_frames_match_the_reference(cuda_library, tag):
    from conftest import check_cropped_golden
    check_cropped_golden(cuda_library, tag)


@pytest.mark.gpu
def test_gpu_wide_box_fuzz(cuda_library):
    """(Last in the file on purpose: added after the round's last GPU session, its gates are
    what the emulator run of the same vectors needs, times ten.) Parameter vectors far
    outside the priors -- tools/emu_fuzz.py, DESIGN.md 4.5 'outside the priors': the same
    walkers are finite in the engine and in the oracle, float64 agrees to the rounding of a
    float64 transform at that dynamic range, float32 (non-finite results repeated in
    float64) to 1e-3 of |lnL|; in the 'typical' box float64 holds its 1e-10 gate."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'tools'))
    from emu_fuzz import draw
    models = {prec: model_from_file('j0005/model_c1.py', prec, library=cuda_library,
                                    obs_dtype=np.float64) for prec in ('fp64', 'fp32')}
    oracle = oracle_from_model(models['fp64'])
    for box, count in (('typical', 96), ('wide', 96)):
        thetas = draw(np.random.RandomState(17), count, box)
        with np.errstate(all='ignore'):
            expect = oracle.lnlike_batch(thetas)
        finite = np.isfinite(expect)
        for prec, model in models.items():
            got = model.log_likelihood_batch(thetas)
            assert np.array_equal(np.isfinite(got), finite), (box, prec)
            rel = np.abs(got[finite] - expect[finite]) / np.abs(expect[finite])
            if prec == 'fp64':
                assert rel.max() <= (1e-10 if box == 'typical' else 1e-6), (box, rel.max())
            else:
                assert rel.max() <= (1e-4 if box == 'typical' else 1e-3), (box, rel.max())
