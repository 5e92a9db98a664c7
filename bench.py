#!/usr/bin/env python
"""
bench.py -- model lnL evaluations per second (walkers x iterations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c3|c4|s128]
                    [--walkers NW] [--impl ours|reference] [--precision fp32|fp64]

One "step" is one emcee iteration of ONE NW-walker ensemble: the two sequentially
dependent half-ensemble batches emcee 2.x maps per iteration (SURVEY.md section
3.2), i.e. NW lnL evaluations through the hot path. The default workload is the
J0005-0006 quasar+host model (C1 frame, 128^2, Sky + PointSource + 2 Sersic, D=18)
for a 4096-walker ensemble -- the configuration BASELINE.json's north_star quotes
its target on.

Multi-GPU (torchrun, one rank per GPU): STRONG scaling. The one ensemble is sharded
over the N ranks: every rank evaluates its contiguous rows of each half-ensemble
on its own engine and the per-walker lnL of ALL rows reaches every rank INSIDE the
timed region, because the next half-ensemble's proposals depend on it. This is the
pool.map of /root/reference/psfMC/fitting.py:55-58. The gather is the lnL kernel's
own: it stores each result into every rank's mailbox over CUDA-IPC peer memory
(psfmc_b200.distributed.PeerExchange / psfmc_lnlike_batch_exchange, device buffers:
`value`; psfmc_lnpost_batch_sharded, host buffers: `e2e`). The same step with an NCCL
all_gather_into_tensor of B doubles is timed beside it (`with_nccl_gather`,
`e2e.with_nccl_gather`; PSFMC_BENCH_GATHER=nccl makes it the headline), and so is the
replica throughput (every rank its own ensemble, no gather; weak scaling) as
`replicas_weak`.

Prints ONE JSON line (rank 0). `value` = device-resident throughput (theta and
lnL stay in HBM, CUDA events on the launching stream); `e2e` = the same metric
through the C ABI with host buffers (pinned theta in, lnL out, copies inside the
timed region); `e2e.pool_map` = through the emcee-facing list protocol
(BatchPool.map / ShardedPool.map: row views in, (lnpost, blob) tuples out, priors
included); `e2e.sampler_loop` = the WHOLE stretch-move iteration of this package's
sampler (proposals, priors, lnL, acceptance, chain storage) -- `library`: the loop
inside the library (psfmc_ensemble_run; N > 1: sharded over the ranks), `numpy_loop`:
the same sampler with its Python loop, both also at the reference example's 250 walkers.

`--impl reference` and the `cpu_baseline` leg run the UNMODIFIED reference
(oracle/_ref/psfMC, placed there verbatim by oracle/make_ref.py; imported through
oracle/refshim.py) on all host cores: its own MultiComponentModel.log_posterior
(priors + the five blob images incl. point_source_subtracted, psfMC/models.py:193-243)
in a multiprocessing pool with one model per worker, plus the lnL-only variant
(priors and point_source_subtracted patched out) for a like-for-like comparison
with the GPU arm. The oracle port is the fallback when oracle/_ref is missing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'model lnL evals/sec (walkers x iters)'
UNIT = 'evals/s'
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c1', choices=['c1', 'c3', 'c4', 's128'])
    ap.add_argument('--walkers', type=int, default=0,
                    help='size of THE ensemble, sharded over the GPUs '
                         '(default: 4096 for c1/s128/c4, 1024 for c3)')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'fp64'])
    ap.add_argument('--cpu-seconds', type=float, default=12.0,
                    help='budget of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-impl', default='auto', choices=['auto', 'reference', 'port'],
                    help='CPU arm: the unmodified reference from oracle/_ref (default when '
                         'present) or the oracle numpy port')
    return ap.parse_args()


# ------------------------------------------------------------------ workload --

def workload_description(name, walkers):
    desc = {
        'c1': 'C1 J0005-0006 quasar+host, 128x128, Sky+PointSource+2 Sersic, D=18, '
              '9176/16384 px unmasked',
        'c3': 'C3 synthetic 256x256 quasar + two-Sersic host, PSF-variance term, D=18',
        'c4': 'C4 synthetic 512x512 PointSource + three Sersic, D=25',
        's128': 'synthetic 128x128, C1 component structure',
    }[name]
    return '{}; one {}-walker ensemble, 2 half-ensemble batches per step'.format(
        desc, walkers)


def default_walkers(name):
    return {'c1': 4096, 'c3': 1024, 'c4': 4096, 's128': 4096}[name]


def build_components(name):
    """Component list of the workload (built from files in this repository)."""
    from psfmc_b200.model_parser import component_list_from_file
    from psfmc_b200.synthetic import WORKLOADS, synthetic_components
    if name == 'c1':
        return component_list_from_file(os.path.join(GOLDEN, 'j0005', 'model_c1.py'))
    size, n_sersic, _ = WORKLOADS[name]
    return synthetic_components(size, n_sersic)


def workload_model_file(name):
    """The workload as a psfMC model file (what the reference's own model class
    reads): C1 from tests/golden, the synthetic ones written to a scratch directory
    with bit-identical arrays (psfmc_b200.synthetic.write_synthetic_files)."""
    if name == 'c1':
        return os.path.join(GOLDEN, 'j0005', 'model_c1.py')
    import tempfile
    from psfmc_b200.synthetic import WORKLOADS, write_synthetic_files
    size, n_sersic, _ = WORKLOADS[name]
    return write_synthetic_files(size, n_sersic, tempfile.mkdtemp(prefix='psfmc_bench_'))


def host_model(name):
    """Host-side pieces of the workload without an engine (components, number of
    parameters, batched priors): draws the same walkers as the GPU arm on a box
    whose GPU is not touched (reference arm)."""
    from psfmc_b200.components import Configuration
    from psfmc_b200.program import compile_program
    comps = build_components(name)
    config = [c for c in comps if isinstance(c, Configuration)][0]
    rest = [c for c in comps if c is not config] + [config.psf_selector]
    program, psf_slot, ndim = compile_program(rest)

    class _Shim(object):
        pass
    shim = _Shim()
    shim.components, shim.num_params, shim.config = rest, ndim, config
    shim.program, shim.psf_index_slot = program, psf_slot

    def log_priors_batch(thetas):
        total, start = np.zeros(len(thetas)), 0
        for comp in rest:
            count = comp.num_stochastics()
            total = total + comp.log_priors_batch(thetas[:, start:start + count])
            start += count
        return total
    shim.log_priors_batch = log_priors_batch
    return shim


def ensemble(model, walkers, which):
    """Seeded prior draws; ensemble `which` is the same on every rank and in both
    arms (GPU and CPU)."""
    from psfmc_b200.synthetic import draw_walkers_fast
    return draw_walkers_fast(model, walkers, seed=1000 + which)


# ------------------------------------------------------------------ CPU arm --
# The reference's CPU implementation of the path on the host cores: a
# multiprocessing pool with one model per worker (the reference's own parallel hook
# is emcee's pool.map; its model object is not picklable, psfMC/fitting.py:55,
# hence one model per worker), OMP/MKL/OPENBLAS threads pinned to 1 per worker.

_WORKER = {}


def reference_available():
    from oracle import refshim
    return refshim.reference_available()


def _worker_init(workload, kind, model_file):
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
        os.environ[var] = '1'
    _WORKER['kind'] = kind
    if kind == 'reference':
        from oracle import refshim
        # mode M1: what the unmodified reference computes under this box's numpy
        # (float32 storage, complex64 FFT)
        full = refshim.build_reference_model(model_file, 'M1')
        lnl_only = refshim.build_reference_model(model_file, 'M1')
        lnl_only.log_priors = lambda: 0.0               # psfMC/models.py:187-191
        lnl_only.point_source_subtracted = lambda: None  # psfMC/models.py:296-306
        _WORKER['full'], _WORKER['lnl_only'] = full, lnl_only
        return
    shim = host_model(workload)
    from oracle import psfmc_oracle as orc
    cfg = shim.config
    # float32 inputs with float64 FFT/tail = the reference's pinned numpy-1.x
    # behaviour (mode M2)
    _WORKER['oracle'] = orc.OracleModel(
        cfg.obs_data, cfg.obs_var, cfg.bad_px, cfg.psf_selector.psf_images,
        cfg.psf_selector.var_images, cfg.mag_zeropoint, shim.program,
        shim.psf_index_slot, fft_upcast=True)


def _worker_eval(task):
    variant, block = task
    if _WORKER['kind'] == 'reference':
        model = _WORKER[variant]
        post = type(model).log_posterior
        with np.errstate(all='ignore'):
            return np.array([float(post(row, model=model)[0]) for row in block])
    return _WORKER['oracle'].lnlike_batch(block)


class CpuArm(object):
    def __init__(self, workload, impl='auto'):
        import multiprocessing as mp
        if impl == 'auto':
            impl = 'reference' if reference_available() else 'port'
        if impl == 'reference' and not reference_available():
            raise SystemExit('bench.py: oracle/_ref is missing -- run oracle/make_ref.py '
                             'where /root/reference exists')
        self.kind = impl
        self.cores = os.cpu_count() or 1
        model_file = workload_model_file(workload) if impl == 'reference' else None
        ctx = mp.get_context('fork')
        self.pool = ctx.Pool(self.cores, initializer=_worker_init,
                             initargs=(workload, impl, model_file))

    def evaluate(self, thetas, variant='full'):
        nblk = max(1, min(len(thetas), self.cores * 4))
        blocks = [(variant, blk) for blk in np.array_split(thetas, nblk)]
        return np.concatenate(self.pool.map(_worker_eval, blocks))

    def rate(self, thetas, seconds, variant='full'):
        """evaluations/s over about `seconds` of wall time, in calls of 4 walkers per
        core (bounded sample of the ensemble)."""
        self.evaluate(thetas[:self.cores], variant)            # warm the workers
        per_call = min(len(thetas), max(self.cores * 4, 32))
        done, start = 0, time.perf_counter()
        while True:
            lo = done % max(1, len(thetas) - per_call)
            self.evaluate(thetas[lo:lo + per_call], variant)
            done += per_call
            elapsed = time.perf_counter() - start
            if elapsed >= seconds:
                return done / elapsed, done, elapsed

    def describe(self):
        if self.kind == 'reference':
            return ('unmodified reference (oracle/_ref/psfMC via oracle/refshim.py, numpy '
                    'branches: numexpr is not installed), MultiComponentModel.log_posterior '
                    '= priors + 5 blob images + lnL, float32 storage / complex64 FFT (numpy '
                    '{})'.format(np.__version__))
        return 'oracle numpy port of the reference path, lnL only (oracle/_ref missing)'

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline_block(workload, thetas, seconds, impl):
    arm = CpuArm(workload, impl)
    try:
        rate, count, elapsed = arm.rate(thetas, seconds * (0.6 if arm.kind == 'reference'
                                                           else 1.0))
        block = {
            'value': round(rate, 1), 'unit': UNIT, 'cores': arm.cores,
            'kind': arm.kind,
            'sample': '{} evaluations of the same ensemble in {:.1f} s on {} worker '
                      'processes; {}'.format(count, elapsed, arm.cores, arm.describe())}
        if arm.kind == 'reference':
            rate2, count2, elapsed2 = arm.rate(thetas, seconds * 0.4, 'lnl_only')
            block['lnl_only'] = {
                'value': round(rate2, 1), 'unit': UNIT,
                'sample': '{} evaluations in {:.1f} s, log_priors and '
                          'point_source_subtracted patched out (like-for-like with the '
                          'GPU arm\'s lnL)'.format(count2, elapsed2)}
            block['per_core'] = round(rate / arm.cores, 2)
        return block
    finally:
        arm.close()


# ------------------------------------------------------------------- clocks --

class ClockSampler(object):
    """SM clock / power / throttle reasons sampled every few milliseconds through
    NVML (nvidia_ml_py) while the benchmark runs; `mark()` / `stop()` delimit the
    samples that fall inside the timed region."""
    REASONS = (('hw_slowdown', 'nvmlClocksEventReasonHwSlowdown', 0x8),
               ('hw_thermal_slowdown', 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
               ('sw_thermal_slowdown', 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
               ('sw_power_cap', 'nvmlClocksEventReasonSwPowerCap', 0x4))

    def __init__(self, index, period=0.004):
        self.samples = []
        self.period = period
        self.running = False
        self.t_mark = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as exc:                      # no NVML: report, do not fail
            self.nvml = None
            self.error = repr(exc)

    def _reasons(self):
        nvml = self.nvml
        for name in ('nvmlDeviceGetCurrentClocksEventReasons',
                     'nvmlDeviceGetCurrentClocksThrottleReasons'):
            func = getattr(nvml, name, None)
            if func is not None:
                try:
                    return int(func(self.handle))
                except Exception:
                    continue
        return 0

    def _loop(self):
        nvml = self.nvml
        while self.running:
            try:
                sm = nvml.nvmlDeviceGetClockInfo(self.handle, nvml.NVML_CLOCK_SM)
                power = nvml.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((time.perf_counter(), sm, power, self._reasons()))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.nvml is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['NVML unavailable']}
        t_end = time.perf_counter()
        self.running = False
        self.thread.join(timeout=1.0)
        inside = [smp for smp in self.samples
                  if self.t_mark is None or self.t_mark <= smp[0] <= t_end]
        where = 'timed region'
        if len(inside) < 3:       # very short region: use every sample taken under load
            inside, where = self.samples, 'warm-up + timed region'
        reasons = set()
        for _, _, _, mask in inside:
            for name, _, bit in self.REASONS:
                if mask & bit:
                    reasons.add(name)
        clocks = [smp[1] for smp in inside]
        return {'sm_mhz': float(np.median(clocks)) if clocks else None,
                'sm_max_mhz': float(self.sm_max),
                'power_w_max': max([smp[2] for smp in inside]) if inside else None,
                'samples': len(inside), 'window': where, 'reasons': sorted(reasons)}


# ------------------------------------------------------------------ our arm --

def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    from psfmc_b200 import BatchPool, MultiComponentModel, fp32_peak_tflops
    from psfmc_b200.distributed import (ShardedPool, shard_bounds,
                                        sharded_lnlike_device)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the engine has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # stdout must carry the one JSON line only: NCCL's log (the version banner it
        # prints at NCCL_DEBUG >= VERSION included) goes to stderr; NCCL_DEBUG itself
        # is left as the launcher set it
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)            # NCCL writes its banner to fd 1 while the communicator is built
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    walkers = args.walkers or default_walkers(args.workload)
    half = walkers // 2
    model = MultiComponentModel(build_components(args.workload),
                                precision=args.precision, devices=[local])
    engine = model.engine
    ndim = model.num_params
    nsets = 4   # distinct ensembles cycled through, so no step repeats its input
    thetas = [ensemble(model, walkers, s) for s in range(nsets)]   # same on every rank
    th_dev = [torch.from_numpy(t).to(dev) for t in thetas]
    th_pin = [torch.from_numpy(t).pin_memory() for t in thetas]
    bounds = shard_bounds(half, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    width = int(bounds[1] - bounds[0])
    even = half % world == 0
    lnl_dev = torch.zeros(2 * world * width, dtype=torch.float64, device=dev)
    send = torch.zeros(width, dtype=torch.float64, device=dev)
    lnl_own = torch.empty(walkers, dtype=torch.float64, device=dev)
    lnl_pin = torch.empty(walkers, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_device(s):
        """One emcee iteration of THE ensemble, device-resident: per half-ensemble this
        rank's shard through the engine, then the lnL all-gather (NCCL) every rank needs
        before it can propose the other half."""
        th = th_dev[s % nsets]
        for h in range(2):     # two sequentially dependent half-ensembles
            recv = lnl_dev[h * world * width:(h + 1) * world * width]
            if world == 1:
                engine.lnlike_device(th.data_ptr() + h * half * ndim * 8, half, ndim,
                                     recv.data_ptr(), stream=stream.cuda_stream)
            else:
                sharded_lnlike_device(engine, th, half, ndim, send, recv, stream,
                                      row_offset=h * half)

    # lnL gather over peer memory (CUDA IPC mailboxes, psfmc_lnlike_batch_exchange) instead
    # of NCCL: the default for N > 1; PSFMC_BENCH_GATHER=nccl keeps the all_gather
    exchange = None
    gather_mode = os.environ.get('PSFMC_BENCH_GATHER', 'peer')
    if world > 1 and gather_mode == 'peer':
        from psfmc_b200.distributed import PeerExchange
        try:
            exchange = PeerExchange(engine, half)
        except Exception as exc:            # no peer access on this box: fall back, say so
            print('bench.py: peer exchange unavailable ({}), using NCCL'.format(exc),
                  file=sys.stderr)
            exchange = None
    lnl_gath = torch.zeros(2 * half, dtype=torch.float64, device=dev)

    def step_peer(s, copy=False):
        # (the gathered values stay in the mailbox, like NCCL's stay in its receive buffer)
        th = th_dev[s % nsets]
        for h in range(2):
            exchange.lnlike(th, half, ndim,
                            lnl_gath[h * half:(h + 1) * half] if copy else None, stream,
                            row_offset=h * half)

    def step_replica(s):
        """Weak-scaling companion: this rank evaluates a whole ensemble on its own."""
        th = th_dev[s % nsets]
        for h in range(2):
            engine.lnlike_device(th.data_ptr() + h * half * ndim * 8, half, ndim,
                                 lnl_own.data_ptr() + h * half * 8,
                                 stream=stream.cuda_stream)

    pool = ShardedPool(model) if world > 1 else BatchPool(model)
    inside_map = [0.0]      # seconds spent inside pool.map itself (step_pool_map)
    lnlike_pool = None
    if world > 1:
        from psfmc_b200.distributed import ShardedEvaluator
        lnlike_pool = ShardedEvaluator(lambda rows: engine.lnlike(rows))

    def step_host(s):
        """The same iteration through the C ABI with host buffers: pinned theta in,
        lnL out (N > 1: each rank its shard, then the gather to every rank's host)."""
        th = th_pin[s % nsets].numpy()
        out = lnl_pin.numpy()
        for h in range(2):
            rows = th[h * half:(h + 1) * half]
            if world == 1:
                engine.lnlike(rows, out=out[h * half:(h + 1) * half])
            elif exchange is not None:
                # one library call per rank: its share of the rows up, lnL of all rows
                # gathered over peer memory, down to the host (psfmc_lnpost_batch_sharded)
                engine.lnpost_sharded(None, rows, out=out[h * half:(h + 1) * half])
            else:
                out[h * half:(h + 1) * half] = lnlike_pool(rows)

    def step_host_nccl(s):
        """N > 1, the torch.distributed form: ShardedEvaluator (NCCL all_gather)."""
        th = th_pin[s % nsets].numpy()
        out = lnl_pin.numpy()
        for h in range(2):
            out[h * half:(h + 1) * half] = lnlike_pool(th[h * half:(h + 1) * half])

    def step_posterior(s):
        th = thetas[s % nsets]
        for h in range(2):
            pool.map_batch(None, th[h * half:(h + 1) * half])

    def step_pool_map(s):
        """emcee 2.x's own protocol (EnsembleSampler._get_lnprob): a list of row views
        in, a list of (lnpost, blob) out, the floats picked out again."""
        th = thetas[s % nsets]
        for h in range(2):
            p = th[h * half:(h + 1) * half]
            rows = [p[i] for i in range(len(p))]
            t_in = time.perf_counter()
            results = list(pool.map(None, rows))
            inside_map[0] += time.perf_counter() - t_in
            np.array([float(r[0]) for r in results])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(value):
        if world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def host_timed(step, steps):
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            step(s)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    step_nccl = step_device
    if exchange is not None:
        # the two gathers must agree bit for bit before anything is timed
        step_nccl(0)
        step_peer(0, copy=True)
        torch.cuda.synchronize()
        a = torch.cat([lnl_dev[:half], lnl_dev[world * width:world * width + half]]) \
            if even else None
        if a is not None and not torch.equal(a.nan_to_num(), lnl_gath.nan_to_num()):
            raise SystemExit('bench.py: peer exchange and NCCL gather disagree')
        step_device = step_peer
    sampler = ClockSampler(local)
    warm = max(args.warmup, 3)
    for s in range(warm):
        step_device(s)
        step_host(s)
        step_posterior(s)
    step_pool_map(0)
    torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) -------------------------------
    launches0 = engine.info()['launches_total']
    engine.profile(True)
    engine.profile_read()
    barrier()
    sampler.mark()
    total_ms = 0.0
    for s in range(args.steps):
        flush.fill_(s & 0xFF)                  # evict L2 between timed steps
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device(s)
        e1.record(stream)
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    barrier()
    launches = engine.info()['launches_total'] - launches0
    clocks = sampler.stop()
    kernel_ms, kernel_launches = engine.profile_read()
    engine.profile(False)
    total_ms_local = total_ms
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = walkers / (ms_per_step * 1e-3)

    # ---- the same with the NCCL all_gather (secondary, when the peer exchange is used) --
    nccl_line = None
    if exchange is not None:
        nsteps = max(10, args.steps // 2)
        for s in range(3):
            step_nccl(s)
        engine.profile(True)        # same conditions as the headline loop
        barrier()
        nccl_ms = 0.0
        for s in range(nsteps):
            flush.fill_(s & 0xFF)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_nccl(s)
            e1.record(stream)
            e1.synchronize()
            nccl_ms += e0.elapsed_time(e1)
        engine.profile_read()
        engine.profile(False)
        nccl_ms = max_over_ranks(nccl_ms) / nsteps
        nccl_line = {'value': round(walkers / (nccl_ms * 1e-3), 1), 'unit': UNIT,
                     'ms_per_step': round(nccl_ms, 4),
                     'gather': 'NCCL all_gather_into_tensor'}

    # ---- replicas (weak scaling, no gather) ---------------------------------
    replicas = None
    if world > 1:
        rsteps = max(10, args.steps // 4)
        for s in range(3):
            step_replica(s)
        barrier()
        rep_ms = 0.0
        for s in range(rsteps):
            flush.fill_(s & 0xFF)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_replica(s)
            e1.record(stream)
            e1.synchronize()
            rep_ms += e0.elapsed_time(e1)
        rep_ms = max_over_ranks(rep_ms) / rsteps
        replicas = {'value': round(world * walkers / (rep_ms * 1e-3), 1), 'unit': UNIT,
                    'ms_per_step': round(rep_ms, 4), 'scaling': 'weak',
                    'note': 'every rank evaluates its own {}-walker ensemble, no '
                            'gather'.format(walkers)}

    # ---- end to end through the C ABI with host buffers (`e2e`) -------------
    rescued0 = engine.info()['rescued_total']
    e2e_s = host_timed(step_host, args.steps)
    e2e_value = walkers * args.steps / e2e_s
    e2e_nccl = None
    if world > 1 and exchange is not None:
        for s in range(3):
            step_host_nccl(s)
        e2e_nccl = walkers * args.steps / host_timed(step_host_nccl, args.steps)
    rescued_per_step = (engine.info()['rescued_total'] - rescued0) / float(args.steps)
    # the same loop with the float64 rescue of non-finite float32 results switched
    # off (for information: prior-drawn ensembles contain a few such walkers, the
    # walkers of a converged chain do not)
    e2e_raw = None
    if args.precision == 'fp32' and world == 1:
        raw_engine = MultiComponentModel(build_components(args.workload), precision='fp32',
                                         devices=[local], fp64_rescue=False).engine
        out = lnl_pin.numpy()

        def step_raw(s):
            th = th_pin[s % nsets].numpy()
            for h in range(2):
                raw_engine.lnlike(th[h * half:(h + 1) * half],
                                  out=out[h * half:(h + 1) * half])
        for s in range(3):
            step_raw(s)
        e2e_raw = walkers * args.steps / host_timed(step_raw, args.steps)
        raw_engine.close()
    post_s = host_timed(step_posterior, args.steps)
    map_steps = max(10, args.steps // 3)
    inside_map[0] = 0.0
    map_s = host_timed(step_pool_map, map_steps)
    map_inside_s = max_over_ranks(inside_map[0])

    # ---- the whole sampler iteration (stretch-move proposals, priors, lnL, acceptance)
    # inside the library: psfmc_ensemble_run, what this package's sampler calls; beside it
    # the same sampler with its numpy loop, at the bench ensemble and at the reference
    # example's 250 walkers (examples/run_example.py:9)
    # (N > 1: every rank runs the same seeded sampler on a ShardedPool; the library loop
    # then evaluates this rank's share of every half-ensemble and gathers the lnL over
    # peer memory, PSFMC_ENS_SHARDED)
    loop = None
    if True:
        from psfmc_b200.sampler import EnsembleSampler
        loop = {}

        def sampler_rate(nwalk, native, iters):
            os.environ['PSFMC_NATIVE_SAMPLER'] = '1' if native else '0'
            start = thetas[0][:nwalk]
            smp = EnsembleSampler(nwalk, ndim, model.log_posterior, kwargs={'model': model},
                                  pool=(ShardedPool(model) if world > 1 else BatchPool(model)),
                                  live_dangerously=True)
            smp._random.seed(7)
            pos, lnp, _ = smp.run_mcmc(start, 16)   # (buffers sized, half-step graphs captured)
            smp.reset()
            barrier()
            t0 = time.perf_counter()
            smp.run_mcmc(pos, iters, lnprob0=lnp)
            return nwalk * iters / max_over_ranks(time.perf_counter() - t0)
        try:
            loop['library'] = round(sampler_rate(walkers, True, args.steps), 1)
            loop['numpy_loop'] = round(sampler_rate(walkers, False, max(10, args.steps // 3)), 1)
            small = min(250, walkers)
            loop['walkers_small'] = small
            loop['library_small'] = round(sampler_rate(small, True, 4 * args.steps), 1)
            loop['numpy_loop_small'] = round(sampler_rate(small, False, args.steps), 1)
            loop['prior_columns_in_python'] = bool(model._sampler_plan and
                                                   model._sampler_plan['python_columns'])
            loop['note'] = ('stretch-move iterations of one ensemble end to end (proposals, '
                            'log-priors, lnL through host buffers, acceptance, chain storage): '
                            'library = psfmc_ensemble_run, numpy_loop = the same sampler with '
                            'its Python loop; *_small = the reference example\'s ensemble size'
                            + ('' if world == 1 else
                               '; {} ranks, each its share of every half-ensemble: library = '
                               'PSFMC_ENS_SHARDED (lnL over peer memory), numpy_loop = '
                               'ShardedPool.map_batch (NCCL all_gather)'.format(world)))
        finally:
            os.environ.pop('PSFMC_NATIVE_SAMPLER', None)

    info = engine.info()
    result = None
    if rank == 0:
        peak_probe = fp32_peak_tflops(local)
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fobj:
                peaks = json.load(fobj)
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        # dominant kernel: events around each launch on the launching stream
        kernel_us = 1e3 * kernel_ms / max(kernel_launches, 1)
        evals_per_launch = hi - lo
        achieved = info['flops_per_eval'] * evals_per_launch / (kernel_us * 1e-6) / 1e12
        fft_achieved = info['fft_flops_per_eval'] * evals_per_launch / (kernel_us * 1e-6) / 1e12
        # dram__bytes and the executed FP32 operation count of one launch of the dominant
        # kernel, from the committed ncu capture of the same workload (profiles/)
        ncu = {}
        ncu_file = {('c1', 1): 'r2_fused_ncu_summary.json',
                    ('c3', 2): 'r2_cluster256_ncu_summary.json'}.get(
                        (args.workload, info['path']))
        if ncu_file and args.walkers == 0 and world == 1:
            for name in (ncu_file.replace('r2_', 'r2b_'), ncu_file,
                         ncu_file.replace('r2_', 'r1_')):
                try:
                    with open(os.path.join(ROOT, 'profiles', name)) as fobj:
                        ncu = json.load(fobj)
                    ncu['file'] = 'profiles/' + name
                    break
                except (OSError, ValueError):
                    continue
        traffic = ncu.get('dram_bytes_per_launch')
        executed = ncu.get('executed_fp32_flop_per_launch')
        roofline = {
            'bound': 'fp32',
            'achieved': round(achieved, 3), 'peak': round(peak_probe, 2),
            'unit': 'TFLOP/s', 'frac': round(achieved / peak_probe, 4),
            'traffic': traffic,
            'kernel': {1: 'fused_lnlike_kernel', 2: 'cluster256_lnlike_kernel',
                       3: 'fused_lnlike_kernel<FWD> + tiled_combine_kernel + '
                          'fused_lnlike_kernel<INV>'}.get(
                info['path'], 'rows_fwd + cols + rows_inv'),
            'kernel_us_per_launch': round(kernel_us, 2),
            'kernel_launches_timed': int(kernel_launches),
            'evals_per_launch': evals_per_launch,
            'kernel_share_of_step': round(kernel_ms / total_ms_local, 4),
            'note': 'FP32 CUDA-core bound path (north_star: no tensor cores; SURVEY.md 8d). '
                    'achieved = (10 N log2 N + (30 n_sersic + 16) N) FLOP/eval x evals per '
                    'launch / mean launch duration of the dominant kernel (CUDA events on the '
                    'launching stream inside the timed region); this nominal count covers the '
                    'full frame although row groups without an unmasked pixel skip their '
                    'inverse transform. peak = FP32 FMA throughput probed in this run, the '
                    'better of scalar FFMA and packed FFMA2 (MEASURED_PEAKS.json has no FP32 '
                    'entry). FFT butterflies are mostly adds (1 FLOP per lane and clock), so '
                    'the practical ceiling of this path is about half of the FMA peak '
                    '(DESIGN.md).',
            'fft_stage_achieved': round(fft_achieved, 3),
            'fft_stage_frac': round(fft_achieved / peak_probe, 4),
            'flops_per_eval': info['flops_per_eval'],
            'hbm': {'achieved': round(info['hbm_bytes_per_eval'] * evals_per_launch
                                      / (kernel_us * 1e-6) / 1e9, 2),
                    'peak': hbm_peak, 'unit': 'GB/s',
                    'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback',
                    'bytes_per_eval': info['hbm_bytes_per_eval']},
            'engine_path': {1: 'fused', 2: 'fused-cluster4', 3: 'tiled-4x4'}.get(
                info['path'], 'staged'),
        }
        if executed:
            # FP32 operations the kernel really executed (ncu op mix: FADD2/FMUL2 = 2,
            # FFMA2 = 4, FADD/FMUL = 1, FFMA = 2 per thread) -- stated beside the nominal
            # count above; scaled to this run's launch duration
            ex_tflops = executed / (kernel_us * 1e-6) / 1e12 * (
                evals_per_launch / float(ncu.get('evals_per_launch', evals_per_launch)))
            roofline['executed'] = {'achieved': round(ex_tflops, 3),
                                    'frac': round(ex_tflops / peak_probe, 4),
                                    'flop_per_launch': executed, 'source': ncu.get('file')}
        result = {
            'metric': METRIC, 'value': round(value, 1), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': warm,
            'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32 render+FFT / f64 accumulate' if args.precision == 'fp32' else 'f64',
            'data': 'synthetic walkers drawn from the model priors (seeded); '
                    + ('J0005-0006 example frames' if args.workload == 'c1'
                       else 'synthetic frames'),
            'config': {'workload': workload_description(args.workload, walkers),
                       'walkers': walkers, 'batch_per_launch': half,
                       'ndim': ndim, 'frame': list(engine.shape)},
            'l2': 'flushed (256 MiB write) between timed steps',
            'sharding': {'ranks': world, 'rows_per_rank_per_half': hi - lo,
                         'gather': 'none' if world == 1 else (
                             'peer memory: every rank stores its lnL into all ranks\' '
                             'mailboxes (CUDA IPC over NVLink) + one flag per peer, {} doubles '
                             'per half-ensemble, inside the timed region'.format(half)
                             if exchange is not None else
                             'NCCL all_gather_into_tensor of {} doubles per '
                             'half-ensemble, inside the timed region'.format(half)),
                         'even': bool(even)},
            'e2e': {'value': round(e2e_value, 1), 'unit': UNIT,
                    'h2d_bytes_per_step': (hi - lo) * 2 * ndim * 8,
                    'd2h_bytes_per_step': walkers * 8,
                    'timer': 'host perf_counter around blocking calls (psfmc_lnlike_batch; '
                             'N > 1: psfmc_lnpost_batch_sharded -- this rank\'s share up, lnL '
                             'of all rows gathered over peer memory, down to every rank\'s '
                             'host), max over ranks',
                    'with_nccl_gather': None if e2e_nccl is None else round(e2e_nccl, 1),
                    'with_python_priors': round(walkers * args.steps / post_s, 1),
                    'pool_map': round(walkers * map_steps / map_s, 1),
                    'pool_map_inside': round(walkers * map_steps / map_inside_s, 1),
                    'pool_map_note': 'emcee 2.x list protocol through {}.map, priors included: '
                                     'pool_map also times emcee\'s own side of it (the list of '
                                     'row views it builds, the floats it picks out of the '
                                     'result tuples), pool_map_inside only the map call'.format(
                                         type(pool).__name__),
                    'sampler_loop': loop,
                    'fp64_rescued_walkers_per_step': round(rescued_per_step, 2),
                    'without_fp64_rescue': None if e2e_raw is None else round(e2e_raw, 1)},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
        }
        if replicas:
            result['replicas_weak'] = replicas
        if nccl_line:
            result['with_nccl_gather'] = nccl_line
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    model.engine.close()
    if rank == 0 and not args.no_cpu_baseline:
        result['cpu_baseline'] = cpu_baseline_block(args.workload, thetas[0],
                                                    args.cpu_seconds, args.cpu_impl)
    if rank == 0:
        print(json.dumps(result))


# ------------------------------------------------------------ reference arm --

def run_reference(args):
    """The reference's CPU implementation of the path on the host cores: the
    UNMODIFIED reference's MultiComponentModel.log_posterior (oracle/_ref, see the
    module docstring) on all host cores; the oracle port if oracle/_ref is absent.
    Each step evaluates a bounded sample of the ensemble so that the whole run ends
    within minutes. Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    walkers = args.walkers or default_walkers(args.workload)
    shim = host_model(args.workload)
    ndim = shim.num_params
    thetas = ensemble(shim, walkers, 0)

    arm = CpuArm(args.workload, args.cpu_impl)
    # bounded sample per step so that the whole run ends within minutes
    probe = arm.cores * 2
    arm.evaluate(thetas[:arm.cores])
    t0 = time.perf_counter()
    arm.evaluate(thetas[:probe])
    per_eval = (time.perf_counter() - t0) / probe
    total_steps = args.steps + args.warmup
    sample = int(min(walkers, max(arm.cores * 4, 60.0 / max(total_steps, 1) / per_eval)))
    for s in range(args.warmup):
        arm.evaluate(thetas[:sample])
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = (s * sample) % max(1, walkers - sample)
        arm.evaluate(thetas[lo:lo + sample])
    elapsed = time.perf_counter() - t0
    rate = sample * args.steps / elapsed
    extra = {}
    if arm.kind == 'reference':
        rate2, count2, elapsed2 = arm.rate(thetas, 8.0, 'lnl_only')
        extra = {'lnl_only': {'value': round(rate2, 1), 'unit': UNIT,
                              'sample': '{} evaluations in {:.1f} s, log_priors and '
                                        'point_source_subtracted patched out'.format(
                                            count2, elapsed2)},
                 'per_core': round(rate / arm.cores, 2)}
    arm.close()
    sample_text = ('{} of the {} evaluations of a step, per step, on {} worker processes; '
                   '{}'.format(sample, walkers, arm.cores, arm.describe()))
    baseline = {'value': round(rate, 1), 'unit': UNIT, 'cores': arm.cores,
                'kind': arm.kind, 'sample': sample_text}
    baseline.update(extra)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': round(rate, 1), 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(1e3 * elapsed / args.steps, 3), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32 storage / complex64 FFT (numpy {})'.format(np.__version__)
                 if arm.kind == 'reference' else 'f64 (numpy, float32 storage)',
        'data': 'synthetic walkers drawn from the model priors (seeded); '
                + ('J0005-0006 example frames' if args.workload == 'c1'
                   else 'synthetic frames'),
        'config': {'workload': workload_description(args.workload, walkers),
                   'walkers': walkers, 'batch_per_launch': walkers // 2,
                   'ndim': ndim, 'frame': list(shim.config.obs_data.shape)},
        'cpu_baseline': baseline,
        'e2e': {'value': round(rate, 1), 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }))


if __name__ == '__main__':
    cli = parse_args()
    if cli.impl == 'reference':
        run_reference(cli)
    else:
        run_ours(cli)
