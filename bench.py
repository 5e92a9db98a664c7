#!/usr/bin/env python
"""
bench.py -- model lnL evaluations per second (walkers x iterations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c3|c4|s128]
                    [--walkers NW] [--impl ours|reference] [--precision fp32|fp64]

One "step" is one emcee iteration of an NW-walker ensemble: the two sequentially
dependent half-ensemble batches emcee 2.x maps per iteration (SURVEY.md section
3.2), i.e. NW lnL evaluations through the hot path. The default workload is the
J0005-0006 quasar+host model (C1 frame, 128^2, Sky + PointSource + 2 Sersic, D=18)
for a 4096-walker ensemble per GPU -- the configuration BASELINE.json's north_star
quotes its target on. Walkers shard over GPUs with no data-path collective
(weak scaling: every rank evaluates its own NW-walker ensemble).

Prints ONE JSON line (rank 0). `value` = device-resident throughput (theta and
lnL stay in HBM, CUDA events on the launching stream); `e2e` = the same metric
through the C ABI with host buffers (pinned theta in, lnL out, copies inside the
timed region).
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
                    help='ensemble size per GPU (default: 4096 for c1/s128/c4, 1024 for c3)')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'fp64'])
    ap.add_argument('--cpu-seconds', type=float, default=12.0,
                    help='budget of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
    return '{}; {}-walker ensemble per GPU, 2 half-ensemble batches per step'.format(
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


def build_oracle(model_like):
    """Oracle (CPU port of the reference path) over a model's arrays + program.
    float32 inputs with float64 FFT/tail = the reference's pinned numpy-1.x
    behaviour (mode M2)."""
    from oracle import psfmc_oracle as orc
    cfg = model_like['config']
    return orc.OracleModel(cfg.obs_data, cfg.obs_var, cfg.bad_px,
                           cfg.psf_selector.psf_images, cfg.psf_selector.var_images,
                           cfg.mag_zeropoint, model_like['program'],
                           model_like['psf_index_slot'], fft_upcast=True)


# ------------------------------------------------------- CPU baseline (port) --

_WORKER = {}


def _worker_init(workload):
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
        os.environ[var] = '1'
    from psfmc_b200.components import Configuration
    from psfmc_b200.program import compile_program
    comps = build_components(workload)
    config = [c for c in comps if isinstance(c, Configuration)][0]
    rest = [c for c in comps if c is not config] + [config.psf_selector]
    program, psf_slot, _ = compile_program(rest)
    _WORKER['oracle'] = build_oracle({'config': config, 'program': program,
                                      'psf_index_slot': psf_slot})


def _worker_eval(block):
    return _WORKER['oracle'].lnlike_batch(block)


class CpuPort(object):
    """The oracle port on all host cores: a multiprocessing pool with one model per
    worker (the reference's own parallel hook is emcee's pool.map; its model object
    is not picklable, psfMC/fitting.py:55, hence one model per worker)."""

    def __init__(self, workload):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        ctx = mp.get_context('fork')
        self.pool = ctx.Pool(self.cores, initializer=_worker_init,
                             initargs=(workload,))

    def evaluate(self, thetas):
        nblk = max(1, min(len(thetas), self.cores * 4))
        blocks = np.array_split(thetas, nblk)
        return np.concatenate(self.pool.map(_worker_eval, blocks))

    def close(self):
        self.pool.close()
        self.pool.join()


def time_cpu_port(workload, thetas, seconds):
    port = CpuPort(workload)
    try:
        port.evaluate(thetas[:port.cores])            # warm the workers
        per_call = max(port.cores * 4, 32)
        done, start = 0, time.perf_counter()
        while True:
            lo = done % max(1, len(thetas) - per_call)
            port.evaluate(thetas[lo:lo + per_call])
            done += per_call
            elapsed = time.perf_counter() - start
            if elapsed >= seconds:
                break
        return done / elapsed, port.cores, done, elapsed
    finally:
        port.close()


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
    from psfmc_b200 import MultiComponentModel, fp32_peak_tflops
    from psfmc_b200.synthetic import draw_walkers_fast

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the engine has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG >= VERSION; stdout must
        # carry the one JSON line only: send NCCL's log to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            del os.environ['NCCL_DEBUG']
        dist.init_process_group('nccl', device_id=dev)

    walkers = args.walkers or default_walkers(args.workload)
    half = walkers // 2
    model = MultiComponentModel(build_components(args.workload),
                                precision=args.precision, devices=[local])
    engine = model.engine
    ndim = model.num_params
    nsets = 4   # distinct ensembles cycled through, so no step repeats its input
    thetas = [draw_walkers_fast(model, walkers, seed=1000 * rank + s) for s in range(nsets)]
    th_dev = [torch.from_numpy(t).to(dev) for t in thetas]
    th_pin = [torch.from_numpy(t).pin_memory() for t in thetas]
    lnl_dev = torch.empty(walkers, dtype=torch.float64, device=dev)
    lnl_pin = torch.empty(walkers, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_device(s):
        th = th_dev[s % nsets]
        base = th.data_ptr()
        for h in range(2):     # two sequentially dependent half-ensembles
            engine.lnlike_device(base + h * half * ndim * 8, half, ndim,
                                 lnl_dev.data_ptr() + h * half * 8,
                                 stream=stream.cuda_stream)

    def step_host(s):
        th = th_pin[s % nsets].numpy()
        out = lnl_pin.numpy()
        for h in range(2):
            engine.lnlike(th[h * half:(h + 1) * half], out=out[h * half:(h + 1) * half])

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

    sampler = ClockSampler(local)
    for s in range(max(args.warmup, 3)):
        step_device(s)
        step_host(s)
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
    value = world * walkers / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers (`e2e`) -------------
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(s)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * walkers * args.steps / e2e_s
    rescued_per_step = engine.info()['rescued_total'] / float(
        args.steps + max(args.warmup, 3))
    # the same loop with the float64 rescue of non-finite float32 results switched
    # off (for information: prior-drawn ensembles contain a few such walkers, the
    # walkers of a converged chain do not)
    e2e_raw = None
    if args.precision == 'fp32':
        raw_engine = MultiComponentModel(build_components(args.workload), precision='fp32',
                                         devices=[local], fp64_rescue=False).engine
        out = lnl_pin.numpy()
        for s in range(3):
            raw_engine.lnlike(th_pin[0].numpy()[:half], out=out[:half])
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            th = th_pin[s % nsets].numpy()
            for h in range(2):
                raw_engine.lnlike(th[h * half:(h + 1) * half],
                                  out=out[h * half:(h + 1) * half])
        torch.cuda.synchronize()
        e2e_raw = world * walkers * args.steps / max_over_ranks(time.perf_counter() - t0)
        raw_engine.close()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        th = thetas[s % nsets]
        for h in range(2):
            model.log_posterior_batch(th[h * half:(h + 1) * half])
    post_s = max_over_ranks(time.perf_counter() - t0)

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
        evals_per_launch = half
        achieved = info['flops_per_eval'] * evals_per_launch / (kernel_us * 1e-6) / 1e12
        fft_achieved = info['fft_flops_per_eval'] * evals_per_launch / (kernel_us * 1e-6) / 1e12
        # dram__bytes of one launch of the dominant kernel from the committed ncu capture
        # of the same workload (profiles/), when there is one
        ncu = {}
        ncu_file = {('c1', 1): 'r1_fused_ncu_summary.json',
                    ('c3', 2): 'r1_cluster256_ncu_summary.json'}.get(
                        (args.workload, info['path']))
        if ncu_file and args.walkers == 0:
            try:
                with open(os.path.join(ROOT, 'profiles', ncu_file)) as fobj:
                    ncu = json.load(fobj)
            except (OSError, ValueError):
                pass
        traffic = ncu.get('dram_bytes_per_launch')
        roofline = {
            'bound': 'fp32',
            'achieved': round(achieved, 3), 'peak': round(peak_probe, 2),
            'unit': 'TFLOP/s', 'frac': round(achieved / peak_probe, 4),
            'traffic': traffic,
            'kernel': {1: 'fused_lnlike_kernel', 2: 'cluster256_lnlike_kernel'}.get(
                info['path'], 'rows_fwd + cols + rows_inv'),
            'kernel_us_per_launch': round(kernel_us, 2),
            'kernel_launches_timed': int(kernel_launches),
            'evals_per_launch': evals_per_launch,
            'kernel_share_of_step': round(kernel_ms / total_ms_local, 4),
            'note': 'FP32 CUDA-core bound path (north_star: no tensor cores; SURVEY.md 8d). '
                    'achieved = (10 N log2 N + (30 n_sersic + 16) N) FLOP/eval x evals per '
                    'launch / mean launch duration of the dominant kernel (CUDA events on the '
                    'launching stream inside the timed region); peak = FP32 FMA throughput '
                    'probed in this run (MEASURED_PEAKS.json has no FP32 entry). FFT '
                    'butterflies are mostly adds, so the practical ceiling of this path is '
                    'about half of the FMA peak (DESIGN.md).',
            'fft_stage_achieved': round(fft_achieved, 3),
            'fft_stage_frac': round(fft_achieved / peak_probe, 4),
            'flops_per_eval': info['flops_per_eval'],
            'hbm': {'achieved': round(info['hbm_bytes_per_eval'] * evals_per_launch
                                      / (kernel_us * 1e-6) / 1e9, 2),
                    'peak': hbm_peak, 'unit': 'GB/s',
                    'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback',
                    'bytes_per_eval': info['hbm_bytes_per_eval']},
            'engine_path': {1: 'fused', 2: 'fused-cluster4'}.get(info['path'], 'staged'),
        }
        result = {
            'metric': METRIC, 'value': round(value, 1), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 render+FFT / f64 accumulate' if args.precision == 'fp32' else 'f64',
            'data': 'synthetic walkers drawn from the model priors (seeded); '
                    + ('J0005-0006 example frames' if args.workload == 'c1'
                       else 'synthetic frames'),
            'config': {'workload': workload_description(args.workload, walkers),
                       'walkers_per_gpu': walkers, 'batch_per_launch': half,
                       'ndim': ndim, 'frame': list(engine.shape),
                       'l2': 'flushed (256 MiB write) between timed steps'},
            'e2e': {'value': round(e2e_value, 1), 'unit': UNIT,
                    'h2d_bytes_per_step': walkers * ndim * 8,
                    'd2h_bytes_per_step': walkers * 8,
                    'timer': 'host perf_counter around blocking C-ABI calls '
                             '(psfmc_lnlike_batch), max over ranks',
                    'with_python_priors': round(world * walkers * args.steps / post_s, 1),
                    'fp64_rescued_walkers_per_step': round(rescued_per_step, 2),
                    'without_fp64_rescue': None if e2e_raw is None else round(e2e_raw, 1)},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
        }
    if rank == 0 and not args.no_cpu_baseline:
        rate, cores, count, elapsed = time_cpu_port(args.workload, thetas[0],
                                                    args.cpu_seconds)
        result['cpu_baseline'] = {
            'value': round(rate, 1), 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '{} evaluations of the same workload in {:.1f} s on {} worker '
                      'processes (oracle numpy port of the reference path, lnL only: no '
                      'priors, no point-source-subtracted blob)'.format(count, elapsed, cores)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


# ------------------------------------------------------------ reference arm --

def run_reference(args):
    """The reference's CPU implementation of the path on the host cores. The
    reference is pure Python and cannot travel to the GPU box, so this runs its
    oracle port (oracle/psfmc_oracle.py, pinned bit-for-bit against the reference
    by tests/golden/make_golden.py) on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from psfmc_b200.components import Configuration
    from psfmc_b200.program import compile_program
    walkers = args.walkers or default_walkers(args.workload)
    comps = build_components(args.workload)
    config = [c for c in comps if isinstance(c, Configuration)][0]
    rest = [c for c in comps if c is not config] + [config.psf_selector]
    _, _, ndim = compile_program(rest)

    class _Shim(object):     # priors only, to draw the same walkers as the GPU arm
        pass
    from psfmc_b200.synthetic import draw_walkers_fast
    shim = _Shim()
    shim.components = rest
    shim.num_params = ndim

    def log_priors_batch(thetas):
        total, start = np.zeros(len(thetas)), 0
        for comp in rest:
            count = comp.num_stochastics()
            total = total + comp.log_priors_batch(thetas[:, start:start + count])
            start += count
        return total
    shim.log_priors_batch = log_priors_batch
    thetas = draw_walkers_fast(shim, walkers, seed=0)

    port = CpuPort(args.workload)
    # bounded sample per step so that the whole run ends within minutes
    probe = port.cores * 4
    t0 = time.perf_counter()
    port.evaluate(thetas[:probe])
    per_eval = (time.perf_counter() - t0) / probe
    total_steps = args.steps + args.warmup
    sample = int(min(walkers, max(port.cores * 4, 60.0 / max(total_steps, 1) / per_eval)))
    for s in range(args.warmup):
        port.evaluate(thetas[:sample])
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = (s * sample) % max(1, walkers - sample)
        port.evaluate(thetas[lo:lo + sample])
    elapsed = time.perf_counter() - t0
    port.close()
    rate = sample * args.steps / elapsed
    sample_text = ('{} of the {} evaluations of a step, per step, on {} worker processes '
                   '(oracle numpy port, lnL only)'.format(sample, walkers, port.cores))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': round(rate, 1), 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(1e3 * elapsed / args.steps, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 (numpy, float32 storage)',
        'data': 'synthetic walkers drawn from the model priors (seeded)',
        'config': {'workload': workload_description(args.workload, walkers),
                   'walkers_per_gpu': walkers, 'batch_per_launch': walkers // 2,
                   'ndim': ndim, 'frame': list(config.obs_data.shape)},
        'cpu_baseline': {'value': round(rate, 1), 'unit': UNIT, 'cores': port.cores,
                         'kind': 'port', 'sample': sample_text},
        'e2e': {'value': round(rate, 1), 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }))


if __name__ == '__main__':
    cli = parse_args()
    if cli.impl == 'reference':
        run_reference(cli)
    else:
        run_ours(cli)
